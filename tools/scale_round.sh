#!/bin/bash
# usage (under `gpurun --gpus 8`): tools/scale_round.sh <tag> — the driver's scaling run: bench.py on 2, 4 and 8 ranks, one JSON line each
TAG=$1
for N in 2 4 8; do
  timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29600 + N)) bench.py --gpus $N --steps 60 --warmup 3 \
    > gpurun_out/bench_n${N}_$TAG.json 2> gpurun_out/bench_n${N}_$TAG.err; echo "N=$N rc=$?"
  python - <<PY
import json
d = json.load(open("gpurun_out/bench_n${N}_$TAG.json"))
c = d["c5"]
print("N=$N value", round(d["value"], 1), "ms", round(d["ms_per_step"], 3), "latency", round(d["frame_latency_ms"], 3), "e2e", round(d["e2e"]["value"], 1),
      "blocking ms", round(d["e2e"]["blocking"]["ms_per_frame"], 3), "parity", d["roofline"].get("parity_rows_within_1_255"), "lanes", d["run"]["lanes"])
print("   c5 value", round(c["value"], 1), "ms", round(c["ms_per_step"], 2), "latency", round(c["frame_latency_ms"], 2), "e2e", round(c["e2e"]["value"], 1))
PY
done
