#!/bin/bash
# usage (under `gpurun --gpus N`): tools/ring_ab.sh <N> <tag> — end to end at N ranks with the host ring and with the device ring, same box
N=$1; TAG=$2
for RING in host device; do
  timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29700 + N)) bench.py --gpus $N --steps 60 --warmup 3 \
    --e2e-ring $RING --no-cpu-baseline > gpurun_out/ring_${RING}_n${N}_$TAG.json 2> gpurun_out/ring_${RING}_n${N}_$TAG.err; echo "N=$N ring=$RING rc=$?"
  tail -3 gpurun_out/ring_${RING}_n${N}_$TAG.err
  python - <<PY
import json
d = json.load(open("gpurun_out/ring_${RING}_n${N}_$TAG.json"))
c = d["c5"]
print("[$RING ring] N=$N value", round(d["value"], 1), "ms", round(d["ms_per_step"], 3), "e2e", round(d["e2e"]["value"], 1), "e2e ms", round(d["e2e"]["ms_per_step"], 3),
      "blocking ms", round(d["e2e"]["blocking"]["ms_per_frame"], 3), "parity", d["roofline"].get("parity_rows_within_1_255"))
print("   c5 value", round(c["value"], 1), "e2e", round(c["e2e"]["value"], 1))
PY
done
