#!/bin/bash
# usage: tools/ncu_traverse.sh <tag> [kernel regex] [workload] — bench line + ncu --set full of the first 4 matching launches of a warm frame (1 GPU, RTB_LANES=1)
TAG=$1; KRE=${2:-k_traverse}; WL=${3:-c4}
python bench.py --workload $WL --steps 30 --warmup 3 --no-cpu-baseline > gpurun_out/bench_${WL}_$TAG.json 2> gpurun_out/bench_${WL}_$TAG.err; echo "bench rc=$?"
python - <<PY
import json
d = json.load(open("gpurun_out/bench_${WL}_$TAG.json"))
r = d["roofline"]
print("value", round(d["value"], 1), "ms", round(d["ms_per_step"], 3), "kernels", d["kernel_ms_per_frame"], "nodes/ray", r.get("gpu_nodes_fetched_per_ray"), "tris/ray", r.get("gpu_tris_tested_per_ray"))
PY
export RTB_LANES=1
CMD="python bench.py --workload $WL --steps 2 --warmup 1 --no-cpu-baseline"
$CMD > gpurun_out/plain_$TAG.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:"$KRE" -s 7 -c 4 -o gpurun_out/prof_$TAG -f $CMD > gpurun_out/ncu_f_$TAG.log 2>&1; echo "ncu rc=$?"
ncu -i gpurun_out/prof_$TAG.ncu-rep --page raw --csv > gpurun_out/prof_${TAG}_raw.csv 2>/dev/null
python tools/ncu_summary.py gpurun_out/prof_${TAG}_raw.csv gpurun_out/prof_${TAG}_summary.csv "ncu --set full, $WL, $KRE, tag $TAG" 2>&1; cut -c1-160 gpurun_out/prof_${TAG}_summary.csv
