#!/bin/bash
# usage: tools/profile_round.sh <tag>   — bench lines + ncu launch list + ncu full capture of k_traverse for C4 (1 GPU)
TAG=$1
python bench.py --steps 100 --warmup 3 > gpurun_out/bench_c4_$TAG.json 2> gpurun_out/bench_c4_$TAG.err; echo "bench rc=$?"
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref_c4_$TAG.json 2>/dev/null
python bench.py --workload c2 --steps 100 --warmup 3 > gpurun_out/bench_c2_$TAG.json 2>/dev/null
python bench.py --workload c2 --bvh reference --steps 100 --warmup 3 > gpurun_out/bench_c2ref_$TAG.json 2>/dev/null
python bench.py --workload c3 --steps 20 --warmup 3 > gpurun_out/bench_c3_$TAG.json 2>/dev/null
python bench.py --workload c5 --steps 3 --warmup 1 > gpurun_out/bench_c5_$TAG.json 2>/dev/null
CMD="python bench.py --steps 2 --warmup 1 --no-cpu-baseline"
export RTB_LANES=1   # one stream: ncu serialises launches anyway; keeps the launch list in program order
$CMD > gpurun_out/plain_$TAG.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches_$TAG.csv $CMD > gpurun_out/ncu_l_$TAG.log 2>&1; echo "launch list rc=$?"
$CMD > gpurun_out/plain2_$TAG.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:"k_traverse" -s 7 -c 7 -o gpurun_out/prof_$TAG $CMD > gpurun_out/ncu_f_$TAG.log 2>&1; echo "full rc=$?"
$CMD > /dev/null 2>&1 && ncu --set full --clock-control none --import-source on -k regex:"k_shade|k_raygen" -s 14 -c 3 -o gpurun_out/prof_stream_$TAG $CMD > gpurun_out/ncu_s_$TAG.log 2>&1; echo "full (streaming kernels) rc=$?"
