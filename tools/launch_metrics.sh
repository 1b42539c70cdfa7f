#!/bin/bash
# usage: tools/launch_metrics.sh <tag> [bench args]  — per-launch metrics of one warm frame
TAG=$1; shift
CMD="python bench.py --steps 2 --warmup 1 --no-cpu-baseline $*"
$CMD > gpurun_out/plain_$TAG.log 2>&1 && ncu --metrics gpu__time_duration.sum,smsp__thread_inst_executed_per_inst_executed.ratio,smsp__inst_executed.sum,sm__warps_active.avg.pct_of_peak_sustained_active,smsp__issue_active.avg.pct_of_peak_sustained_active --clock-control none -k regex:"k_traverse|k_shade" -s 13 -c 13 --csv --log-file gpurun_out/launches_$TAG.csv $CMD > gpurun_out/ncu_$TAG.log 2>&1
echo "ncu rc=$?"
tail -1 gpurun_out/plain_$TAG.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('value',d['value'],'ms',d['ms_per_step'],'e2e',d['e2e']['value'],d['kernel_ms_per_frame'])"
