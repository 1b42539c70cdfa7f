#!/bin/bash
# usage: tools/wide_sweep.sh <tag> — A/B of the LBVH node format: RTB_WIDE=1 (8-wide quantised records, default) against 0 (binary two-box records)
TAG=${1:-r2}
OUT=gpurun_out/sweep_wide_$TAG.log
: > $OUT
for W in 1 0; do
  for WL in ${SWEEP_WORKLOADS:-c4 c3 c2}; do
    STEPS=${SWEEP_STEPS:-60}; [ "$WL" = c3 ] && STEPS=20; [ "$WL" = c5 ] && STEPS=3
    RTB_WIDE=$W python bench.py --workload $WL --steps $STEPS --warmup 3 --no-cpu-baseline 2>gpurun_out/sweep_err.log | tail -1 \
      | python tools/oneline.py "[RTB_WIDE=$W] $WL" >> $OUT 2>&1 || tail -3 gpurun_out/sweep_err.log >> $OUT
  done
done
cat $OUT
