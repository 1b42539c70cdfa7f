#!/bin/bash
# usage: tools/env_sweep.sh VAR v1 v2 ...   — benches c2 c4 (or $SWEEP_WORKLOADS) with VAR set to each value
VAR=$1; shift
for V in "$@"; do
  for WL in ${SWEEP_WORKLOADS:-c2 c4}; do
    STEPS=${SWEEP_STEPS:-100}; [ "$WL" = c5 ] && STEPS=3
    env $VAR=$V python bench.py --workload $WL --steps $STEPS --warmup 3 --no-cpu-baseline 2>/dev/null | tail -1 | python tools/oneline.py "[$VAR=$V] $WL"
  done
done
