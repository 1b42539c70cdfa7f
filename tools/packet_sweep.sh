#!/bin/bash
# usage: tools/packet_sweep.sh <tag> — A/B of the packet kernels: RTB_PACKET_CLOSEST / RTB_PACKET_SHADOW = max depth served by
# k_primary / k_packet (-1 = none: k_raygen + per-lane k_traverse, the round-1 schedule).  One line per setting and workload.
TAG=${1:-r2}
OUT=gpurun_out/sweep_packet_$TAG.log
: > $OUT
for PAIR in "-1 -1" "0 -1" "0 0" "1 0" "1 1" "2 2" "16 16"; do
  set -- $PAIR
  for WL in ${SWEEP_WORKLOADS:-c4 c3 c2}; do
    STEPS=${SWEEP_STEPS:-60}; [ "$WL" = c3 ] && STEPS=20; [ "$WL" = c5 ] && STEPS=3
    RTB_PACKET_CLOSEST=$1 RTB_PACKET_SHADOW=$2 python bench.py --workload $WL --steps $STEPS --warmup 3 --no-cpu-baseline 2>gpurun_out/sweep_err.log | tail -1 \
      | python tools/oneline.py "[closest<=$1 shadow<=$2] $WL" >> $OUT 2>&1 || tail -3 gpurun_out/sweep_err.log >> $OUT
  done
done
cat $OUT
