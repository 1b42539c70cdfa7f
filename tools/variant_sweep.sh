#!/bin/bash
# usage (on the GPU box): tools/variant_sweep.sh tag1 tag2 ...   — benches $SWEEP_WORKLOADS (default c4) with each prebuilt
# tools/variants/librtb200_<tag>.so in place of the default library, then restores it.
cp cosig-raytracing_b200/librtb200.so /tmp/librtb200_default.so
for TAG in "$@"; do
  if [ "$TAG" = default ]; then cp /tmp/librtb200_default.so cosig-raytracing_b200/librtb200.so; else cp tools/variants/librtb200_$TAG.so cosig-raytracing_b200/librtb200.so; fi
  for WL in ${SWEEP_WORKLOADS:-c4}; do
    python bench.py --workload $WL --steps ${SWEEP_STEPS:-100} --warmup 3 --no-cpu-baseline 2>/dev/null | tail -1 | python tools/oneline.py "[$TAG] $WL"
  done
done
cp /tmp/librtb200_default.so cosig-raytracing_b200/librtb200.so
