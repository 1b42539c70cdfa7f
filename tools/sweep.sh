#!/bin/bash
# usage: tools/sweep.sh "<EXTRA flags 1>" "<EXTRA flags 2>" ...   — rebuilds the library with each flag set and benches c2 c3 c4
for X in "$@"; do
  touch cosig-raytracing_b200/csrc/*.cu
  make -C cosig-raytracing_b200/csrc -j8 EXTRA="$X" > /dev/null 2>&1 || { echo "build failed: $X"; continue; }
  for WL in ${SWEEP_WORKLOADS:-c2 c3 c4}; do
    python bench.py --workload $WL --steps 50 --warmup 3 --no-cpu-baseline 2>/dev/null | tail -1 | python tools/oneline.py "[$X] $WL"
  done
done
