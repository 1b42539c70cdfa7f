#!/bin/bash
# usage: tools/build_variants.sh tag1 "<EXTRA flags 1>" tag2 "<EXTRA flags 2>" ...
# Builds librtb200.so once per flag set HERE (nvcc cross-compiles) into tools/variants/librtb200_<tag>.so; the files travel to
# the GPU box with the snapshot (git-ignored, not gpurun-ignored), where tools/variant_sweep.sh benches them.  The default
# library is rebuilt at the end.
set -e
mkdir -p tools/variants
while [ $# -ge 2 ]; do
  TAG=$1; X=$2; shift 2
  touch cosig-raytracing_b200/csrc/*.cu cosig-raytracing_b200/csrc/*.cpp
  make -C cosig-raytracing_b200/csrc -j8 EXTRA="$X" > /dev/null
  cp cosig-raytracing_b200/librtb200.so tools/variants/librtb200_$TAG.so
  echo "$TAG: $X" >> tools/variants/FLAGS.txt
done
touch cosig-raytracing_b200/csrc/*.cu cosig-raytracing_b200/csrc/*.cpp
make -C cosig-raytracing_b200/csrc -j8 > /dev/null
