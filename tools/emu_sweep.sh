cp cosig-raytracing_b200/librtb200.so /tmp/librtb200_default.so
for TAG in ${EMU_TAGS:-default}; do
  if [ $TAG != default ]; then cp tools/variants/librtb200_$TAG.so cosig-raytracing_b200/librtb200.so; else cp /tmp/librtb200_default.so cosig-raytracing_b200/librtb200.so; fi
  for EW in 8 1; do
    python bench.py --emulate-world $EW --steps 60 --warmup 3 --no-cpu-baseline 2>/dev/null | tail -1 | python tools/oneline.py "[$TAG emu$EW]"
  done
done
cp /tmp/librtb200_default.so cosig-raytracing_b200/librtb200.so
