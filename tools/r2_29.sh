#!/bin/bash
export PYTHONUNBUFFERED=1
mkdir -p gpurun_out
L=gpurun_out/r2_29.log
: > $L
for v in default walk4 walk6 walk8; do
  lib=build/variants/librt_backend_$v.so
  [ $v = default ] && lib=rsoderh-raytracing_b200/librt_backend.so
  echo "=== $v" >> $L
  RT_BACKEND_LIB=$PWD/$lib timeout 300 python tools/quick_bench.py 16 3 synthetic:99 >> $L 2>&1
  RT_BACKEND_LIB=$PWD/$lib timeout 200 python tools/quick_bench.py 256 2 house >> $L 2>&1
done
cat $L
