#!/bin/bash
export PYTHONUNBUFFERED=1
mkdir -p gpurun_out
L=gpurun_out/r2_32.log
: > $L
timeout 300 python -m pytest tests/test_cli.py -x -q -m gpu >> $L 2>&1
echo "--- pytest exit $?" >> $L
for v in default shade6 shade7 finish6 prepass5 prepass3; do
  lib=build/variants/librt_backend_$v.so
  [ $v = default ] && lib=rsoderh-raytracing_b200/librt_backend.so
  RT_BACKEND_LIB=$PWD/$lib timeout 200 python tools/quick_bench.py 256 2 house >> $L 2>&1
done
cat $L | tail -10
