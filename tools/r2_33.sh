#!/bin/bash
export PYTHONUNBUFFERED=1
mkdir -p gpurun_out
L=gpurun_out/r2_33.log
: > $L
run() { echo "=== $1 blocks $2" >> $L; RT_WAVE_BLOCKS2=$2 RT_BACKEND_LIB=$PWD/build/variants/librt_backend_$1.so timeout 200 python tools/quick_bench.py 256 2 house >> $L 2>&1; }
run shade9 9,8,4,5
run shade10 10,8,4,5
run finish10 8,10,4,5
run both10 10,10,4,5
echo "=== default blocks 8,8,4,5" >> $L
timeout 200 python tools/quick_bench.py 256 2 house >> $L 2>&1
cat $L
