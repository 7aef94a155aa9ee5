#!/bin/bash
export PYTHONUNBUFFERED=1
mkdir -p gpurun_out
L=gpurun_out/r2_24.log
: > $L
for spp in 1 2 4 8 16; do
  echo "=== spp $spp (adaptive default)" >> $L
  timeout 100 python tools/quick_bench.py $spp 10 house >> $L 2>&1
done
cat $L
