#!/bin/bash
export PYTHONUNBUFFERED=1
mkdir -p gpurun_out
L=gpurun_out/r2_34.log
: > $L
for b in 8,8,4,5 8,8,4,4 8,8,4,6 8,8,5,5 8,8,3,5 6,8,4,5 8,6,4,5 7,7,4,5 8,8,2,5 8,8,4,3; do
  echo "=== RT_WAVE_BLOCKS2=$b" >> $L
  RT_WAVE_BLOCKS2=$b timeout 200 python tools/quick_bench.py 256 2 house >> $L 2>&1
done
for w in 2 4 6 8; do
  echo "=== RT_WAVE_BODY=$w" >> $L
  RT_WAVE_BODY=$w timeout 200 python tools/quick_bench.py 32 5 house >> $L 2>&1
done
grep -v "^$" $L
