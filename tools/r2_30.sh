#!/bin/bash
export PYTHONUNBUFFERED=1
mkdir -p gpurun_out
L=gpurun_out/r2_30.log
: > $L
for g in 1 2 3 4; do
  echo "=== config D groups $g" >> $L
  RT_WAVE_GROUPS=$g timeout 300 python tools/quick_bench.py 16 3 synthetic:99 >> $L 2>&1
done
for s in 4194304 2097152 16777216; do
  echo "=== config D slots $s" >> $L
  RT_WAVE_SLOTS=$s timeout 300 python tools/quick_bench.py 16 3 synthetic:99 >> $L 2>&1
done
echo "=== config D groups 4 slots 16Mi" >> $L
RT_WAVE_GROUPS=4 RT_WAVE_SLOTS=16777216 timeout 300 python tools/quick_bench.py 16 3 synthetic:99 >> $L 2>&1
cat $L
