#!/bin/bash
export PYTHONUNBUFFERED=1
mkdir -p gpurun_out
L=gpurun_out/r2_18.log
: > $L
for s in 8388608 4194304 2097152 1048576 524288; do
 for g in 1 2; do
  echo "=== config A slots $s groups $g" >> $L
  RT_WAVE_SLOTS=$s RT_WAVE_GROUPS=$g timeout 200 python tools/quick_bench.py 64 8 default 640 360 >> $L 2>&1
 done
done
for s in 8388608 4194304 2097152; do
  echo "=== house 1080p 4spp slots $s" >> $L
  RT_WAVE_SLOTS=$s timeout 200 python tools/quick_bench.py 4 8 house >> $L 2>&1
  echo "=== house 1080p 1spp slots $s" >> $L
  RT_WAVE_SLOTS=$s timeout 200 python tools/quick_bench.py 1 8 house >> $L 2>&1
done
cat $L
