#!/bin/bash
export PYTHONUNBUFFERED=1
mkdir -p gpurun_out
L=gpurun_out/r2_25.log
: > $L
for spp in 1 2 4 8; do
  echo "=== spp $spp (paths / 2)" >> $L
  timeout 100 python tools/quick_bench.py $spp 10 house >> $L 2>&1
done
echo "=== config A" >> $L
timeout 100 python tools/quick_bench.py 64 10 default 640 360 >> $L 2>&1
echo "=== 1 spp, slots forced" >> $L
for s in 2097152 1572864 1048576 524288; do RT_WAVE_SLOTS=$s timeout 100 python tools/quick_bench.py 1 10 house >> $L 2>&1; done
cat $L
