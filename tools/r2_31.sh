#!/bin/bash
export PYTHONUNBUFFERED=1
mkdir -p gpurun_out
L=gpurun_out/r2_31.log
: > $L
timeout 900 python -m pytest tests -x -q -m gpu >> $L 2>&1
echo "--- pytest exit $?" >> $L
for n4 in 1 0; do
  echo "=== RT_WALK_N4=$n4" >> $L
  RT_WALK_N4=$n4 timeout 300 python tools/quick_bench.py 16 3 synthetic:99 >> $L 2>&1
  RT_WALK_N4=$n4 timeout 300 python tools/quick_bench.py 32 3 suzanne >> $L 2>&1
  RT_WALK_N4=$n4 timeout 300 python tools/quick_bench.py 32 3 synthetic:20 >> $L 2>&1
done
timeout 200 python tools/quick_bench.py 256 2 house >> $L 2>&1
tail -14 $L
