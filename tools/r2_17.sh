#!/bin/bash
export PYTHONUNBUFFERED=1
mkdir -p gpurun_out
L=gpurun_out/r2_17.log
: > $L
run() { echo "=== $*" >> $L; timeout ${T:-600} "$@" >> $L 2>&1; echo "--- exit $?" >> $L; }
T=900 run python -m pytest tests -x -q -m gpu
for prio in 0,0,0,0 1,1,0,0 0,0,1,1 1,0,0,0 0,1,0,0 0,0,0,1 0,0,1,0 2,1,0,0 1,1,0,1; do
  echo "=== RT_WAVE_PRIO=$prio" >> $L
  RT_WAVE_PRIO=$prio timeout 200 python tools/quick_bench.py 256 3 house >> $L 2>&1
  RT_WAVE_PRIO=$prio timeout 200 python tools/quick_bench.py 32 5 house >> $L 2>&1
done
tail -40 $L
