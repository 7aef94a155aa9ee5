#!/bin/bash
# the wave loop as a CUDA graph (WHILE node): correctness first, then timings against the host-paced loop
export PYTHONUNBUFFERED=1
mkdir -p gpurun_out
L=gpurun_out/r2_sixth.log
: > $L
run() { echo "=== $*" >> $L; timeout ${T:-300} "$@" >> $L 2>&1; echo "--- exit $?" >> $L; }
T=60 run python tools/quick_bench.py 2 1 house 256 144
T=60 RT_WAVE_GRAPH=0 run python tools/quick_bench.py 2 1 house 256 144
T=120 run python tools/quick_bench.py 32 3 house
T=120 RT_WAVE_GRAPH=0 run python tools/quick_bench.py 32 3 house
for b in 2 8 16; do T=120 RT_WAVE_BODY=$b run python tools/quick_bench.py 32 3 house; done
T=200 run python tools/quick_bench.py 256 2 house
T=200 RT_WAVE_GRAPH=0 run python tools/quick_bench.py 256 2 house
T=200 RT_WAVE_BODY=8 run python tools/quick_bench.py 256 2 house
T=120 run python tools/quick_bench.py 64 3 default 640 360
T=120 RT_WAVE_GRAPH=0 run python tools/quick_bench.py 64 3 default 640 360
T=120 run python tools/quick_bench.py 1 5 house
T=120 RT_WAVE_GRAPH=0 run python tools/quick_bench.py 1 5 house
T=300 run python tools/quick_bench.py 16 2 synthetic:99
T=900 run python -m pytest tests -x -q -m gpu
for c in A B; do
  timeout 900 python bench.py --config $c --steps 5 --warmup 3 > gpurun_out/bench_r2_v2_$c.json 2>> $L
  echo "--- bench $c exit $?" >> $L
  tail -c 300 gpurun_out/bench_r2_v2_$c.json >> $L
done
tail -5 $L
