#!/bin/bash
# first contact of the per-warp engines (small, under timeouts), then tests at BASELINE sizes and one
# bench line per config on one GPU
export PYTHONUNBUFFERED=1
mkdir -p gpurun_out
L=gpurun_out/r2_fourth.log
: > $L
run() { echo "=== $*" >> $L; timeout ${T:-900} "$@" >> $L 2>&1; echo "--- exit $?" >> $L; }
T=60 RT_KERNEL=local run python tools/quick_bench.py 1 1 house 64 32
T=60 RT_KERNEL=local run python tools/quick_bench.py 2 1 house 256 144
T=60 RT_KERNEL=wave run python tools/quick_bench.py 2 1 house 256 144
T=120 RT_KERNEL=local RT_LOCAL_PROFILE=1 run python tools/quick_bench.py 32 2 house
T=120 RT_KERNEL=local RT_LOCAL_FLAT=1 RT_LOCAL_PROFILE=1 run python tools/quick_bench.py 32 2 house
for s in 64 96 192 256; do T=120 RT_KERNEL=local RT_LOCAL_SLOTS=$s run python tools/quick_bench.py 32 2 house; done
for w in 2 8 16; do T=120 RT_KERNEL=local RT_LOCAL_WINDOW=$w run python tools/quick_bench.py 32 2 house; done
T=300 RT_KERNEL=local run python tools/quick_bench.py 256 2 house
T=300 RT_KERNEL=local run python tools/quick_bench.py 16 2 synthetic:99
T=600 run python -m pytest tests/test_gpu_wavefront.py -x -q
T=900 run python -m pytest tests -x -q -m gpu --durations=8
for c in A B C D; do
  echo "=== bench config $c" >> $L
  timeout 900 python bench.py --config $c --steps 3 --warmup 3 > gpurun_out/bench_r2_v1_$c.json 2>> $L
  echo "--- exit $?" >> $L
  tail -c 600 gpurun_out/bench_r2_v1_$c.json >> $L
done
timeout 600 python bench.py --impl reference --config B --steps 2 --warmup 1 > gpurun_out/bench_r2_v1_ref_B.json 2>> $L
tail -5 $L
