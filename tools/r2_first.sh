#!/bin/bash
# first GPU contact of the persistent kernel: sanity, sanitizer, tests, first timings
export PYTHONUNBUFFERED=1
mkdir -p gpurun_out
L=gpurun_out/r2_first.log
: > $L
run() { echo "=== $*" >> $L; timeout 600 "$@" >> $L 2>&1; echo "--- exit $?" >> $L; }
run python tools/quick_bench.py 2 1 house 256 144
RT_KERNEL=wave run python tools/quick_bench.py 2 1 house 256 144
run compute-sanitizer --tool memcheck --print-limit 20 python tools/quick_bench.py 1 1 house 128 72
run compute-sanitizer --tool racecheck --print-limit 20 python tools/quick_bench.py 1 1 house 128 72
run python -m pytest tests/test_gpu_wavefront.py -x -q
run python tools/quick_bench.py 32 2 house
RT_KERNEL=wave run python tools/quick_bench.py 32 2 house
RT_LOCAL_FLAT=0 run python tools/quick_bench.py 32 2 house
for s in 512 2048 4096; do RT_LOCAL_SLOTS=$s run python tools/quick_bench.py 32 2 house; done
for s in 2048 4096 16384; do RT_LOCAL_STAGE=$s run python tools/quick_bench.py 32 2 house; done
run python tools/quick_bench.py 256 2 house
run python tools/quick_bench.py 8 2 suzanne
RT_KERNEL=wave run python tools/quick_bench.py 8 2 suzanne
tail -5 $L
