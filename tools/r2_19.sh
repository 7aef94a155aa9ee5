#!/bin/bash
export PYTHONUNBUFFERED=1
mkdir -p gpurun_out
L=gpurun_out/r2_19.log
: > $L
run() { echo "=== $*" >> $L; timeout ${T:-600} "$@" >> $L 2>&1; echo "--- exit $?" >> $L; }
T=900 run python -m pytest tests -x -q -m gpu
for mb in 8192 4096 2048 1024 512; do
  echo "=== staging $mb MiB" >> $L
  RT_WAVE_STAGING_MB=$mb timeout 200 python tools/quick_bench.py 256 3 house >> $L 2>&1
done
echo "=== staging 2048 MiB, 4 Mi slots" >> $L
RT_WAVE_STAGING_MB=2048 RT_WAVE_SLOTS=4194304 timeout 200 python tools/quick_bench.py 256 3 house >> $L 2>&1
echo "=== config A, 4 spp 1080p, 32 spp with the adaptive slot count" >> $L
timeout 200 python tools/quick_bench.py 64 8 default 640 360 >> $L 2>&1
timeout 200 python tools/quick_bench.py 4 8 house >> $L 2>&1
timeout 200 python tools/quick_bench.py 32 5 house >> $L 2>&1
timeout 900 python bench.py --config A --steps 10 --warmup 3 > gpurun_out/bench_r2_v6_A.json 2>> $L
tail -c 600 gpurun_out/bench_r2_v6_A.json >> $L
tail -32 $L
