#!/bin/bash
# r2 experiment set 3: which kernel should be the default, and what the footprints cost
export PYTHONUNBUFFERED=1
mkdir -p gpurun_out
L=gpurun_out/r2_third.log
: > $L
run() { echo "=== $*" >> $L; timeout 600 "$@" >> $L 2>&1; echo "--- exit $?" >> $L; }
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv >> $L
run python -m pytest tests -x -q -m gpu
RT_LOCAL_FLAT=0 run python tools/quick_bench.py 32 2 house
RT_LOCAL_FLAT=0 RT_LOCAL_SLOTS=2048 run python tools/quick_bench.py 32 2 house
RT_LOCAL_FLAT=0 QB_ENV=128x64 run python tools/quick_bench.py 32 2 house
QB_ENV=128x64 run python tools/quick_bench.py 32 2 house
RT_KERNEL=wave QB_ENV=128x64 run python tools/quick_bench.py 32 2 house
RT_LOCAL_FLAT=0 run python tools/quick_bench.py 256 2 house
RT_KERNEL=wave run python tools/quick_bench.py 256 2 house
RT_KERNEL=wave RT_WAVE_STAGING_MB=8192 run python tools/quick_bench.py 256 2 house
RT_KERNEL=wave RT_WAVE_STAGING_MB=1024 run python tools/quick_bench.py 256 2 house
RT_KERNEL=wave RT_WAVE_SLOTS=4194304 run python tools/quick_bench.py 256 2 house
RT_KERNEL=wave RT_WAVE_SLOTS=2097152 run python tools/quick_bench.py 256 2 house
RT_KERNEL=wave RT_WAVE_SLOTS=1048576 run python tools/quick_bench.py 256 2 house
RT_KERNEL=wave RT_WAVE_SLOTS=524288 run python tools/quick_bench.py 32 2 house
RT_KERNEL=wave RT_WAVE_SLOTS=2097152 run python tools/quick_bench.py 32 2 house
RT_KERNEL=wave RT_WAVE_SLOTS=4194304 run python tools/quick_bench.py 32 2 house
RT_LOCAL_FLAT=0 run python tools/quick_bench.py 16 2 synthetic:99
RT_KERNEL=wave run python tools/quick_bench.py 16 2 synthetic:99
tail -5 $L
