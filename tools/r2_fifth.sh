#!/bin/bash
# why are the per-warp engines slow? one full ncu capture + variants + the deep-tree crash bisect
export PYTHONUNBUFFERED=1
mkdir -p gpurun_out
L=gpurun_out/r2_fifth.log
: > $L
run() { echo "=== $*" >> $L; timeout ${T:-300} "$@" >> $L 2>&1; echo "--- exit $?" >> $L; }
V=build/variants
export RT_KERNEL=local
for v in busy1 busy8 busy24 t256b4; do RT_BACKEND_LIB=$V/librt_backend_$v.so run python tools/quick_bench.py 32 2 house; done
RT_LOCAL_SLOTS=32 run python tools/quick_bench.py 32 2 house
T=60 run python tools/quick_bench.py 2 1 synthetic:3 256 144
T=60 RT_BACKEND_LIB=$V/librt_backend_busy1.so run python tools/quick_bench.py 2 1 synthetic:3 256 144
T=120 RT_BACKEND_LIB=$V/librt_backend_busy1.so run python tools/quick_bench.py 2 1 synthetic:99 256 144
T=120 run python tools/quick_bench.py 2 1 synthetic:20 256 144
timeout 900 ncu --set full --clock-control none --import-source on -k regex:local_wave --launch-skip 1 --launch-count 1 \
   -o gpurun_out/r2_warp_v1 -f python tools/quick_bench.py 8 1 house >> $L 2>&1
tail -3 $L
