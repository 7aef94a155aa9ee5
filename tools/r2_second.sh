#!/bin/bash
export PYTHONUNBUFFERED=1
mkdir -p gpurun_out
L=gpurun_out/r2_second.log
: > $L
run() { echo "=== $*" >> $L; timeout 600 "$@" >> $L 2>&1; echo "--- exit $?" >> $L; }
V=build/variants
RT_LOCAL_PROFILE=1 run python tools/quick_bench.py 32 1 house
RT_LOCAL_PROFILE=1 RT_LOCAL_FLAT=0 run python tools/quick_bench.py 32 1 house
RT_LOCAL_PROFILE=1 RT_LOCAL_SLOTS=2048 run python tools/quick_bench.py 32 1 house
for cfg in "t256b4 256" "t256b4 512" "t256b4 1024" "t256b3 512" "t256b3 1024" "t128b8 128" "t128b8 256" "t128b8 512" "t1024b1 2048" "t1024b1 4096"; do
  set -- $cfg
  RT_BACKEND_LIB=$V/librt_backend_$1.so RT_LOCAL_SLOTS=$2 RT_LOCAL_PROFILE=1 run python tools/quick_bench.py 32 2 house
  RT_BACKEND_LIB=$V/librt_backend_$1.so RT_LOCAL_SLOTS=$2 RT_LOCAL_FLAT=0 run python tools/quick_bench.py 32 2 house
done
# one full capture of the persistent kernel (8 spp; launch 0 is quick_bench's 2-spp warm-up)
timeout 900 ncu --set full --clock-control none --import-source on -k regex:local_wave --launch-skip 1 --launch-count 1 \
   -o gpurun_out/r2_local_flat_v1 -f python tools/quick_bench.py 8 1 house >> $L 2>&1
tail -3 $L
