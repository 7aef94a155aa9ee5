#!/bin/bash
# full ncu captures of the device BVH build's main kernels, the first-hit kernel (config C) and the deep walk (config D)
export PYTHONUNBUFFERED=1
mkdir -p gpurun_out
L=gpurun_out/r2_28.log
: > $L
cap() { # name regex skip cmd...
  local name=$1 regex=$2 skip=$3; shift 3
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$regex --launch-skip $skip -c 1 \
      -o gpurun_out/prof_r2_v7_$name -f "$@" > gpurun_out/ncu_r2_v7_$name.log 2>&1
  echo "--- $name exit $? $(tail -1 gpurun_out/ncu_r2_v7_$name.log)" >> $L
}
export QB_NO_WARMUP=1
cap bvh_bucket "bucket_kernel" 12 python tools/bvh_build_probe.py 99
cap bvh_move "move_kernel" 12 python tools/bvh_build_probe.py 99
cap bvh_split "split_kernel" 20 python tools/bvh_build_probe.py 99
unset QB_NO_WARMUP
QB_FIRST_HIT=1 cap first_hit_C "first_hit_fast" 5 python tools/quick_bench.py 1 2 suzanne 3840 2160
RT_WAVE_GRAPH=0 cap walk_D "wave_walk" 20 python tools/quick_bench.py 16 1 synthetic:99
cat $L
ls -la gpurun_out/*.ncu-rep | tail -8
