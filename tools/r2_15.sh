#!/bin/bash
export PYTHONUNBUFFERED=1
mkdir -p gpurun_out
L=gpurun_out/r2_15.log
: > $L
run() { echo "=== $*" >> $L; timeout ${T:-600} "$@" >> $L 2>&1; echo "--- exit $?" >> $L; }
T=300 run python -m pytest tests/test_cli.py -x -q -m gpu
RT_BVH_FUZZ_SEEDS=200 T=600 run python -m pytest tests/test_gpu_bvh_build.py -x -q -k fuzz
RT_FUZZ_SEEDS=300 T=900 run python -m pytest tests/test_gpu_wavefront.py -x -q -k fuzz
# config D: every launch of one 16-spp step
RT_WAVE_GRAPH=0 QB_NO_WARMUP=1 timeout 900 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,smsp__thread_inst_executed.sum,dram__bytes_read.sum,dram__bytes_write.sum \
  --clock-control none --csv --log-file gpurun_out/ncu_r2_v5_step_D.csv python tools/quick_bench.py 16 1 synthetic:99 > gpurun_out/ncu_r2_v5_step_D.log 2>&1
echo "--- ncu step D exit $? $(wc -l < gpurun_out/ncu_r2_v5_step_D.csv) lines" >> $L
# the device BVH build's kernels
T=300 run python tools/bvh_build_probe.py 99
QB_NO_WARMUP=1 timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__inst_executed.sum,smsp__thread_inst_executed.sum \
  --clock-control none --csv --log-file gpurun_out/ncu_r2_v5_bvh_build.csv python tools/bvh_build_probe.py 99 > gpurun_out/ncu_r2_v5_bvh_build.log 2>&1
echo "--- ncu bvh exit $? $(wc -l < gpurun_out/ncu_r2_v5_bvh_build.csv) lines" >> $L
tail -30 $L
