#!/bin/bash
# full GPU suite on the fixed build, then the ncu evidence of round 2 and the other configs' bench lines
export PYTHONUNBUFFERED=1
mkdir -p gpurun_out
L=gpurun_out/r2_14.log
: > $L
timeout 900 python -m pytest tests -x -q -m gpu >> $L 2>&1
echo "--- pytest exit $?" >> $L
# 1. every launch of one 256-spp step: exact instruction and DRAM byte counts per kernel
RT_WAVE_GRAPH=0 QB_NO_WARMUP=1 timeout 900 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,smsp__thread_inst_executed.sum,dram__bytes_read.sum,dram__bytes_write.sum \
  --clock-control none --csv --log-file gpurun_out/ncu_r2_v5_step.csv python tools/quick_bench.py 256 1 > gpurun_out/ncu_r2_v5_step.log 2>&1
echo "--- ncu step exit $? $(wc -l < gpurun_out/ncu_r2_v5_step.csv) lines" >> $L
# 2. steady-state full captures of the four wave kernels
NCU_CMD="env RT_WAVE_GRAPH=0 python tools/quick_bench.py 32 1" timeout 900 bash tools/ncu_capture.sh r2_v5 32 >> $L 2>&1
echo "--- ncu captures exit $?" >> $L
# 3. launch list of the bench command (host-paced loop: the same kernels, one launch each)
RT_WAVE_GRAPH=0 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 2400 --csv \
  --log-file gpurun_out/ncu_r2_v5_launches_default_command.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_r2_v5_launches.log 2>&1
echo "--- ncu launch list exit $?" >> $L
# 4. the other configs
for c in A C D; do
  timeout 900 python bench.py --config $c --steps 5 --warmup 3 > gpurun_out/bench_r2_v5_$c.json 2>> $L
  echo "--- bench $c exit $?" >> $L
done
timeout 900 python bench.py --config E --spp 128 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/bench_r2_v5_E_share_g1.json 2>> $L
echo "--- bench E share exit $?" >> $L
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_r2_v5_reference_B.json 2>> $L
echo "--- bench reference exit $?" >> $L
ls -la gpurun_out >> $L
tail -40 $L
