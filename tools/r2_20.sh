#!/bin/bash
export PYTHONUNBUFFERED=1
mkdir -p gpurun_out
L=gpurun_out/r2_20.log
: > $L
M="gpu__time_duration.sum,smsp__inst_executed.sum,smsp__thread_inst_executed.sum,dram__bytes_read.sum,dram__bytes_write.sum"
RT_WAVE_GRAPH=0 QB_NO_WARMUP=1 timeout 600 ncu --metrics $M --clock-control none --csv --log-file gpurun_out/ncu_r2_v6_step_A.csv python tools/quick_bench.py 64 1 default 640 360 > gpurun_out/ncu_r2_v6_step_A.log 2>&1
echo "--- ncu step A exit $? $(wc -l < gpurun_out/ncu_r2_v6_step_A.csv) lines" >> $L
QB_FIRST_HIT=1 timeout 600 ncu --metrics $M --clock-control none -k regex:first_hit --csv --log-file gpurun_out/ncu_r2_v6_step_C.csv python tools/quick_bench.py 1 1 suzanne 3840 2160 > gpurun_out/ncu_r2_v6_step_C.log 2>&1
echo "--- ncu step C exit $? $(wc -l < gpurun_out/ncu_r2_v6_step_C.csv) lines" >> $L
for c in B A C; do
  timeout 900 python bench.py --config $c --steps 10 --warmup 3 > gpurun_out/bench_r2_v6_$c.json 2>> $L
  echo "--- bench $c exit $?" >> $L
  python - gpurun_out/bench_r2_v6_$c.json >> $L <<'PY'
import json,sys
d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print(d['config']['config'], 'value %.0f e2e %.0f ms %.3f'%(d['value'], d['e2e']['value'], d['ms_per_step']), d['config'].get('environment_build_s'), d['config'].get('state_new_s'), d['roofline'].get('measured'), d['roofline'].get('traffic'))
PY
done
cat $L
