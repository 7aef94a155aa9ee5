#!/bin/bash
# final tree on one 8-GPU box: strong-scaling line at N = 8 (and 4), config E, group tests
export PYTHONUNBUFFERED=1
mkdir -p gpurun_out
L=gpurun_out/r2_multi8b.log
: > $L
line() {
python - $1 >> $L <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(d['n_gpus'], d['scaling'], d['config']['config'], 'value %.0f e2e %.0f ms %.2f e2e_ms %.2f'%(d['value'], d['e2e']['value'], d['ms_per_step'], d['e2e']['ms_per_step']), 'render_max', d['config'].get('render_ms_rank_max'), 'reduce', d['config'].get('reduce_ms'), d.get('reduce_check',{}).get('ok'), d['config'].get('step_ms_rank0'))
except Exception as e:
    print('unparsed', e)
PY
}
port=29700
for n in 8 4; do
  port=$((port+1))
  out=gpurun_out/bench_r2_final_B_strong_g$n.json
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $port \
      bench.py --gpus $n --steps 20 --warmup 5 --no-cpu-baseline > $out 2>> $L
  echo "--- strong g$n exit $?" >> $L
  line $out
done
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29750 \
   bench.py --gpus 8 --config E --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_r2_final_E_strong_g8.json 2>> $L
echo "--- E g8 exit $?" >> $L
line gpurun_out/bench_r2_final_E_strong_g8.json
timeout 300 python -m pytest tests/test_gpu_configs.py -x -q -k "group or cli" >> $L 2>&1
grep -v "^\*\*\*\|OMP_NUM\|^$\|NCCL version" $L | cut -c1-500 | tail -12
