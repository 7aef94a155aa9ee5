#!/bin/bash
# strong scaling of the headline frame on one 8-GPU box (+ config E, + the weak line at N=8)
export PYTHONUNBUFFERED=1 BENCH_DEBUG=1
mkdir -p gpurun_out
L=gpurun_out/r2_multi8.log
: > $L
N=$(nvidia-smi -L | wc -l)
echo "GPUs: $N" >> $L
port=29600
line() {
python - $1 >> $L <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(d['n_gpus'], d['scaling'], d['config']['config'], 'value %.0f e2e %.0f ms %.2f'%(d['value'], d['e2e']['value'], d['ms_per_step']), 'render_max', d['config'].get('render_ms_rank_max'), 'reduce', d['config'].get('reduce_ms'), d.get('reduce_check'), d['config'].get('step_ms_rank0'))
except Exception as e:
    print('unparsed', e)
PY
}
for n in 8 4 2 1; do
  [ $n -le $N ] || continue
  port=$((port+1))
  out=gpurun_out/bench_r2_v5_B_strong_g$n.json
  if [ $n -eq 1 ]; then
    timeout 600 python bench.py --gpus 1 --steps 10 --warmup 3 --no-cpu-baseline > $out 2>> $L
  else
    timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $port \
      bench.py --gpus $n --steps 10 --warmup 3 --no-cpu-baseline > $out 2>> $L
  fi
  echo "--- strong g$n exit $?" >> $L
  line $out
done
if [ $N -ge 8 ]; then
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29650 \
     bench.py --gpus 8 --config E --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_r2_v5_E_strong_g8.json 2>> $L
  echo "--- E g8 exit $?" >> $L
  line gpurun_out/bench_r2_v5_E_strong_g8.json
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29660 \
     bench.py --gpus 8 --scaling weak --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_r2_v5_B_weak_g8.json 2>> $L
  echo "--- weak g8 exit $?" >> $L
  line gpurun_out/bench_r2_v5_B_weak_g8.json
  timeout 300 python -m pytest tests/test_gpu_configs.py -x -q -k "group or cli" >> $L 2>&1
fi
grep -v "^\*\*\*\|OMP_NUM\|^$\|NCCL version" $L | cut -c1-600 | tail -40
