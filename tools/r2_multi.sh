#!/bin/bash
# multi-GPU: group tests + strong-scaling bench lines at N = 1, 2, 4, 8 (as many as the box has)
export PYTHONUNBUFFERED=1
mkdir -p gpurun_out
TAG=${TAG:-v1}
L=gpurun_out/r2_multi_$TAG.log
: > $L
N=$(nvidia-smi -L | wc -l)
echo "GPUs: $N" >> $L
timeout 600 python -m pytest tests/test_gpu_configs.py -x -q -k "group or cli" >> $L 2>&1
echo "--- exit $?" >> $L
port=29500
for n in 1 2 4 8; do
  [ $n -le $N ] || continue
  for mode in strong weak; do
    [ $mode = weak ] && [ $n -ne ${WEAK_AT:-8} ] && continue
    port=$((port+1))
    out=gpurun_out/bench_r2_${TAG}_B_${mode}_g$n.json
    if [ $n -eq 1 ]; then
      timeout 900 python bench.py --gpus 1 --steps 5 --warmup 3 > $out 2>> $L
    else
      timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $port \
        bench.py --gpus $n --steps 5 --warmup 3 --scaling $mode --no-cpu-baseline > $out 2>> $L
    fi
    echo "--- $mode g$n exit $?" >> $L
    python - $out >> $L <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(d['n_gpus'], d['scaling'], 'value %.0f e2e %.0f ms %.2f'%(d['value'], d['e2e']['value'], d['ms_per_step']), d['config'].get('render_ms_rank_max'), d['config'].get('reduce_ms'), d.get('reduce_check'))
except Exception as e:
    print('unparsed', e)
PY
  done
done
if [ $N -ge 8 ] && [ -n "$WITH_E" ]; then
  timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29600 \
     bench.py --gpus 8 --config E --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/bench_r2_${TAG}_E_strong_g8.json 2>> $L
  echo "--- E g8 exit $?" >> $L
fi
tail -20 $L
