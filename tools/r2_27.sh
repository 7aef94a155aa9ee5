#!/bin/bash
# final 2-GPU check: group / CLI tests and the strong-scaling line
export PYTHONUNBUFFERED=1
mkdir -p gpurun_out
L=gpurun_out/r2_27.log
: > $L
timeout 600 python -m pytest tests/test_gpu_configs.py tests/test_cli.py -x -q -m gpu -k "group or cli or multi" >> $L 2>&1
echo "--- pytest exit $?" >> $L
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 \
  bench.py --gpus 2 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_r2_final_B_g2.json 2>> $L
echo "--- bench g2 exit $?" >> $L
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29534 \
  bench.py --impl reference --gpus 2 --steps 2 --warmup 1 > gpurun_out/bench_r2_final_reference_g2.json 2>> $L
echo "--- reference g2 exit $? lines $(wc -l < gpurun_out/bench_r2_final_reference_g2.json)" >> $L
python - >> $L <<'PY'
import json
d=json.loads(open('gpurun_out/bench_r2_final_B_g2.json').read().strip().splitlines()[-1])
print(d['n_gpus'], d['scaling'], 'value %.0f e2e %.0f ms %.2f'%(d['value'], d['e2e']['value'], d['ms_per_step']), d['reduce_check']['ok'], d['config']['step_ms_rank0'])
PY
grep -v "^\*\*\*\|OMP_NUM\|^$\|NCCL version" $L | tail -12
