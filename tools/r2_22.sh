#!/bin/bash
export PYTHONUNBUFFERED=1
mkdir -p gpurun_out
L=gpurun_out/r2_22.log
: > $L
timeout 900 python -m pytest tests/test_gpu_configs.py tests/test_gpu_parity.py tests/test_cli.py -x -q -m gpu >> $L 2>&1
echo "--- pytest exit $?" >> $L
timeout 900 python bench.py --config C --steps 20 --warmup 5 > gpurun_out/bench_r2_v7_C.json 2>> $L
python - >> $L <<'PY'
import json
d=json.loads(open('gpurun_out/bench_r2_v7_C.json').read().strip().splitlines()[-1])
print('C value %.0f e2e %.0f ms %.3f e2e ms %.3f'%(d['value'], d['e2e']['value'], d['ms_per_step'], d['e2e']['ms_per_step']), d['gpu_launches'])
PY
QB_FIRST_HIT=1 timeout 100 python tools/quick_bench.py 1 20 suzanne 3840 2160 >> $L 2>&1
tail -8 $L
