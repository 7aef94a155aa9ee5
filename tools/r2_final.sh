#!/bin/bash
# what the driver runs at round end, on the final tree
export PYTHONUNBUFFERED=1
mkdir -p gpurun_out
L=gpurun_out/r2_final.log
: > $L
timeout 900 python -m pytest tests -x -q -m gpu >> $L 2>&1
echo "--- pytest exit $?" >> $L
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" >> $L 2>&1
echo "--- smoke exit $?" >> $L
timeout 900 python bench.py --impl reference --gpus 1 --steps 5 --warmup 2 > gpurun_out/bench_r2_final_reference.json 2>> $L
echo "--- reference exit $?" >> $L
timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/bench_r2_final_B.json 2>> $L
echo "--- bench exit $?" >> $L
python - >> $L <<'PY'
import json
d=json.loads(open('gpurun_out/bench_r2_final_B.json').read().strip().splitlines()[-1])
r=json.loads(open('gpurun_out/bench_r2_final_reference.json').read().strip().splitlines()[-1])
print('B value %.0f e2e %.0f ms %.2f'%(d['value'], d['e2e']['value'], d['ms_per_step']), d['clocks'], 'launches', d['gpu_launches'])
print('reference %.1f'%r['value'], 'e2e ratio %.1f'%(d['e2e']['value']/r['value']))
print(d['roofline'])
PY
tail -14 $L
