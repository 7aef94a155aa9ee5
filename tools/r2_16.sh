#!/bin/bash
export PYTHONUNBUFFERED=1
mkdir -p gpurun_out
L=gpurun_out/r2_16.log
: > $L
run() { echo "=== $*" >> $L; timeout ${T:-600} "$@" >> $L 2>&1; echo "--- exit $?" >> $L; }
T=900 run python -m pytest tests -x -q -m gpu
T=300 run python tools/bvh_build_probe.py 99
timeout 900 python bench.py --config D --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_r2_v6_D.json 2>> $L
echo "--- bench D exit $?" >> $L
python - >> $L <<'PY'
import json
d=json.loads(open('gpurun_out/bench_r2_v6_D.json').read().strip().splitlines()[-1])
print('D value %.0f e2e %.0f ms %.2f'%(d['value'], d['e2e']['value'], d['ms_per_step']), d['config'].get('bvh_build'), d['config'].get('scene_build_s'), d['e2e'].get('note'), d['roofline'].get('traffic'))
PY
tail -25 $L
