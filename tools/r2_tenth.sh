#!/bin/bash
# device BVH build vs host build; bench line with the NVML sampler and graph parameter updates
export PYTHONUNBUFFERED=1
mkdir -p gpurun_out
L=gpurun_out/r2_tenth.log
: > $L
run() { echo "=== $*" >> $L; timeout ${T:-300} "$@" >> $L 2>&1; echo "--- exit $?" >> $L; }
T=600 run python -m pytest tests/test_gpu_bvh_build.py -x -q -s
BENCH_DEBUG=1 RT_WAVE_GRAPH_DEBUG=1 timeout 900 python bench.py --steps 8 --warmup 3 > gpurun_out/bench_r2_v4_B_g1.json 2>> $L
echo "--- bench exit $?" >> $L
python - >> $L <<'PY'
import json
d=json.loads(open('gpurun_out/bench_r2_v4_B_g1.json').read().strip().splitlines()[-1])
print('value %.0f e2e %.0f ms %.2f e2e_ms %.2f'%(d['value'], d['e2e']['value'], d['ms_per_step'], d['e2e']['ms_per_step']), d['clocks'], d['config']['step_ms_rank0'])
PY
T=300 run python -m pytest tests/test_gpu_parity.py tests/test_gpu_wavefront.py -x -q
tail -40 $L
