#!/bin/bash
export PYTHONUNBUFFERED=1
mkdir -p gpurun_out
L=gpurun_out/r2_13.log
: > $L
echo "=== graph loop, no cudaMemGetInfo per render" >> $L
timeout 300 python tools/variance_probe.py 32 80 2>&1 | grep -v "Not Active\|^$\|NVSMI\|Timestamp\|Driver Ver\|CUDA Ver\|Attached\|^GPU 0\|Performance State\|Clocks Event\|Sync Boost\|Thermal\|Power Brak\|Sparse" | awk '/<-- slow/ || !/^ *[0-9]+:/' >> $L
BENCH_DEBUG=1 RT_WAVE_GRAPH_DEBUG=1 timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_r2_v5_B_g1.json 2>> $L
echo "--- bench exit $?" >> $L
python - >> $L <<'PY'
import json
d=json.loads(open('gpurun_out/bench_r2_v5_B_g1.json').read().strip().splitlines()[-1])
print('value %.0f e2e %.0f ms %.2f e2e_ms %.2f'%(d['value'], d['e2e']['value'], d['ms_per_step'], d['e2e']['ms_per_step']), d['clocks'], d['config']['step_ms_rank0'])
PY
timeout 600 python -m pytest tests -x -q -m gpu >> $L 2>&1
tail -30 $L
