#!/bin/bash
# bench.py's device-timed step: graph loop vs host-paced loop, sampler interval
export PYTHONUNBUFFERED=1
mkdir -p gpurun_out
L=gpurun_out/r2_seventh.log
: > $L
one() { echo "=== $*" >> $L; env "$@" timeout 600 python bench.py --config B --steps 5 --warmup 3 --no-cpu-baseline 2>> $L | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('value %.0f e2e %.0f ms %.2f render_ms %.2f launches %d clocks %s'%(d['value'],d['e2e']['value'],d['ms_per_step'],d['config']['render_ms_rank_max'],d['gpu_launches'],d['clocks']))" >> $L; }
one X=1
one RT_WAVE_GRAPH=0
one BENCH_SAMPLER_MS=1000
one RT_WAVE_GRAPH=0 BENCH_SAMPLER_MS=1000
one X=2
one RT_WAVE_GRAPH=0
timeout 300 python -m pytest tests/test_gpu_parity.py tests/test_gpu_wavefront.py -x -q >> $L 2>&1
tail -14 $L
