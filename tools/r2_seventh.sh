#!/bin/bash
# why is bench.py's device-timed step slower than its own e2e step? sampler interval / graph / stream
export PYTHONUNBUFFERED=1
mkdir -p gpurun_out
L=gpurun_out/r2_seventh.log
: > $L
one() { echo "=== $*" >> $L; env "$@" timeout 600 python bench.py --config B --steps 5 --warmup 3 --no-cpu-baseline 2>> $L | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('value %.0f e2e %.0f ms %.2f render_ms %.2f launches %d clocks %s'%(d['value'],d['e2e']['value'],d['ms_per_step'],d['config']['render_ms_rank_max'],d['gpu_launches'],d['clocks']))" >> $L; }
one X=1
one BENCH_SAMPLER_MS=1000
one RT_WAVE_GRAPH=0
one RT_WAVE_GRAPH=0 BENCH_SAMPLER_MS=1000
one X=2
tail -12 $L
