#!/bin/bash
# why do the 128-spp renders of the N=2 strong-scaling line take 137-236 ms? per-rank step lists under
# the sampler / graph-loop variants
export PYTHONUNBUFFERED=1 BENCH_DEBUG=1
mkdir -p gpurun_out
L=gpurun_out/r2_ninth.log
: > $L
port=29520
one() {
  port=$((port+1))
  echo "=== $*" >> $L
  env "$@" timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $port \
    bench.py --gpus 2 --steps 8 --warmup 3 --no-cpu-baseline 2>> $L | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('value %.0f e2e %.0f ms %.2f render_ms %.2f clocks %s'%(d['value'],d['e2e']['value'],d['ms_per_step'],d['config']['render_ms_rank_max'],d['clocks']))" >> $L
}
one X=1
one BENCH_SAMPLER_MS=100000
one RT_WAVE_GRAPH=0
one RT_WAVE_GRAPH=0 BENCH_SAMPLER_MS=100000
one CUDA_DEVICE_MAX_CONNECTIONS=32 BENCH_SAMPLER_MS=100000
grep -v "^\*\*\*\|OMP_NUM\|^$\|NCCL version" $L | tail -40
