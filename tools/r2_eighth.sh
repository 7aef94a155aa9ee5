#!/bin/bash
# state check of the restored tree on a 2-GPU box: the whole GPU suite (group tests included), the
# default bench line, the strong-scaling line at N=2, and a slot-pool sweep under the graph loop
export PYTHONUNBUFFERED=1
mkdir -p gpurun_out
L=gpurun_out/r2_eighth.log
: > $L
nvidia-smi -L >> $L
run() { echo "=== $*" >> $L; timeout ${T:-300} "$@" >> $L 2>&1; echo "--- exit $?" >> $L; }
T=900 run python -m pytest tests -x -q -m gpu
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_r2_v3_B_g1.json 2>> $L
echo "--- bench g1 exit $?" >> $L
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
  bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/bench_r2_v3_B_strong_g2.json 2>> $L
echo "--- bench g2 exit $?" >> $L
for f in gpurun_out/bench_r2_v3_B_g1.json gpurun_out/bench_r2_v3_B_strong_g2.json; do
python - $f >> $L <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(d['n_gpus'], d['scaling'], 'value %.0f e2e %.0f ms %.2f'%(d['value'], d['e2e']['value'], d['ms_per_step']), d['config'].get('render_ms_rank_max'), d['config'].get('reduce_ms'), d.get('reduce_check'), d.get('roofline'))
except Exception as e:
    print('unparsed', e)
PY
done
# slot-pool sweep: does an L2-sized pool pay now that the loop is one graph launch?
for spp in 32 256; do
for s in 16777216 8388608 4194304 2097152 1048576 524288; do
  for g in 2 4; do
    echo "=== slots $s groups $g spp $spp" >> $L
    RT_WAVE_SLOTS=$s RT_WAVE_GROUPS=$g timeout 200 python tools/quick_bench.py $spp 2 house >> $L 2>&1
  done
done
done
tail -60 $L
