#!/bin/bash
export PYTHONUNBUFFERED=1
mkdir -p gpurun_out
L=gpurun_out/r2_26.log
: > $L
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_wavefront.py tests/test_gpu_configs.py -x -q -m gpu >> $L 2>&1
echo "--- pytest exit $?" >> $L
for spp in 1 4 32; do timeout 100 python tools/quick_bench.py $spp 10 house >> $L 2>&1; done
timeout 200 python tools/quick_bench.py 256 3 house >> $L 2>&1
timeout 100 python tools/quick_bench.py 64 10 default 640 360 >> $L 2>&1
timeout 300 python tools/quick_bench.py 16 3 synthetic:99 >> $L 2>&1
tail -12 $L
