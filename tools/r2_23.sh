#!/bin/bash
export PYTHONUNBUFFERED=1
mkdir -p gpurun_out
L=gpurun_out/r2_23.log
: > $L
timeout 900 python -m pytest tests -x -q -m gpu >> $L 2>&1
echo "--- pytest exit $?" >> $L
timeout 200 python tools/quick_bench.py 256 3 house >> $L 2>&1
timeout 200 python tools/quick_bench.py 32 5 house >> $L 2>&1
tail -8 $L
