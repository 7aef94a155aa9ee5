#!/bin/bash
export PYTHONUNBUFFERED=1
mkdir -p gpurun_out
L=gpurun_out/r2_eleventh.log
: > $L
echo "=== graph loop, 100 ms" >> $L
timeout 300 python tools/sampler_probe.py 100 32 >> $L 2>&1
echo "=== host-paced loop, 100 ms" >> $L
RT_WAVE_GRAPH=0 timeout 300 python tools/sampler_probe.py 100 32 >> $L 2>&1
cat $L
