#!/bin/bash
export PYTHONUNBUFFERED=1
mkdir -p gpurun_out
L=gpurun_out/r2_twelfth.log
: > $L
nproc >> $L; cat /sys/fs/cgroup/cpu.max >> $L 2>&1; cat /sys/fs/cgroup/cpu.stat >> $L 2>&1
echo "=== graph loop" >> $L
timeout 300 python tools/variance_probe.py 32 80 >> $L 2>&1
cat /sys/fs/cgroup/cpu.stat >> $L 2>&1
echo "=== host-paced loop, clock read after every render" >> $L
VP_CLOCK=1 RT_WAVE_GRAPH=0 timeout 300 python tools/variance_probe.py 32 60 >> $L 2>&1
echo "=== RT_KERNEL=fast (thread-per-pixel megakernel: one plain launch), 8 spp" >> $L
RT_KERNEL=fast timeout 300 python tools/variance_probe.py 8 60 2>&1 | grep -v "^ *[0-9]*: .* [0-9]*$" >> $L
tail -5 $L
