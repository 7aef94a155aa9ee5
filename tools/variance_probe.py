#!/usr/bin/env python
"""Where do the occasional slow renders come from? Per render: device time (CUDA events), the host
time of the enqueue, and around the whole series the NVML violation-time counters and the
`nvidia-smi -q -d PERFORMANCE` throttle counters. usage: tools/variance_probe.py [spp] [renders]"""
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import _pkg

rt = _pkg.load()
spp = int(sys.argv[1]) if len(sys.argv) > 1 else 32
renders = int(sys.argv[2]) if len(sys.argv) > 2 else 80
scene = rt.Scene.load_toml(os.path.join(ROOT, "assets", "scenes", "house.toml"))
env = rt.EnvironmentMaps([rt.Environment.synthetic(2048, 1024)])
st = rt.State(scene, env, 1920, 1080, max_bounces=8)
st.render_samples(spp)
st.sync()

import pynvml
pynvml.nvmlInit()
h = pynvml.nvmlDeviceGetHandleByIndex(0)
POLICIES = {n: getattr(pynvml, n) for n in dir(pynvml) if n.startswith("NVML_PERF_POLICY_") and n != "NVML_PERF_POLICY_COUNT"}


def violations():
    out = {}
    for name, pol in POLICIES.items():
        try:
            v = pynvml.nvmlDeviceGetViolationStatus(h, pol)
            out[name.replace("NVML_PERF_POLICY_", "")] = int(v.violationTime)
        except pynvml.NVMLError:
            pass
    return out


def smi_perf():
    try:
        return subprocess.run(["nvidia-smi", "-i", "0", "-q", "-d", "PERFORMANCE"], capture_output=True, text=True, timeout=20).stdout
    except Exception as e:
        return f"nvidia-smi failed: {e}"


print(smi_perf())
v0 = violations()
rows = []
for i in range(renders):
    st.reset()
    t0 = time.perf_counter()
    st.timer_start()
    st.render_samples(spp)
    t1 = time.perf_counter()
    ms = st.timer_stop()
    t2 = time.perf_counter()
    clock = pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM) if os.environ.get("VP_CLOCK") else 0
    rows.append((ms, (t1 - t0) * 1e3, (t2 - t0) * 1e3, clock))
v1 = violations()
base = min(r[0] for r in rows)
print(f"spp {spp}: device ms min {base:.2f} median {sorted(r[0] for r in rows)[len(rows) // 2]:.2f} max {max(r[0] for r in rows):.2f}")
print("render: device_ms enqueue_ms wall_ms clock")
for i, r in enumerate(rows):
    flag = " <-- slow" if r[0] > 1.15 * base else ""
    print(f"{i:3d}: {r[0]:8.2f} {r[1]:8.3f} {r[2]:8.2f} {r[3]}{flag}")
print("violation time deltas (ns):", {k: v1[k] - v0[k] for k in v0 if v1.get(k, 0) != v0[k]}, "all:", v1)
print(smi_perf())
