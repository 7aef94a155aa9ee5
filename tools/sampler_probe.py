#!/usr/bin/env python
"""Which clock / throttle query disturbs a running render? 60 renders of 32 spp (about 35 ms each)
per variant while a thread polls one NVML call (or an nvidia-smi child runs) every `interval` ms.
usage: tools/sampler_probe.py [interval_ms] [spp]"""
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import _pkg

rt = _pkg.load()
interval = float(sys.argv[1]) / 1e3 if len(sys.argv) > 1 else 0.1
spp = int(sys.argv[2]) if len(sys.argv) > 2 else 32
scene = rt.Scene.load_toml(os.path.join(ROOT, "assets", "scenes", "house.toml"))
env = rt.EnvironmentMaps([rt.Environment.synthetic(2048, 1024)])
st = rt.State(scene, env, 1920, 1080, max_bounces=8)
st.render_samples(spp)
st.sync()

import pynvml
pynvml.nvmlInit()
h = pynvml.nvmlDeviceGetHandleByIndex(0)
CALLS = {
    "none": None,
    "clock_sm": lambda: pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM),
    "event_reasons": lambda: pynvml.nvmlDeviceGetCurrentClocksEventReasons(h),
    "power": lambda: pynvml.nvmlDeviceGetPowerUsage(h),
    "temperature": lambda: pynvml.nvmlDeviceGetTemperature(h, pynvml.NVML_TEMPERATURE_GPU),
    "clock+reasons": lambda: (pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM),
                              pynvml.nvmlDeviceGetCurrentClocksEventReasons(h)),
    "smi_clock_reasons": "clocks.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.sw_power_cap",
    "smi_all": ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
                "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
                "clocks_event_reasons.sw_power_cap"),
    "none_again": None,
}


def run(name, call):
    stop = threading.Event()
    durations = []
    proc = None
    if isinstance(call, str):
        proc = subprocess.Popen(["nvidia-smi", "-i", "0", f"--query-gpu={call}", "--format=csv,noheader,nounits",
                                 "-lms", str(int(interval * 1e3))], stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
        time.sleep(1.5)   # its start-up is not what is measured
    elif call:
        def poll():
            while not stop.is_set():
                t0 = time.perf_counter()
                try:
                    call()
                except Exception as e:
                    durations.append(-1.0)
                durations.append((time.perf_counter() - t0) * 1e3)
                stop.wait(interval)
        th = threading.Thread(target=poll, daemon=True)
        th.start()
    times = []
    for _ in range(60):
        st.reset()
        st.timer_start()
        st.render_samples(spp)
        times.append(st.timer_stop())
    stop.set()
    if proc:
        proc.terminate()
        proc.wait()
    times.sort()
    q = f"query ms median {statistics.median(durations):.3f} max {max(durations):.3f} n {len(durations)}" if durations else ""
    print(f"{name:18} render ms: min {times[0]:.2f} median {times[30]:.2f} p90 {times[54]:.2f} max {times[-1]:.2f} "
          f"mean {sum(times) / len(times):.2f}  {q}", flush=True)


for name, call in CALLS.items():
    run(name, call)
