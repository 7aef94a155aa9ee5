#!/usr/bin/env python
"""Quick kernel timing for tuning: house.toml 1080p, given spp; prints ms, Mrays/s and an
md5 of the accumulation buffer (all build variants must print the same hash).
usage: RT_BACKEND_LIB=path tools/quick_bench.py [spp] [reps] [scene] [w] [h]"""
import hashlib
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import _pkg

rt = _pkg.load()
spp = int(sys.argv[1]) if len(sys.argv) > 1 else 32
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
scene_name = sys.argv[3] if len(sys.argv) > 3 else "house"
w = int(sys.argv[4]) if len(sys.argv) > 4 else 1920
h = int(sys.argv[5]) if len(sys.argv) > 5 else 1080
import time
t0 = time.perf_counter()
if scene_name.startswith("synthetic:"):
    # BASELINE.json configs[3]: the house + suzanne meshes instanced on an N x N grid
    scene = rt.Scene.synthetic_field(os.path.join(ROOT, "assets"), int(scene_name.split(":")[1]))
else:
    scene = rt.Scene.load_toml(os.path.join(ROOT, "assets", "scenes", f"{scene_name}.toml"))
build_s = time.perf_counter() - t0
ew, eh = (int(v) for v in os.environ.get("QB_ENV", "2048x1024").split("x"))  # diagnostic: environment size
env = rt.EnvironmentMaps([rt.Environment.synthetic(ew, eh)])
st = rt.State(scene, env, w, h, max_bounces=8)
if os.environ.get("QB_FIRST_HIT"):
    # coherent primary rays only (BASELINE.json configs[2]): kernel + device-to-host copies
    st.first_hit(0)
    t0 = time.perf_counter()
    for _ in range(reps):
        slots, _t = st.first_hit(0)
    dt = (time.perf_counter() - t0) / reps
    kms = st.first_hit_kernel_ms()
    print(f"{scene_name} {w}x{h} first-hit: kernel {kms:.3f} ms = {w * h / kms / 1e3:.0f} Mrays/s; "
          f"{dt * 1e3:.3f} ms per call incl. 2 D2H copies, hit fraction {(slots != 0xFFFFFFFF).mean():.3f}, "
          f"{scene.counts()['triangles']} triangles, depth {scene.bvh_depth}, scene build {build_s:.2f} s")
    sys.exit(0)
if not os.environ.get("QB_NO_WARMUP"):   # (ncu whole-step passes count every launch of ONE render)
    st.render_samples(2)  # warm-up
    st.sync()
best = 1e30
for _ in range(reps):
    st.reset()
    st.reset_counters()
    st.timer_start()
    st.render_samples(spp)
    ms = st.timer_stop()
    best = min(best, ms)
c = st.counters()
rays = c["rays"] + c["shadow_rays"]
digest = hashlib.md5(st.read_accum().tobytes()).hexdigest()[:12]
print(f"{os.path.basename(os.environ.get('RT_BACKEND_LIB', 'default')):16} {scene_name} {w}x{h} spp={spp} "
      f"best {best:9.3f} ms  {rays / best / 1e3:9.1f} Mrays/s  md5 {digest}"
      + (f"  ({scene.counts()['triangles']} triangles, depth {scene.bvh_depth}, build {build_s:.1f} s)"
         if scene_name.startswith("synthetic") else ""))
