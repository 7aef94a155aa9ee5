#!/usr/bin/env python
"""Builds the BVH of the synthetic field on the device once (for an ncu pass over its kernels) and
prints the timings. usage: tools/bvh_build_probe.py [grid]"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import _pkg

rt = _pkg.load()
grid = int(sys.argv[1]) if len(sys.argv) > 1 else 99
t0 = time.perf_counter()
scene = rt.Scene.synthetic_field(os.path.join(ROOT, "assets"), grid, build_bvh=False)
t1 = time.perf_counter()
b = rt.DeviceBvhBuilder(0)
if not os.environ.get("QB_NO_WARMUP"):
    b.build(rt.Scene.synthetic_field(os.path.join(ROOT, "assets"), 2, build_bvh=False))
out = b.build(scene)
print(f"grid {grid}: {scene.counts()['triangles']} triangles generated in {t1 - t0:.2f} s; device build {out['device_ms']:.2f} ms, "
      f"call {out['build_call_ms']:.1f} ms, fetch {out['fetch_ms']:.1f} ms, install {out['install_ms']:.1f} ms; "
      f"{out['node_count']} nodes, depth {out['depth']}")
