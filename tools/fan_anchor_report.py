#!/usr/bin/env python
"""What moves if wavefront_obj anchored its polygon fans at the FIRST corner instead of the last?

The crate (11.0.0, Cargo.lock of the reference) is not in /root/reference and cannot be fetched, so
its fan triangulation (reference call site: src/mesh.rs:30, 50-70) is restated from memory as
last-anchored. This script builds every shipped scene both ways through the numpy front-end
(oracle/scene_oracle.py -> oracle/independent.py), renders both with the CPU oracle and records
what differs: triangle vertex order, BVH topology and primitive order, first-hit slot maps,
hit distances, and the image. Output: tests/golden/fan_anchor_report.json (checked by
tests/test_oracle_shader.py::test_fan_anchor_report_is_current).  usage: tools/fan_anchor_report.py [--check]
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle")):
    if p not in sys.path:
        sys.path.insert(0, p)
import _pkg  # noqa: E402
import independent  # noqa: E402
import oracle as orc  # noqa: E402
import scene_oracle  # noqa: E402

W, H, SPP = 320, 180, 4


def build(path, anchor):
    scene_oracle.FAN_ANCHOR = anchor
    try:
        ref = scene_oracle.load_scene_toml(path)
        nodes, ordered = scene_oracle.build_bvh(ref)
        return independent.IndependentScene(path), ref, nodes, ordered
    finally:
        scene_oracle.FAN_ANCHOR = "last"


def report():
    rt = _pkg.load()
    sky = rt.Environment.synthetic(64, 32)
    env = independent.IndependentEnvironments([sky.rgb])
    out = {"frame": f"{W}x{H}", "spp": SPP, "max_bounces": 8, "scenes": {}}
    for name in ("house", "default", "suzanne"):
        path = os.path.join(ROOT, "assets", "scenes", f"{name}.toml")
        a_scene, a_ref, a_nodes, a_ord = build(path, "last")
        b_scene, b_ref, b_nodes, b_ord = build(path, "first")
        topo = lambda nodes: [(idx, length, axis) for _mn, _mx, idx, length, axis in nodes]
        oa, ob = orc.Oracle(a_scene, env, W, H, max_bounces=8), orc.Oracle(b_scene, env, W, H, max_bounces=8)
        sa, ta, _ = oa.first_hit(0)
        sb, tb, _ = ob.first_hit(0)
        ia, _ = oa.render(0, SPP)
        ib, _ = ob.render(0, SPP)
        hit = sa != 0xFFFFFFFF
        rel = np.abs(ta - tb)[hit] / np.maximum(1.0, ta[hit])
        diff = (ia[..., :3] - ib[..., :3]).astype(np.float64)
        out["scenes"][name] = {
            "triangles": len(a_ref.triangles),
            "triangles_with_other_vertex_order": int(sum(x[:6] != y[:6] for x, y in zip(a_ref.triangles, b_ref.triangles))),
            "bvh_topology_identical": topo(a_nodes) == topo(b_nodes),
            "bvh_bounds_identical": all(x[0].tobytes() == y[0].tobytes() and x[1].tobytes() == y[1].tobytes()
                                        for x, y in zip(a_nodes, b_nodes)),
            "primitive_order_identical": [tuple(p) for p in a_ord] == [tuple(p) for p in b_ord],
            "first_hit_slots_differing": int((sa != sb).sum()),
            "first_hit_pixels_hit": int(hit.sum()),
            "first_hit_t_bits_differing": int((ta.view(np.uint32) != tb.view(np.uint32)).sum()),
            "first_hit_t_max_rel_diff": float(rel.max()) if rel.size else 0.0,
            "image_pixels_differing": int((ia[..., :3] != ib[..., :3]).any(axis=-1).sum()),
            "image_rel_rmse": float(np.sqrt((diff ** 2).mean()) / max(1e-12, np.sqrt((ia[..., :3].astype(np.float64) ** 2).mean()))),
        }
    return out


if __name__ == "__main__":
    path = os.path.join(ROOT, "tests", "golden", "fan_anchor_report.json")
    rep = report()
    if "--check" in sys.argv:
        with open(path) as fh:
            want = json.load(fh)
        sys.exit(0 if want == rep else 1)
    with open(path, "w") as fh:
        json.dump(rep, fh, indent=1)
    print(json.dumps(rep, indent=1))
