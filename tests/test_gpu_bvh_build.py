"""rt_bvh_build (csrc/rt_bvh_build.cuh): build_bvh (src/bvh.rs:13-28, 215-337) on the device must
return the host builder's node array and primitive order byte for byte — which the CPU suite in
turn pins to the SURVEY goldens and to oracle/scene_oracle.py."""
import os
import time

import numpy as np
import pytest

from conftest import ASSETS

pytestmark = pytest.mark.gpu


def _check(rt, builder, scene):
    want_nodes, want_prims, want_depth = scene.bvh_nodes_words(), scene.primitives_words(), scene.bvh_depth
    out = builder.build(scene, install=False)
    assert out["primitives"].tobytes() == want_prims.tobytes(), (
        f"{int((out['primitives'] != want_prims).any(axis=1).sum())} primitive slots differ")
    assert out["nodes"].shape == want_nodes.shape
    bad = np.argwhere((out["nodes"] != want_nodes).any(axis=1))
    assert len(bad) == 0, f"{len(bad)} nodes differ, first {bad[:3].ravel().tolist()}: {out['nodes'][bad[0, 0]]} vs {want_nodes[bad[0, 0]]}"
    assert out["depth"] == want_depth
    return out


@pytest.fixture(scope="module")
def builder(rt):
    b = rt.DeviceBvhBuilder(0)
    yield b
    b.close()


@pytest.mark.parametrize("name", ["house", "default", "suzanne"])
def test_device_bvh_equals_host_bvh(rt, builder, name):
    _check(rt, builder, rt.Scene.load_toml(os.path.join(ASSETS, "scenes", f"{name}.toml")))


@pytest.mark.parametrize("grid", [1, 2, 7, 20])
def test_device_bvh_equals_host_bvh_field(rt, builder, grid):
    _check(rt, builder, rt.Scene.synthetic_field(ASSETS, grid))


def test_device_bvh_signed_zeros_degenerate_leaves_small_inputs(rt, builder):
    rng = np.random.default_rng(3)
    b = rt.SceneBuilder()
    b.add_material((0.8, 0.8, 0.8), 1.0, 0.0)
    for _ in range(9):   # one centroid, nine spheres: a leaf of nine (bvh.rs:238-244)
        b.add_sphere((1.0, 2.0, 3.0), float(rng.uniform(0.1, 0.4)), 0)
    for i in range(40):
        z = 0.0 if i % 3 else -0.0
        b.add_sphere((float(rng.uniform(-4, 4)), z, float(rng.uniform(-4, 4))), 0.0 if i % 2 else 0.25, 0)
    verts = np.array([[-0.0, 0.0, 0.0], [1.0, -0.0, 0.0], [0.0, 1.0, -0.0], [0.0, -0.0, 1.0],
                      [-1.0, 0.0, -0.0], [0.0, -1.0, 0.0]], np.float32)
    tris = np.array([[0, 1, 2, 0, 0, 0], [0, 2, 3, 0, 0, 0], [0, 3, 4, 0, 0, 0], [0, 4, 5, 0, 0, 0],
                     [1, 2, 3, 0, 0, 0], [2, 3, 4, 0, 0, 0], [3, 4, 5, 0, 0, 0]], np.uint32)
    b.add_mesh_raw(verts, np.array([[0, 1, 0]], np.float32), tris, 0)
    b.add_plane((-0.0, 0.0, -0.0), (0, 0, 5), (5, 0, 0), 0)
    _check(rt, builder, b.build())
    for count in (1, 5, 6, 300):
        s = rt.SceneBuilder()
        s.add_material((0.8, 0.8, 0.8), 1.0, 0.0)
        for i in range(count):
            s.add_sphere((float(rng.uniform(-9, 9)), float(rng.uniform(-9, 9)), float(rng.uniform(-9, 9))), 0.3, 0)
        _check(rt, builder, s.build())


def test_device_bvh_random_soups(rt, builder):
    """Random triangle soups with many exact ties (snapped coordinates) across sizes that cross the
    block and scan-tile boundaries of the kernels."""
    rng = np.random.default_rng(11)
    for count in (255, 256, 257, 1023, 1024, 1025, 5000, 70001):
        s = rt.SceneBuilder()
        s.add_material((0.8, 0.8, 0.8), 1.0, 0.0)
        v = np.round(rng.uniform(-8, 8, size=(count * 3, 3)) * 4) / 4     # quarter-unit lattice: ties, zeros
        v = v.astype(np.float32)
        v[rng.random(len(v)) < 0.02, 1] = -0.0
        t = np.zeros((count, 6), np.uint32)
        t[:, 0:3] = np.arange(count * 3, dtype=np.uint32).reshape(count, 3)
        s.add_mesh_raw(v, np.array([[0, 1, 0]], np.float32), t, 0)
        _check(rt, builder, s.build())


def test_device_bvh_ten_million_triangles(rt, builder):
    """BASELINE.json configs[3]'s scene: same tree as the host builder. The device part is timed with
    CUDA events and asserted (measured: 30-42 ms); the whole call with the upload of the primitive
    arrays and the download of nodes and primitives is printed (measured 0.18-0.43 s in a fresh
    process, tools/bvh_build_probe.py) but only sanity-checked: it is host page-fault and PCIe time
    and swings several-fold on a busy host."""
    scene = rt.Scene.synthetic_field(ASSETS, 99)
    assert scene.counts()["triangles"] == 9997020
    builder.build(rt.Scene.synthetic_field(ASSETS, 3), install=False)   # warm the context
    t0 = time.perf_counter()
    out = _check(rt, builder, scene)
    _ = time.perf_counter() - t0
    total_ms = out["build_call_ms"] + out["fetch_ms"]
    print(f"device BVH build, 9 997 021 primitives: device {out['device_ms']:.1f} ms, call {out['build_call_ms']:.1f} ms, "
          f"fetch {out['fetch_ms']:.1f} ms, {out['node_count']} nodes, depth {out['depth']}")
    assert out["device_ms"] <= 250.0, f"device build took {out['device_ms']:.0f} ms on the GPU"
    assert total_ms <= 20000.0, f"device build call took {total_ms:.0f} ms"
    # installed, the scene renders like the host-built one (same bytes, so only a smoke check)
    fresh = rt.Scene.synthetic_field(ASSETS, 3, build_bvh=False)
    assert fresh.counts()["bvh_nodes"] == 0
    built = builder.build(fresh)
    assert fresh.counts()["bvh_nodes"] == built["node_count"] and fresh.bvh_depth == built["depth"]
    assert fresh.bvh_nodes_words().tobytes() == rt.Scene.synthetic_field(ASSETS, 3).bvh_nodes_words().tobytes()


def _fuzz_scene(rt, rng):
    """Mixed primitive soups: clustered, planar, duplicated and lattice-snapped triangles, spheres of
    radius zero, axis-aligned and skewed planes."""
    b = rt.SceneBuilder()
    b.add_material((0.8, 0.8, 0.8), 1.0, 0.0)
    for _ in range(int(rng.integers(0, 40))):
        b.add_sphere(tuple(float(v) for v in rng.uniform(-6, 6, 3)), float(rng.choice([0.0, 0.1, 0.7])), 0)
    for _ in range(int(rng.integers(0, 4))):
        p = rng.uniform(-5, 5, 3)
        b.add_plane(tuple(map(float, p)), tuple(map(float, rng.uniform(-3, 3, 3))), tuple(map(float, rng.uniform(-3, 3, 3))), 0)
    count = int(rng.choice([0, 1, 7, 60, 700, 5000, 20000]))
    if count:
        kind = int(rng.integers(0, 4))
        centers = rng.uniform(-8, 8, size=(count, 1, 3))
        if kind == 1:      # a few tight clusters: deep, unbalanced trees
            centers = rng.choice(rng.uniform(-8, 8, size=(5, 3)), size=count)[:, None, :] + rng.normal(0, 0.05, (count, 1, 3))
        elif kind == 2:    # everything in one plane: zero extents, signed zeros
            centers[:, :, 1] = 0.0
        v = centers + rng.normal(0, 0.3, size=(count, 3, 3))
        if kind == 2:
            v[:, :, 1] = np.where(rng.random((count, 3)) < 0.5, 0.0, -0.0)
        if kind == 3:      # lattice: exact ties of centroids and bounds, many duplicates
            v = np.round(v * 2) / 2
        v = v.reshape(-1, 3).astype(np.float32)
        t = np.zeros((count, 6), np.uint32)
        t[:, 0:3] = np.arange(count * 3, dtype=np.uint32).reshape(count, 3)
        b.add_mesh_raw(v, np.array([[0, 1, 0]], np.float32), t, 0)
    elif rng.random() < 0.5:
        b.add_sphere((0.0, 0.0, 0.0), 1.0, 0)   # never an empty scene
    return b


@pytest.mark.parametrize("seed", range(int(os.environ.get("RT_BVH_FUZZ_SEEDS", "24"))))
def test_device_bvh_fuzz(rt, builder, seed):
    _check(rt, builder, _fuzz_scene(rt, np.random.default_rng(500 + seed)).build())


def test_device_bvh_rejects_bad_input(rt, builder):
    import ctypes as C
    lib = rt._ffi.backend_lib()
    inp = rt._ffi.RtBvhBuildInput()
    n, d = C.c_uint32(), C.c_uint32()
    assert lib.rt_bvh_build(builder._ctx, C.byref(inp), C.byref(n), C.byref(d), None) == rt._ffi.RT_ERR_SCENE
    assert b"no primitives" in lib.rt_last_error(builder._ctx)      # bvh.rs:222
    verts = np.zeros((3, 4), np.float32)
    tris = np.array([[0, 1, 7, 0, 0, 0, 0]], np.uint32)            # vertex 7 does not exist
    inp.vertices, inp.vertex_count = verts.ctypes.data, 3
    inp.triangles, inp.triangle_count = tris.ctypes.data, 1
    assert lib.rt_bvh_build(builder._ctx, C.byref(inp), C.byref(n), C.byref(d), None) == rt._ffi.RT_ERR_INVALID_ARGUMENT
    assert b"vertex out of range" in lib.rt_last_error(builder._ctx)
    assert lib.rt_bvh_build_fetch(builder._ctx, None, None) == rt._ffi.RT_ERR_STATE
