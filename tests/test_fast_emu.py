"""CPU check of the production traversal (rsoderh-raytracing_b200/csrc/rt_fast.cuh).

tests/emu/fast_emu.cu compiles trace_fast and the fast-layout builder for the host. Here
its answers — slot / fallback code and hit distance for closest-hit rays, hit-or-not for
shadow rays — are compared bit for bit with the oracle's literal cast_ray / cast_ray_bvh
(un-pruned, reference order) on camera rays, random rays, secondary rays that start on
surfaces, axis-aligned rays (infinite 1/d components), and on scenes built to exercise the
paths the shipped scenes do not: exact ties across leaves, mixed leaves, many analytic
primitives (hoisting off), a root that is a leaf, leaves larger than the inline length field.
"""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from conftest import ASSETS, ROOT

EMU_DIR = os.path.join(ROOT, "tests", "emu")
EMU_LIB = os.path.join(EMU_DIR, "libfast_emu.so")


@pytest.fixture(scope="module")
def emu():
    src = os.path.join(EMU_DIR, "fast_emu.cu")
    deps = [src] + [os.path.join(ROOT, "rsoderh-raytracing_b200", "csrc", f)
                    for f in ("rt_fast.cuh", "rt_fast_host.hpp", "rt_device.cuh")]
    if not os.path.exists(EMU_LIB) or any(os.path.getmtime(d) > os.path.getmtime(EMU_LIB) for d in deps):
        subprocess.run(["make", "-s", "emu"], cwd=ROOT, check=True)
    lib = C.CDLL(EMU_LIB)
    lib.emu_scene_create.restype = C.c_void_p
    lib.emu_scene_create.argtypes = [C.c_void_p]
    lib.emu_scene_free.argtypes = [C.c_void_p]
    lib.emu_scene_hoisted.argtypes = [C.c_void_p]
    lib.emu_trace.argtypes = [C.c_void_p, C.c_uint32, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p,
                              C.c_void_p, C.c_void_p]
    lib.emu_trace_flat.argtypes = lib.emu_trace.argtypes
    lib.emu_scene_flat_leaves.argtypes = [C.c_void_p]
    lib.emu_check_layout.argtypes = [C.c_void_p, C.c_void_p, C.POINTER(C.c_uint32 * 8)]
    lib.emu_check_layout.restype = C.c_int
    return lib


class Emu:
    def __init__(self, lib, scene):
        self.lib, self.scene = lib, scene
        self.h = lib.emu_scene_create(C.addressof(scene.desc))

    def __del__(self):
        self.lib.emu_scene_free(self.h)

    def check_layout(self):
        """(status, stats) of the structural check of the effective tree."""
        out = (C.c_uint32 * 8)()
        st = self.lib.emu_check_layout(self.h, C.addressof(self.scene.desc), C.byref(out))
        keys = ("interior", "leaves", "triangles", "depth", "pre_leaves", "pre_boxes", "pre_prims", "hoisted")
        return st, dict(zip(keys, list(out)))

    @property
    def hoisted(self):
        return bool(self.lib.emu_scene_hoisted(self.h))

    @property
    def flat_leaves(self):
        """Leaves of the flat walk program; 0 when the scene has none (too many leaves, not hoisted)."""
        return int(self.lib.emu_scene_flat_leaves(self.h))

    def trace(self, o, d, shadow=False, flat=False):
        o = np.ascontiguousarray(o, dtype=np.float32)
        d = np.ascontiguousarray(d, dtype=np.float32)
        n = len(o)
        code = np.empty(n, np.uint32)
        t = np.empty(n, np.float32)
        fn = self.lib.emu_trace_flat if flat else self.lib.emu_trace
        fn(self.h, n, o.ctypes.data, d.ctypes.data, int(shadow), code.ctypes.data, t.ctypes.data, None)
        return code, t


def _unit(v):
    v = np.asarray(v, dtype=np.float32)
    return (v / np.sqrt((v * v).sum(-1, keepdims=True, dtype=np.float32))).astype(np.float32)


def _random_rays(rng, n, lo, hi):
    o = rng.uniform(lo, hi, size=(n, 3)).astype(np.float32)
    d = _unit(rng.normal(size=(n, 3)))
    return o, d


def _check(emu_scene, ora, o, d, what):
    ocode, ot = ora.cast_rays(o, d)
    socc, _ = ora.cast_rays(o, d, shadow=True)
    # the stack walk always; the flat walk (leaf-box mask) whenever the scene has its program
    for flat in ([False, True] if emu_scene.flat_leaves else [False]):
        tag = f"{what} ({'flat' if flat else 'stack'} walk)"
        code, t = emu_scene.trace(o, d, flat=flat)
        bad = np.flatnonzero(code != ocode)
        assert len(bad) == 0, f"{tag}: {len(bad)} of {len(o)} hits differ, first {bad[:5]}: " \
                              f"{code[bad[:5]]} vs {ocode[bad[:5]]} o={o[bad[:2]]} d={d[bad[:2]]}"
        assert t.tobytes() == ot.tobytes(), f"{tag}: distances differ"
        scode, _ = emu_scene.trace(o, d, shadow=True, flat=flat)
        assert np.array_equal(scode != 0xFFFFFFFF, socc != 0), f"{tag}: shadow-ray occlusion differs"
    return code, t


def _exercise(emu_scene, ora, rng, lo, hi, n=20000):
    o, d = _random_rays(rng, n, lo, hi)
    code, t = _check(emu_scene, ora, o, d, "random rays")
    # secondary rays: start exactly on the surfaces found above, like shader.wgsl:1249,1295
    hit = code != 0xFFFFFFFF
    o2 = (o[hit] + d[hit] * t[hit, None]).astype(np.float32)
    d2 = _unit(rng.normal(size=o2.shape))
    _check(emu_scene, ora, o2, d2, "secondary rays")
    # axis-aligned and planar directions: 1/d has infinite components, slabs produce NaN
    axes = np.array([[1, 0, 0], [-1, 0, 0], [0, 1, 0], [0, -1, 0], [0, 0, 1], [0, 0, -1],
                     [1, 1, 0], [0, -1, 1], [-1, 0, 1]], dtype=np.float32)
    o3 = np.repeat(o[:600], len(axes), axis=0)
    d3 = _unit(np.tile(axes, (600, 1)))
    _check(emu_scene, ora, o3, d3, "axis-aligned rays")
    return hit.mean()


@pytest.mark.parametrize("name,lo,hi", [("house", (-12, -1, -12), (12, 8, 12)),
                                        ("default", (-5, -1.5, -6), (7, 3, 6)),
                                        ("suzanne", (-2, -2, -2), (2, 2, 2.5))])
def test_shipped_scenes_random_rays(rt, orc, emu, env_small, name, lo, hi):
    scene = rt.Scene.load_toml(os.path.join(ASSETS, "scenes", f"{name}.toml"))
    es = Emu(emu, scene)
    assert es.hoisted
    ora = orc.Oracle(scene, env_small, 64, 36)
    frac = _exercise(es, ora, np.random.default_rng(7), lo, hi)
    assert frac > 0.05


@pytest.mark.parametrize("name,w,h", [("house", 240, 135), ("default", 160, 90), ("suzanne", 192, 108)])
def test_shipped_scenes_camera_rays(rt, orc, emu, env_small, name, w, h):
    scene = rt.Scene.load_toml(os.path.join(ASSETS, "scenes", f"{name}.toml"))
    es = Emu(emu, scene)
    ora = orc.Oracle(scene, env_small, w, h)
    slots, t, _ = ora.first_hit(3)
    o = np.empty((h * w, 3), np.float32)
    d = np.empty((h * w, 3), np.float32)
    for y in range(h):
        for x in range(w):
            o[y * w + x], d[y * w + x], _ = ora.camera_ray(x, y, 3)
    for flat in ([False, True] if es.flat_leaves else [False]):
        code, et = es.trace(o, d, flat=flat)
        assert np.array_equal(code.reshape(h, w), slots)
        assert et.tobytes() == t.tobytes()


def test_flat_program_of_shipped_scenes(rt, emu):
    # house / default fit the flat walk (few triangle leaves); suzanne (hundreds) does not
    leaves = {}
    for name in ("house", "default", "suzanne"):
        scene = rt.Scene.load_toml(os.path.join(ASSETS, "scenes", f"{name}.toml"))
        es = Emu(emu, scene)
        _, stats = es.check_layout()
        leaves[name] = (es.flat_leaves, stats["leaves"])
    assert leaves["house"][0] == leaves["house"][1] and 0 < leaves["house"][0] <= 32, leaves
    assert leaves["default"][0] == leaves["default"][1] and 0 < leaves["default"][0] <= 32, leaves
    assert leaves["suzanne"][0] == 0 and leaves["suzanne"][1] > 32, leaves


CUBE = """o cube
v -1 -1 -1
v 1 -1 -1
v 1 1 -1
v -1 1 -1
v -1 -1 1
v 1 -1 1
v 1 1 1
v -1 1 1
vn 0 0 -1
vn 0 0 1
vn -1 0 0
vn 1 0 0
vn 0 -1 0
vn 0 1 0
f 1//1 4//1 3//1 2//1
f 5//2 6//2 7//2 8//2
f 1//3 5//3 8//3 4//3
f 2//4 3//4 7//4 6//4
f 1//5 2//5 6//5 5//5
f 4//6 8//6 7//6 3//6
"""


def _cube_at(offset, scale=1.0):
    lines = []
    for line in CUBE.splitlines():
        if line.startswith("v "):
            x, y, z = (float(s) for s in line.split()[1:])
            lines.append(f"v {x * scale + offset[0]} {y * scale + offset[1]} {z * scale + offset[2]}")
        else:
            lines.append(line)
    return "\n".join(lines) + "\n"


def test_ties_across_leaves_and_duplicates(rt, orc, emu, env_small):
    # Coincident geometry everywhere: the same plane three times, two identical spheres,
    # two identical cubes, a plane lying exactly on a cube face — rays tie bit for bit and
    # the winner is decided by the reference's visiting order.
    b = rt.SceneBuilder().add_material((0.8, 0.8, 0.8), 1.0, 0.0)
    for _ in range(3):
        b.add_plane((-4, -1, -4), (0, 0, 8), (8, 0, 0), 0)
    b.add_plane((-1, 1, -1), (0, 0, 2), (2, 0, 0), 0)      # on the cube's top face
    b.add_sphere((3, 0, 0), 1.0, 0).add_sphere((3, 0, 0), 1.0, 0)
    b.add_sphere((0, 0, 3), 0.5, 0)
    b.add_mesh_obj(_cube_at((0, 0, 0)), 0).add_mesh_obj(_cube_at((0, 0, 0)), 0)
    b.add_mesh_obj(_cube_at((-3, 0, 0), 0.5), 0)
    scene = b.set_camera((0, 1, 6), 0.0, 0.0, 1.2).build()
    es = Emu(emu, scene)
    assert es.hoisted
    ora = orc.Oracle(scene, env_small, 64, 36)
    rng = np.random.default_rng(11)
    _exercise(es, ora, rng, (-5, -2, -5), (5, 4, 5), n=30000)
    # rays aimed at shared vertices, edges and face centres from many origins
    targets = np.array([[1, 1, 1], [-1, 1, -1], [0, 1, 0], [1, 0, 1], [0, -1, 0], [3, 1, 0], [0, 1, 0.5],
                        [-3, 0.5, 0], [1, 1, 0]], dtype=np.float32)
    o = rng.uniform(-5, 5, size=(4000, 3)).astype(np.float32)
    tg = targets[rng.integers(0, len(targets), size=4000)]
    _check(es, ora, o, _unit(tg - o), "aimed rays")


def test_many_analytic_primitives_not_hoisted(rt, orc, emu, env_small):
    rng = np.random.default_rng(5)
    b = rt.SceneBuilder().add_material((0.8, 0.8, 0.8), 1.0, 0.0)
    for _ in range(40):
        b.add_sphere(tuple(rng.uniform(-4, 4, 3)), float(rng.uniform(0.2, 0.8)), 0)
    for _ in range(6):
        b.add_plane(tuple(rng.uniform(-4, 4, 3)), tuple(rng.uniform(-2, 2, 3)), tuple(rng.uniform(-2, 2, 3)), 0)
    b.add_mesh_obj(_cube_at((0, 0, 0)), 0)
    scene = b.set_camera((0, 1, 6), 0.0, 0.0, 1.2).build()
    es = Emu(emu, scene)
    assert not es.hoisted
    ora = orc.Oracle(scene, env_small, 64, 36)
    _exercise(es, ora, rng, (-5, -5, -5), (5, 5, 5))


def test_tiny_scenes_root_is_leaf(rt, orc, emu, env_small):
    rng = np.random.default_rng(3)
    for build in (lambda b: b.add_sphere((0, 0, 0), 1.0, 0),
                  lambda b: b.add_plane((-1, 0, -1), (0, 0, 2), (2, 0, 0), 0).add_sphere((0, 1, 0), 0.5, 0),
                  lambda b: b.add_mesh_raw([[0, 0, 0], [1, 0, 0], [0, 1, 0]], [[0, 0, 1]], [[0, 1, 2, 0, 0, 0]], 0)):
        b = rt.SceneBuilder().add_material((0.8, 0.8, 0.8), 1.0, 0.0)
        scene = build(b).set_camera((0, 0, 3), 0.0, 0.0, 1.0).build()
        es = Emu(emu, scene)
        ora = orc.Oracle(scene, env_small, 64, 36)
        _exercise(es, ora, rng, (-2, -2, -2), (2, 2, 2), n=6000)


def test_degenerate_big_leaf(rt, orc, emu, env_small):
    # 40 triangles whose boxes share one centre cannot be split (bvh.rs:238-244): one leaf longer
    # than the inline length field, next to ordinary geometry.
    rng = np.random.default_rng(9)
    verts, tris = [], []
    for i in range(40):
        # box centre (min*0.5 + max*0.5, scene.rs:102-104) exactly at the origin for every triangle
        a = np.round(rng.uniform(0.5, 2.0, size=3) * rng.choice([-1, 1], size=3) * 8) / 8
        base = len(verts)
        verts += [a, -a, a * np.array([0.5, -0.5, 0.25])]
        tris.append([base, base + 1, base + 2, 0, 0, 0])
    verts = np.array(verts)
    b = rt.SceneBuilder().add_material((0.8, 0.8, 0.8), 1.0, 0.0)
    b.add_mesh_raw(verts, [[0, 0, 1]], tris, 0).add_mesh_obj(_cube_at((4, 0, 0)), 0)
    b.add_sphere((-4, 0, 0), 1.0, 0)
    scene = b.set_camera((0, 0, 6), 0.0, 0.0, 1.0).build()
    nodes = scene.bvh_nodes_words()
    assert nodes[:, 9].max() >= 31, "expected a leaf with >= 31 primitives"
    es = Emu(emu, scene)
    ora = orc.Oracle(scene, env_small, 64, 36)
    _exercise(es, ora, rng, (-6, -3, -3), (6, 3, 3), n=20000)


def test_effective_tree_structure(rt, emu):
    # The tree the walk uses: every triangle slot covered exactly once, children inside their
    # parents' (tight) boxes, no empty children; hoisted leaves are gone and single-child
    # nodes spliced out, so it has fewer interior records than the reference's tree.
    cases = {"house": (52, 39), "default": (12, 15), "suzanne": (968, 545)}
    for name, (tris, ref_nodes) in cases.items():
        scene = rt.Scene.load_toml(os.path.join(ASSETS, "scenes", f"{name}.toml"))
        st, stats = Emu(emu, scene).check_layout()
        assert st == 0, (name, st, stats)
        assert stats["triangles"] == tris and stats["hoisted"] == 1
        ref_interior = (ref_nodes - 1) // 2
        assert stats["interior"] == stats["leaves"] - 1 <= ref_interior
        if name != "suzanne":
            assert stats["interior"] < ref_interior and stats["pre_prims"] > 0
    # house: two analytic leaves (planes, spheres), the duplicate ground plane dropped
    st, stats = Emu(emu, rt.Scene.load_toml(os.path.join(ASSETS, "scenes", "house.toml"))).check_layout()
    assert (stats["pre_leaves"], stats["pre_prims"]) == (2, 3)
    field = rt.Scene.synthetic_field(ASSETS, 6)
    st, stats = Emu(emu, field).check_layout()
    assert st == 0 and stats["triangles"] == field.counts()["triangles"] and stats["depth"] <= field.bvh_depth
