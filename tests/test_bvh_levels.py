"""The level-synchronous BVH build of the device builder (csrc/rt_bvh_build.cuh), checked on the
CPU through its numpy model (oracle/bvh_levels.py): the closed form of the reference's swap
partition (src/bvh.rs:304-315) against the literal loop, and whole trees against the recursive
host builder (host/bvh.cpp) byte for byte."""
import itertools
import os

import numpy as np
import pytest

import bvh_levels
from conftest import ASSETS


def test_partition_closed_form_exhaustive():
    for n in range(0, 11):
        for flags in itertools.product([False, True], repeat=n):
            order, split = bvh_levels.partition_literal(flags)
            dest, good = bvh_levels.partition_dest(flags)
            assert good == split
            got = [None] * n
            for p in range(n):
                got[dest[p]] = p
            assert got == order, (flags, got, order)


def test_partition_closed_form_random():
    rng = np.random.default_rng(7)
    for _ in range(300):
        n = int(rng.integers(1, 400))
        flags = rng.random(n) < rng.random()
        order, split = bvh_levels.partition_literal(list(flags))
        dest, good = bvh_levels.partition_dest(flags)
        assert good == split
        got = np.empty(n, dtype=np.int64)
        got[dest] = np.arange(n)
        assert got.tolist() == order


def primitive_boxes(scene):
    """What init_records_kernel computes: boxes in build_array order (bvh.rs:40-72)."""
    sp = scene.spheres_words()
    pb = scene.plane_bounds()
    tri = scene.triangles_words()
    v = scene.vertices()
    lo, hi, types, idx = [], [], [], []
    for i in range(len(sp)):
        pos, r = sp[i, 0:3], sp[i, 3]
        lo.append(pos - r); hi.append(pos + r); types.append(0); idx.append(i)
    for i in range(len(pb)):
        lo.append(pb[i, 0:3]); hi.append(pb[i, 3:6]); types.append(1); idx.append(i)
    if len(tri):
        pts = v[tri[:, 0:3].astype(np.int64)]     # [T, 3 corners, 3]
        tlo = np.full((len(tri), 3), np.float32(3.40282347e38), np.float32)
        thi = -tlo
        for c in range(3):                         # f32::min / max keep the first argument on ties
            tlo = np.where(pts[:, c] < tlo, pts[:, c], tlo)
            thi = np.where(pts[:, c] > thi, pts[:, c], thi)
        lo.extend(tlo); hi.extend(thi); types.extend([2] * len(tri)); idx.extend(range(len(tri)))
    return (np.array(lo, np.float32).reshape(-1, 3), np.array(hi, np.float32).reshape(-1, 3),
            np.array(types, np.uint32), np.array(idx, np.uint32))


def assert_same_tree(scene):
    nodes, prims, depth = bvh_levels.build(*primitive_boxes(scene))
    want_nodes = scene.bvh_nodes_words()
    want_prims = scene.primitives_words()
    assert prims.tobytes() == want_prims.tobytes()
    assert nodes.shape == want_nodes.shape
    assert nodes.tobytes() == want_nodes.tobytes()
    assert depth == scene.bvh_depth


@pytest.mark.parametrize("name", ["house", "default", "suzanne"])
def test_level_build_equals_host_builder(rt, name):
    assert_same_tree(rt.Scene.load_toml(os.path.join(ASSETS, "scenes", f"{name}.toml")))


def test_level_build_equals_host_builder_field(rt):
    assert_same_tree(rt.Scene.synthetic_field(ASSETS, 2))


def test_level_build_signed_zeros_and_degenerate_leaves(rt):
    """-0.0 / +0.0 box corners resolve by position (f32::min keeps its first argument), and more
    than five primitives with one centroid stay one leaf (bvh.rs:238-244)."""
    b = rt.SceneBuilder()
    b.add_material((0.8, 0.8, 0.8), 1.0, 0.0)
    rng = np.random.default_rng(3)
    for _ in range(9):   # one centroid, nine spheres
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
    assert_same_tree(b.build())


def test_level_build_small_inputs(rt):
    for count in (1, 5, 6):
        b = rt.SceneBuilder()
        b.add_material((0.8, 0.8, 0.8), 1.0, 0.0)
        for i in range(count):
            b.add_sphere((float(i), 0.5 * i, 0.0), 0.3, 0)
        assert_same_tree(b.build())


def test_level_build_random_soups(rt):
    """Triangle soups on a quarter-unit lattice: many exact ties and zeros of both signs."""
    rng = np.random.default_rng(11)
    for count in (255, 257, 1025, 5000):
        s = rt.SceneBuilder()
        s.add_material((0.8, 0.8, 0.8), 1.0, 0.0)
        v = (np.round(rng.uniform(-8, 8, size=(count * 3, 3)) * 4) / 4).astype(np.float32)
        v[rng.random(len(v)) < 0.02, 1] = -0.0
        t = np.zeros((count, 6), np.uint32)
        t[:, 0:3] = np.arange(count * 3, dtype=np.uint32).reshape(count, 3)
        s.add_mesh_raw(v, np.array([[0, 1, 0]], np.float32), t, 0)
        assert_same_tree(s.build())
