"""GPU tests of the render kernels — the persistent CTA-local wavefront (rt_local.cuh, the
product path), the launch-per-phase wavefront pipeline (rt_wave.cuh) and the fast traversal
(rt_fast.cuh):
bit-exact against the oracle and against the backend's other kernels, on the cases the
shipped scenes do not reach — tiny slot pools and several staging batches, frame sizes
with partial 8x8 tiles, exact ties across leaves, scenes whose analytic primitives are
not hoisted (with a plane that sticks out of its leaf box), a deep triangle-only tree.
"""
import os

import numpy as np
import pytest

from conftest import ASSETS
from test_fast_emu import _cube_at

pytestmark = pytest.mark.gpu


def _render(rt, scene, env, w, h, spp, bounces, begin=0, **kw):
    st = rt.State(scene, env, w, h, device=0, max_bounces=bounces, **kw)
    st.reset_counters()
    st.render_range(begin, spp)
    return st, st.read_accum()


def _report(a, b):
    bad = np.argwhere((a != b).any(axis=-1))
    return f"{len(bad)} pixels differ, first {bad[:4].tolist()}"


def test_all_kernels_agree_bit_for_bit(rt, house, env_small, monkeypatch):
    w, h = 200, 112
    images = {}
    for kernel in ("local", "wave", "fast", "pixel"):
        monkeypatch.setenv("RT_KERNEL", kernel)  # "wave" is the default (the product path)
        st, images[kernel] = _render(rt, house, env_small, w, h, 3, 8)
        c = st.counters()
        assert c["paths"] == w * h * 3
        images[kernel + "_counts"] = (c["paths"], c["rays"], c["shadow_rays"])
    assert images["local"].tobytes() == images["wave"].tobytes() == images["fast"].tobytes() == images["pixel"].tobytes()
    assert images["local_counts"] == images["wave_counts"] == images["fast_counts"] == images["pixel_counts"]


@pytest.mark.parametrize("slots,stage,window,w,h,spp", [(64, 64, 1, 203, 117, 5), (256, 1024, 4, 320, 180, 9),
                                                        (1024, 8192, 16, 161, 90, 20), (4096, 65536, 16, 161, 90, 4),
                                                        (512, 2048, 8, 97, 61, 11)])
def test_local_pools_windows_and_rings(rt, orc, house, env_small, monkeypatch, slots, stage, window, w, h, spp):
    # The persistent kernel with tiny pools (every slot traces many paths), a one-window ring
    # (stage = 64 * window: the slots wait for every commit), windows that do not divide the
    # sample count, sizes with partial 8x8 tiles: the image must not depend on any of it.
    monkeypatch.setenv("RT_KERNEL", "local")
    monkeypatch.setenv("RT_LOCAL_SLOTS", str(slots))
    monkeypatch.setenv("RT_LOCAL_STAGE", str(stage))
    monkeypatch.setenv("RT_LOCAL_WINDOW", str(window))
    st, got = _render(rt, house, env_small, w, h, spp, 8, begin=3)
    want, counters = orc.Oracle(house, env_small, w, h, max_bounces=8).render(3, spp)
    assert got.tobytes() == want.tobytes(), _report(got, want)
    c = st.counters()
    assert (c["paths"], c["rays"], c["shadow_rays"]) == (counters["paths"], counters["rays"],
                                                         counters["shadow_rays"])
    assert (got[..., 3] == 1.0).all()
    assert st.launch_count() <= 4  # clear + ONE render launch (+ the uploads' material bake)


def test_local_flat_walk_equals_stack_walk(rt, house, env_small, monkeypatch):
    # house.toml fits the flat walk (18 leaves); RT_LOCAL_FLAT=0 sends it through the N2 stack walk
    w, h = 160, 90
    monkeypatch.setenv("RT_KERNEL", "local")
    monkeypatch.setenv("RT_LOCAL_FLAT", "1")
    _, flat = _render(rt, house, env_small, w, h, 4, 8)
    monkeypatch.setenv("RT_LOCAL_FLAT", "0")
    _, stack = _render(rt, house, env_small, w, h, 4, 8)
    assert flat.tobytes() == stack.tobytes()


@pytest.mark.parametrize("slots,staging_mb,w,h,spp", [(1024, 1, 203, 117, 5), (4096, 1, 320, 180, 9),
                                                     (1 << 20, 2048, 161, 90, 4)])
def test_slot_pool_and_batches(rt, orc, house, env_small, monkeypatch, slots, staging_mb, w, h, spp):
    # Few slots: every slot traces many paths; a 1 MB staging budget: several batches
    # (1 MB / (w*h*16 B) samples each); sizes with partial 8x8 tiles: skipped path indices.
    monkeypatch.setenv("RT_KERNEL", "wave")
    monkeypatch.setenv("RT_WAVE_SLOTS", str(slots))
    monkeypatch.setenv("RT_WAVE_STAGING_MB", str(staging_mb))
    st, got = _render(rt, house, env_small, w, h, spp, 8, begin=3)
    want, counters = orc.Oracle(house, env_small, w, h, max_bounces=8).render(3, spp)
    assert got.tobytes() == want.tobytes(), _report(got, want)
    c = st.counters()
    assert (c["paths"], c["rays"], c["shadow_rays"]) == (counters["paths"], counters["rays"],
                                                         counters["shadow_rays"])
    assert (got[..., 3] == 1.0).all()


def test_walk_stack_in_shared_or_local_memory(rt, house, env_small, monkeypatch):
    # house.toml is 6 levels deep: its walk keeps the stack in shared memory unless told not to
    w, h = 160, 90
    monkeypatch.setenv("RT_KERNEL", "wave")
    monkeypatch.setenv("RT_WALK_SHALLOW", "1")
    _, shallow = _render(rt, house, env_small, w, h, 4, 8)
    monkeypatch.setenv("RT_WALK_SHALLOW", "0")
    _, deep = _render(rt, house, env_small, w, h, 4, 8)
    assert shallow.tobytes() == deep.tobytes()


def test_progressive_calls_equal_one_call(rt, house, env_small):
    w, h = 128, 72
    st, one = _render(rt, house, env_small, w, h, 7, 8)
    st.reset()
    for k in range(7):
        st.render_range(k, 1)
    assert st.read_accum().tobytes() == one.tobytes()


def test_ties_and_duplicates_scene(rt, orc, env_small):
    b = rt.SceneBuilder().add_material((0.8, 0.8, 0.8), 1.0, 0.0).add_material((0.9, 0.6, 0.3), 0.2, 1.0)
    for _ in range(3):
        b.add_plane((-4, -1, -4), (0, 0, 8), (8, 0, 0), 0)
    b.add_plane((-1, 1, -1), (0, 0, 2), (2, 0, 0), 1)      # on the cube's top face
    b.add_sphere((3, 0, 0), 1.0, 1).add_sphere((3, 0, 0), 1.0, 0)
    b.add_sphere((0, 0, 3), 0.5, 0)
    b.add_mesh_obj(_cube_at((0, 0, 0)), 0).add_mesh_obj(_cube_at((0, 0, 0)), 1)
    b.add_mesh_obj(_cube_at((-3, 0, 0), 0.5), 0)
    scene = b.set_camera((0.5, 1.5, 6), 0.0, -0.15, 1.2).build()
    w, h = 192, 108
    st, got = _render(rt, scene, env_small, w, h, 3, 6)
    ora = orc.Oracle(scene, env_small, w, h, max_bounces=6)
    want, _ = ora.render(0, 3)
    assert got.tobytes() == want.tobytes(), _report(got, want)
    slots, t = st.first_hit(0)
    oslots, ot, _ = ora.first_hit(0)
    assert np.array_equal(slots, oslots) and t.tobytes() == ot.tobytes()


def test_many_analytic_primitives_not_hoisted(rt, orc, env_small):
    # 46 spheres / planes: more than the pre-pass takes, so they stay in the BVH leaves; the
    # random planes are skewed parallelograms that stick out of their leaf boxes
    # (scene.rs:203-207 bounds two corners only), so the walk must not prune.
    rng = np.random.default_rng(5)
    b = rt.SceneBuilder().add_material((0.8, 0.8, 0.8), 1.0, 0.0).add_material((0.7, 0.7, 0.9), 0.1, 1.0)
    for _ in range(40):
        b.add_sphere(tuple(rng.uniform(-4, 4, 3)), float(rng.uniform(0.2, 0.8)), int(rng.integers(0, 2)))
    for _ in range(6):
        b.add_plane(tuple(rng.uniform(-4, 4, 3)), tuple(rng.uniform(-2, 2, 3)), tuple(rng.uniform(-2, 2, 3)), 0)
    b.add_mesh_obj(_cube_at((0, 0, 0)), 0)
    scene = b.set_camera((0, 1, 9), 0.0, 0.0, 1.2).build()
    w, h = 160, 90
    st, got = _render(rt, scene, env_small, w, h, 2, 5)
    ora = orc.Oracle(scene, env_small, w, h, max_bounces=5)
    want, _ = ora.render(0, 2)
    assert got.tobytes() == want.tobytes(), _report(got, want)
    slots, t = st.first_hit(0)
    oslots, ot, _ = ora.first_hit(0)
    assert np.array_equal(slots, oslots) and t.tobytes() == ot.tobytes()


def test_synthetic_field_deep_tree(rt, orc, env_small):
    # The config-D generator at a small grid: ~9k triangles over one ground plane.
    scene = rt.Scene.synthetic_field(ASSETS, 3)
    assert scene.counts()["triangles"] > 5000
    w, h = 256, 144
    st, got = _render(rt, scene, env_small, w, h, 1, 4)
    ora = orc.Oracle(scene, env_small, w, h, max_bounces=4)
    want, _ = ora.render(0, 1)
    assert got.tobytes() == want.tobytes(), _report(got, want)
    slots, t = st.first_hit(0)
    oslots, ot, _ = ora.first_hit(0)
    assert np.array_equal(slots, oslots) and t.tobytes() == ot.tobytes()
    assert (slots != 0xFFFFFFFF).mean() > 0.3


def test_profile_reports_every_kernel(rt, house, env_small, monkeypatch):
    monkeypatch.setenv("RT_KERNEL", "wave")  # per-kernel event timing of the launch-per-phase pipeline
    st = rt.State(house, env_small, 256, 144, device=0, max_bounces=8)
    st.set_profiling(True)
    st.render_samples(4)
    st.sync()
    prof = st.profile()
    st.set_profiling(False)
    assert prof["renders"] == 1 and prof["waves"] >= 8
    for name, k in prof["kernels"].items():
        assert k["launches"] == prof["waves"] and k["ms"] > 0, name
    assert sum(k["ms"] for k in prof["kernels"].values()) <= prof["total_ms"] * 1.05
    assert st.launch_count() >= 4 * prof["waves"]


def test_tiny_scenes_through_the_pipeline(rt, orc, env_small):
    # Scenes whose BVH root is a leaf, or whose effective tree is empty (spheres / planes only,
    # all hoisted): every ray is settled by the pre-pass kernel.
    builders = [lambda b: b.add_sphere((0, 0, 0), 1.0, 0),
                lambda b: b.add_plane((-1, 0, -1), (0, 0, 2), (2, 0, 0), 0).add_sphere((0, 1, 0), 0.5, 0),
                lambda b: b.add_mesh_raw([[0, 0, 0], [1, 0, 0], [0, 1, 0]], [[0, 0, 1]], [[0, 1, 2, 0, 0, 0]], 0)]
    for build in builders:
        b = rt.SceneBuilder().add_material((0.8, 0.7, 0.6), 0.4, 0.0)
        scene = build(b).set_camera((0.2, 0.4, 3), 0.0, 0.0, 1.0).build()
        w, h = 96, 64
        st, got = _render(rt, scene, env_small, w, h, 3, 6)
        ora = orc.Oracle(scene, env_small, w, h, max_bounces=6)
        want, _ = ora.render(0, 3)
        assert got.tobytes() == want.tobytes(), _report(got, want)
        slots, t = st.first_hit(0)
        oslots, ot, _ = ora.first_hit(0)
        assert np.array_equal(slots, oslots) and t.tobytes() == ot.tobytes()


def test_degenerate_big_leaf_on_gpu(rt, orc, env_small):
    # 40 triangles whose boxes share one centre cannot be split (bvh.rs:238-244): a leaf longer
    # than the inline length field of the N2 records, next to ordinary geometry.
    rng = np.random.default_rng(9)
    verts, tris = [], []
    for _ in range(40):
        a = np.round(rng.uniform(0.5, 2.0, size=3) * rng.choice([-1, 1], size=3) * 8) / 8
        base = len(verts)
        verts += [a, -a, a * np.array([0.5, -0.5, 0.25])]
        tris.append([base, base + 1, base + 2, 0, 0, 0])
    b = rt.SceneBuilder().add_material((0.8, 0.8, 0.8), 1.0, 0.0)
    b.add_mesh_raw(np.array(verts), [[0, 0, 1]], tris, 0).add_mesh_obj(_cube_at((4, 0, 0)), 0)
    b.add_sphere((-4, 0, 0), 1.0, 0)
    scene = b.set_camera((0, 0, 7), 0.0, 0.0, 1.0).build()
    assert scene.bvh_nodes_words()[:, 9].max() >= 31
    w, h = 160, 90
    st, got = _render(rt, scene, env_small, w, h, 2, 4)
    ora = orc.Oracle(scene, env_small, w, h, max_bounces=4)
    want, _ = ora.render(0, 2)
    assert got.tobytes() == want.tobytes(), _report(got, want)
    slots, t = st.first_hit(0)
    oslots, ot, _ = ora.first_hit(0)
    assert np.array_equal(slots, oslots) and t.tobytes() == ot.tobytes()


def _random_scene(rt, rng):
    """Random mix: 0-40 spheres, 0-8 planes (some skewed, some duplicated), cubes and loose
    triangles, rough and mirror-like materials, a random camera."""
    b = rt.SceneBuilder()
    n_mat = int(rng.integers(2, 6))
    for _ in range(n_mat):
        b.add_material(rng.uniform(0.1, 1.0, 3), float(rng.choice([0.0, 0.05, 0.3, 1.0])),
                       float(rng.integers(0, 2)), rng.uniform(0, 0.5, 3) * (rng.random() < 0.25))
    mat = lambda: int(rng.integers(0, n_mat))
    for _ in range(int(rng.choice([0, 1, 3, 10, 40]))):
        b.add_sphere(rng.uniform(-4, 4, 3), float(rng.uniform(0.15, 1.2)), mat())
    for _ in range(int(rng.choice([0, 1, 2, 8]))):
        pos = rng.uniform(-5, 5, 3)
        if rng.random() < 0.5:   # axis-aligned rectangle, like the reference's scenes
            f, r = (0, 0, float(rng.uniform(2, 10))), (float(rng.uniform(2, 10)), 0, 0)
        else:                     # skewed parallelogram: sticks out of its two-corner box
            f, r = rng.uniform(-3, 3, 3), rng.uniform(-3, 3, 3)
        m = mat()
        b.add_plane(pos, f, r, m)
        if rng.random() < 0.3:
            b.add_plane(pos, f, r, mat())  # the same plane twice (house.toml does this)
    for _ in range(int(rng.integers(0, 4))):
        b.add_mesh_obj(_cube_at(rng.uniform(-3, 3, 3), float(rng.uniform(0.3, 1.5))), mat())
    if rng.random() < 0.7:
        n = int(rng.integers(3, 60))
        verts = rng.uniform(-3, 3, (n, 3)).astype(np.float32)
        normals = rng.normal(size=(n, 3)).astype(np.float32)
        normals /= np.linalg.norm(normals, axis=1, keepdims=True)
        tris = np.array([[i, i + 1, i + 2, i, i + 1, i + 2] for i in range(n - 2)], dtype=np.uint32)
        b.add_mesh_raw(verts, normals, tris, mat())
    elif rng.random() < 0.5:
        b.add_sphere((0, 0, 0), 0.5, mat())  # make sure the scene is not empty
    else:
        b.add_mesh_obj(_cube_at((0, 0, 0)), mat())
    cam = (float(rng.uniform(-2, 2)), float(rng.uniform(-1, 3)), float(rng.uniform(4, 9)))
    return b.set_camera(cam, float(rng.uniform(-0.3, 0.3)), float(rng.uniform(-0.3, 0.2)),
                        float(rng.uniform(0.7, 1.6))).build()


@pytest.mark.parametrize("seed", range(int(os.environ.get("RT_FUZZ_SEEDS", "16"))))
def test_random_scene_fuzz(rt, orc, env_small, seed):
    rng = np.random.default_rng(1000 + seed)
    scene = _random_scene(rt, rng)
    w, h = 96, 64
    bounces = int(rng.integers(1, 9))
    st, got = _render(rt, scene, env_small, w, h, 2, bounces, begin=seed)
    ora = orc.Oracle(scene, env_small, w, h, max_bounces=bounces)
    want, counters = ora.render(seed, 2)
    assert got.tobytes() == want.tobytes(), (seed, scene.counts(), _report(got, want))
    c = st.counters()
    assert (c["rays"], c["shadow_rays"]) == (counters["rays"], counters["shadow_rays"])
    slots, t = st.first_hit(seed)
    oslots, ot, _ = ora.first_hit(seed)
    assert np.array_equal(slots, oslots) and t.tobytes() == ot.tobytes(), (seed, scene.counts())


def test_graph_loop_safety_bound_is_reported(rt, house, env_small, monkeypatch):
    """The graph's WHILE loop carries an upper bound on its bodies (far above what a batch needs); a
    loop cut off there makes rt_sync / rt_read_accum fail instead of returning a partial image."""
    st = rt.State(house, env_small, 320, 180, max_bounces=8)
    st.render_samples(4)
    st.sync()                                   # a normal render stays far below its bound
    monkeypatch.setenv("RT_WAVE_MAX_BODIES", "1")
    st.render_samples(4)
    with pytest.raises(rt.RtError, match="did not terminate"):
        st.sync()
    monkeypatch.delenv("RT_WAVE_MAX_BODIES")
    st.reset()
    st.render_samples(4)                        # the context keeps working
    st.sync()
    st.close()
