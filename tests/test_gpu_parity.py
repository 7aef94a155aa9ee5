"""GPU parity tests proper: the CUDA backend through the C ABI vs the CPU oracle.

Bars (DESIGN.md "Parity"):
  * first-hit slot maps and hit distances: bit-exact (integer/index work);
  * accumulation buffers: bit-exact per launch, because both sides implement the same
    arithmetic contract and a pixel's samples are summed in index order on both;
  * work counters (rays, shadow rays, node/primitive tests): exactly equal;
  * sample-split across contexts: max abs diff <= 1e-5 * spp relative to the
    single-context image (float re-association only);
  * tone-mapped 8-bit output: within 1 code value (powf in the sRGB encode).
"""
import os

import numpy as np
import pytest

from conftest import ASSETS

pytestmark = pytest.mark.gpu

MISS = 0xFFFFFFFF


def _state(rt, scene, env, w, h, **kw):
    return rt.State(scene, env, w, h, device=0, **kw)


def _report(name, a, b):
    diff = np.argwhere(a != b)
    return f"{name}: {len(diff)} mismatches, first at {diff[:5].tolist()}"


@pytest.mark.parametrize("name,w,h", [("house", 480, 270), ("default", 320, 180), ("suzanne", 512, 288)])
def test_first_hit_bit_exact(rt, orc, env_small, name, w, h):
    scene = rt.Scene.load_toml(os.path.join(ASSETS, "scenes", f"{name}.toml"))
    st = _state(rt, scene, env_small, w, h)
    ora = orc.Oracle(scene, env_small, w, h)
    for sample in (0, 7):
        slots, t = st.first_hit(sample)
        oslots, ot, _ = ora.first_hit(sample)
        assert np.array_equal(slots, oslots), _report("slots", slots, oslots)
        assert t.tobytes() == ot.tobytes(), _report("t", t, ot)
    assert (slots != MISS).mean() > 0.2  # the camera actually sees geometry


@pytest.mark.parametrize("name", ["house", "default"])
def test_single_sample_radiance_bit_exact(rt, orc, env_small, name):
    scene = rt.Scene.load_toml(os.path.join(ASSETS, "scenes", f"{name}.toml"))
    w, h = 256, 144
    st = _state(rt, scene, env_small, w, h, max_bounces=8)
    ora = orc.Oracle(scene, env_small, w, h, max_bounces=8)
    for sample in (0, 1, 12345):
        st.reset()
        st.render_range(sample, 1)
        got = st.read_accum()
        want, _ = ora.render(sample, 1)
        assert got.tobytes() == want.tobytes(), _report("accum", got, want)
    assert np.isfinite(got).all() and got[..., :3].max() > 0


def test_accumulation_order_bit_exact(rt, orc, house, env_small):
    # 16 samples in one launch, and 4 + 12 over two launches, equal the oracle's
    # sequential sum bit for bit (shader.wgsl:1367-1371 accumulates frame by frame).
    w, h = 192, 108
    ora = orc.Oracle(house, env_small, w, h, max_bounces=8)
    want, ocn = ora.render(0, 16)
    st = _state(rt, house, env_small, w, h, max_bounces=8)
    st.render_samples(16)
    one = st.read_accum()
    assert one.tobytes() == want.tobytes(), _report("accum16", one, want)
    st.reset()
    st.render_samples(4)
    st.render_samples(12)
    two = st.read_accum()
    assert two.tobytes() == want.tobytes()
    assert st.sample_count == 16
    assert (one[..., 3] == 1.0).all()


def test_counters_equal_oracle(rt, orc, house, env_small):
    w, h = 160, 90
    st = _state(rt, house, env_small, w, h, max_bounces=8)
    st.set_detailed_counters(True)
    st.reset_counters()
    st.render_samples(4)
    got = st.counters()
    # The backend prunes with the closest hit and stops shadow rays at the first hit;
    # the oracle's prune=True mode restates exactly that traversal.
    _, want = orc.Oracle(house, env_small, w, h, max_bounces=8, prune=True).render(0, 4)
    assert got == want
    # Rays / paths do not depend on the traversal variant.
    _, literal = orc.Oracle(house, env_small, w, h, max_bounces=8, prune=False).render(0, 4)
    for k in ("paths", "rays", "shadow_rays", "hits", "tri_hits", "shaded_bounces", "fallback_rays"):
        assert got[k] == literal[k]
    # Cheap counters (always on) agree with the detailed run.
    st.set_detailed_counters(False)
    st.reset()
    st.reset_counters()
    st.render_samples(4)
    light = st.counters()
    assert (light["paths"], light["rays"], light["shadow_rays"]) == (
        want["paths"], want["rays"], want["shadow_rays"])


def test_reference_bounce_count_and_mirror_materials(rt, orc, default_scene, env_small):
    # default.toml has roughness-0 metallic and dielectric mirrors (alpha clamps to 1e-3)
    # and emissive markers; run at the reference's MAX_BOUNCES = 10.
    w, h = 200, 112
    st = _state(rt, default_scene, env_small, w, h)
    assert st.max_bounces == 10
    st.render_samples(3)
    got = st.read_accum()
    want, _ = orc.Oracle(default_scene, env_small, w, h, max_bounces=10).render(0, 3)
    assert got.tobytes() == want.tobytes(), _report("accum", got, want)


def test_rotated_camera_and_ragged_resolution(rt, orc, house, env_small):
    # Width/height not multiples of the 16x8 tile; yaw/pitch != 0.
    w, h = 203, 117
    cam = rt.Camera((2.5, 1.6, 4.0), 0.6, -0.2, 1.1)
    st = _state(rt, house, env_small, w, h, max_bounces=6)
    st.update(cam)
    st.render_samples(2)
    got = st.read_accum()
    want, _ = orc.Oracle(house, env_small, w, h, camera=cam, max_bounces=6).render(0, 2)
    assert got.tobytes() == want.tobytes(), _report("accum", got, want)
    slots, t = st.first_hit(1)
    oslots, ot, _ = orc.Oracle(house, env_small, w, h, camera=cam).first_hit(1)
    assert np.array_equal(slots, oslots) and t.tobytes() == ot.tobytes()


def test_random_scenes_property(rt, orc, env_small):
    # Seeded random mixes of the three primitive types, rough and glossy materials.
    rng = np.random.default_rng(1234)
    for case in range(3):
        b = rt.SceneBuilder()
        for _ in range(4):
            b.add_material(rng.uniform(0.2, 1.0, 3), float(rng.uniform(0, 1)),
                           float(rng.integers(0, 2)), rng.uniform(0, 0.3, 3) * (rng.random() < 0.3))
        for _ in range(int(rng.integers(3, 12))):
            b.add_sphere(rng.uniform(-3, 3, 3), float(rng.uniform(0.2, 0.9)), int(rng.integers(0, 4)))
        b.add_plane((-6, -1.5, -6), (0, 0, 12), (12, 0, 0), int(rng.integers(0, 4)))
        b.add_plane((-6, -1.5, -6), (0, 6, 0), (12, 0, 0), int(rng.integers(0, 4)))
        verts = rng.uniform(-3, 3, (30, 3)).astype(np.float32)
        normals = rng.normal(size=(30, 3)).astype(np.float32)
        normals /= np.linalg.norm(normals, axis=1, keepdims=True)
        tris = np.array([[i, i + 1, i + 2, i, i + 1, i + 2] for i in range(0, 27, 1)], dtype=np.uint32)
        b.add_mesh_raw(verts, normals, tris, int(rng.integers(0, 4)))
        scene = b.set_camera((0.1, 0.3, 7.5), 0.05, -0.03, 1.0).build()
        w, h = 128, 96
        st = _state(rt, scene, env_small, w, h, max_bounces=5)
        st.render_range(case * 100, 2)
        got = st.read_accum()
        want, _ = orc.Oracle(scene, env_small, w, h, max_bounces=5).render(case * 100, 2)
        assert got.tobytes() == want.tobytes(), _report(f"case {case}", got, want)
        slots, _t = st.first_hit(0)
        oslots, _ot, _ = orc.Oracle(scene, env_small, w, h).first_hit(0)
        assert np.array_equal(slots, oslots)


def test_sample_split_matches_single_context(rt, house, env_small):
    # Multi-GPU partitioning run on one device: 4 disjoint sample ranges summed on the
    # host equal the 32-sample render up to float re-association.
    w, h = 160, 90
    spp = 32
    st = _state(rt, house, env_small, w, h, max_bounces=8)
    st.render_samples(spp)
    full = st.read_accum()
    total = np.zeros_like(full)
    for r in range(4):
        st.reset()
        st.render_range(r * 8, 8)
        total += st.read_accum()
    scale = np.maximum(1.0, np.abs(full[..., :3]))
    assert (np.abs(total[..., :3] - full[..., :3]) / scale).max() <= 1e-5 * spp


def test_tonemap_matches_restatement(rt, orc, house, env_small):
    w, h = 160, 90
    st = _state(rt, house, env_small, w, h, max_bounces=8)
    st.render_samples(8)
    got = st.tonemap_rgb8()
    want = orc.tonemap_rgb8(st.read_accum(), 8)
    assert np.abs(got.astype(int) - want.astype(int)).max() <= 1
    assert got.std() > 5  # a real image, not a constant


def test_resume_from_written_accumulation(rt, house, env_small):
    w, h = 96, 54
    st = _state(rt, house, env_small, w, h, max_bounces=4)
    st.render_samples(6)
    want = st.read_accum()
    st.reset()
    st.render_samples(2)
    part = st.read_accum()
    st2 = _state(rt, house, env_small, w, h, max_bounces=4)
    st2.write_accum(part)
    st2.render_range(2, 4)
    assert st2.read_accum().tobytes() == want.tobytes()


def test_error_behaviour(rt, house, env_small):
    st = _state(rt, house, env_small, 32, 32)
    with pytest.raises(rt.RtError) as e:
        st.environment_index = 3
        st.render()
    assert e.value.status == rt._ffi.RT_ERR_INVALID_ARGUMENT and "environment index" in str(e.value)
    st.environment_index = 0
    with pytest.raises(rt.RtError):
        st.resize(0, 10)
    with pytest.raises(rt.RtError) as e:
        rt.State(house, env_small, 16, 16, device=99)
    assert "device ordinal" in str(e.value)
    # max_bounces = 0: no rays, black image with alpha 1.
    st.max_bounces = 0
    st.resize(32, 32)
    st.reset_counters()
    st.render_samples(2)
    acc = st.read_accum()
    assert (acc[..., :3] == 0).all() and (acc[..., 3] == 1).all()
    assert st.counters()["rays"] == 0


def test_full_size_headline_properties(rt, orc, house, env_2k):
    # BASELINE.json configs[1] shape (house.toml, 1920x1080, 8 bounces) at full resolution:
    # first-hit map bit-exact vs the oracle, and size-independent properties of the render.
    w, h = 1920, 1080
    st = _state(rt, house, env_2k, w, h, max_bounces=8)
    slots, t = st.first_hit(0)
    oslots, ot, _ = orc.Oracle(house, env_2k, w, h).first_hit(0)
    assert np.array_equal(slots, oslots) and t.tobytes() == ot.tobytes()
    # Rows of the full-size image, sampled: bit-exact vs oracle on a 1920x8 strip is not
    # expressible (seeds depend on y*W+x), so check determinism + linearity instead.
    st.reset_counters()
    st.render_samples(4)
    a = st.read_accum()
    c1 = st.counters()
    st.reset()
    st.reset_counters()
    st.render_samples(4)
    b = st.read_accum()
    assert a.tobytes() == b.tobytes() and st.counters() == c1  # idempotent / deterministic
    assert c1["paths"] == w * h * 4 and c1["rays"] >= c1["paths"]
    assert c1["rays"] + c1["shadow_rays"] <= c1["paths"] * 16
    assert np.isfinite(a).all() and (a[..., :3] >= 0).all()
    # A handful of pixels against the oracle's per-pixel tracer at the full resolution.
    ora = orc.Oracle(house, env_2k, w, h, max_bounces=8)
    for (x, y) in [(0, 0), (959, 540), (1919, 1079), (700, 800), (1300, 300)]:
        acc = np.zeros(3, dtype=np.float32)
        for s in range(4):
            acc = acc + ora.trace_pixel(x, y, s)
        assert acc.tobytes() == a[y, x, :3].tobytes(), (x, y, acc, a[y, x])


def test_suzanne_4k_first_hit(rt, orc, suzanne, env_small):
    # BASELINE.json configs[2]: 3840x2160 primary rays, slot map bit-exact.
    w, h = 3840, 2160
    st = _state(rt, suzanne, env_small, w, h)
    slots, t = st.first_hit(0)
    oslots, ot, _ = orc.Oracle(suzanne, env_small, w, h).first_hit(0)
    assert np.array_equal(slots, oslots), _report("slots", slots, oslots)
    assert t.tobytes() == ot.tobytes()
    assert 0.3 < (slots != MISS).mean() < 0.95
