"""GPU parity on the workloads BASELINE.json names, at their own sizes, through the C ABI.

  B  house.toml 1920x1080: 16 spp of the whole frame bit-exact vs the oracle, and all 256 spp on
     >= 1000 strided pixels (oracle trace_pixel, summed in sample order);
  C  suzanne.toml 3840x2160: first-hit slot map + distances, whole frame, bit-exact;
  D  the ~10 M-triangle synthetic field at 1920x1080: first-hit maps and a 1-spp image bit-exact;
  E  house.toml 7680x4320: first-hit maps on a strided subset bit-exact, plus a 1-spp row band;
  +  an oracle fed from buffers the product's host library never touched (oracle/independent.py);
  +  rt_group_* (single-process multi-GPU) against one context, two progressive rounds;
  +  the memory budget knobs leave every bit of the image where it was.
"""
import ctypes
import os
import subprocess

import numpy as np
import pytest

from conftest import ASSETS, ROOT

pytestmark = pytest.mark.gpu
MISS = 0xFFFFFFFF


def _report(name, a, b):
    diff = np.argwhere(a != b)
    return f"{name}: {len(diff)} mismatches, first at {diff[:5].tolist()}"


def _device_count(rt):
    n = ctypes.c_int(0)
    rt._ffi.backend_lib().rt_device_count(ctypes.byref(n))
    return n.value


# ---------------------------------------------------------------- config B
def test_config_b_full_frame_16spp_bit_exact(rt, orc, house, env_2k):
    w, h = 1920, 1080
    st = rt.State(house, env_2k, w, h, max_bounces=8)
    st.reset_counters()
    st.render_samples(16)
    got = st.read_accum()
    want, oc = orc.Oracle(house, env_2k, w, h, max_bounces=8).render(0, 16)
    assert got.tobytes() == want.tobytes(), _report("accum", got, want)
    c = st.counters()
    assert (c["paths"], c["rays"], c["shadow_rays"]) == (oc["paths"], oc["rays"], oc["shadow_rays"])


def test_config_b_256spp_strided_pixels_bit_exact(rt, orc, house, env_2k):
    w, h, spp = 1920, 1080, 256
    st = rt.State(house, env_2k, w, h, max_bounces=8)
    st.render_samples(spp)
    got = st.read_accum()
    ora = orc.Oracle(house, env_2k, w, h, max_bounces=8)
    xs, ys = np.arange(7, w, 53), np.arange(5, h, 37)   # 37 x 30 = 1110 pixels
    assert len(xs) * len(ys) >= 1000
    bad = []
    for y in ys:
        for x in xs:
            acc = np.zeros(3, dtype=np.float32)
            for s in range(spp):
                acc = acc + ora.trace_pixel(int(x), int(y), s)   # shader.wgsl:1367: total += sample
            if acc.tobytes() != got[y, x, :3].tobytes():
                bad.append((int(x), int(y), acc.tolist(), got[y, x, :3].tolist()))
    assert not bad, f"{len(bad)} of {len(xs) * len(ys)} pixels differ, first {bad[:3]}"
    assert (got[..., 3] == 1.0).all()


# ---------------------------------------------------------------- config C
def test_config_c_suzanne_4k_first_hit_bit_exact(rt, orc, suzanne, env_small):
    w, h = 3840, 2160
    st = rt.State(suzanne, env_small, w, h)
    maps = (rt.pinned_array((h, w), np.uint32), rt.pinned_array((h, w), np.float32))
    slots, t = st.first_hit(0, out=maps)
    oslots, ot, _ = orc.Oracle(suzanne, env_small, w, h).first_hit(0)
    assert np.array_equal(slots, oslots), _report("slots", slots, oslots)
    assert t.tobytes() == ot.tobytes(), _report("t", t, ot)
    assert (slots != MISS).mean() > 0.2


# ---------------------------------------------------------------- config D
@pytest.fixture(scope="module")
def field99(rt):
    return rt.Scene.synthetic_field(ASSETS, 99)


def test_config_d_ten_million_triangles_first_hit_and_image(rt, orc, field99, env_small):
    assert field99.counts()["triangles"] == 9997020
    w, h = 1920, 1080
    st = rt.State(field99, env_small, w, h, max_bounces=8)
    ora = orc.Oracle(field99, env_small, w, h, max_bounces=8)
    slots, t = st.first_hit(0)
    oslots, ot, _ = ora.first_hit(0)
    assert np.array_equal(slots, oslots), _report("slots", slots, oslots)
    assert t.tobytes() == ot.tobytes(), _report("t", t, ot)
    assert (slots != MISS).mean() > 0.2
    st.reset_counters()
    st.render_samples(1)
    got = st.read_accum()
    want, oc = ora.render(0, 1)
    assert got.tobytes() == want.tobytes(), _report("accum", got, want)
    c = st.counters()
    assert (c["rays"], c["shadow_rays"]) == (oc["rays"], oc["shadow_rays"])


# ---------------------------------------------------------------- config E
def test_config_e_8k_frame_strided_first_hit_and_band(rt, orc, house, env_small):
    w, h = 7680, 4320
    st = rt.State(house, env_small, w, h, max_bounces=8)
    slots, t = st.first_hit(3)
    ora = orc.Oracle(house, env_small, w, h, max_bounces=8)
    xs, ys = np.arange(3, w, 29), np.arange(1, h, 31)
    origins, dirs, where = [], [], []
    for y in ys[::4]:              # camera_ray is one ctypes call per pixel: ~9 k pixels
        for x in xs[::4]:
            o, d, _ = ora.camera_ray(int(x), int(y), 3)
            origins.append(o)
            dirs.append(d)
            where.append((int(y), int(x)))
    oslots, ot = ora.cast_rays(np.array(origins), np.array(dirs))
    yy, xx = np.array(where).T
    assert np.array_equal(slots[yy, xx], oslots)
    assert np.where(oslots == MISS, np.float32(0), ot).tobytes() == t[yy, xx].tobytes()
    # and a 1-spp image of the whole 8K frame against the oracle on a band of rows
    st.render_samples(1)
    got = st.read_accum()
    for y in (0, 2159, 4319):
        for x in range(0, w, 97):
            assert ora.trace_pixel(x, y, 0).tobytes() == got[y, x, :3].tobytes(), (x, y)


# ---------------------------------------------------------------- independent front-end
@pytest.mark.parametrize("name", ["house", "default"])
def test_image_vs_oracle_fed_by_the_numpy_front_end(rt, orc, name):
    import independent
    path = os.path.join(ASSETS, "scenes", f"{name}.toml")
    sky = rt.Environment.synthetic(64, 32)
    iscene = independent.IndependentScene(path)
    ienv = independent.IndependentEnvironments([sky.rgb])
    w, h = 256, 144
    want, oc = orc.Oracle(iscene, ienv, w, h, max_bounces=8).render(0, 4)
    oslots, ot, _ = orc.Oracle(iscene, ienv, w, h).first_hit(0)

    scene = rt.Scene.load_toml(path)
    st = rt.State(scene, rt.EnvironmentMaps([sky]), w, h, max_bounces=8)
    slots, t = st.first_hit(0)
    assert np.array_equal(slots, oslots) and t.tobytes() == ot.tobytes()
    st.reset_counters()
    st.render_samples(4)
    got = st.read_accum()
    assert got.tobytes() == want.tobytes(), _report("accum", got, want)
    c = st.counters()
    assert (c["rays"], c["shadow_rays"]) == (oc["rays"], oc["shadow_rays"])


# ---------------------------------------------------------------- rt_group_*
def _group_devices(rt):
    n = _device_count(rt)
    return list(range(min(n, 4)))


def test_group_two_progressive_rounds_match_one_context(rt, house, env_small):
    devices = _group_devices(rt)
    w, h = 320, 180
    one = rt.State(house, env_small, w, h, max_bounces=8)
    one.render_samples(8)
    one.render_samples(12)
    want = one.read_accum()

    grp = rt.StateGroup(house, env_small, w, h, devices, max_bounces=8)
    assert grp.size == len(devices)
    grp.render_samples(8)
    grp.reduce(0)
    grp.render_samples(12)     # progressive: a second round on top of the first
    grp.reduce(0)
    grp.sync()
    got = grp.read_accum(0)
    assert (got[..., 3] == 1.0).all()
    tol = 1e-5 * 20 * max(1.0, float(want[..., :3].max()))
    assert np.abs(got[..., :3] - want[..., :3]).max() <= tol
    if len(devices) == 1:
        assert got.tobytes() == want.tobytes()
    else:
        for r in range(1, grp.size):           # the other members are zero again
            assert not grp.read_accum(r).any()
        rays = sum(grp.counters(r)["rays"] for r in range(grp.size))
        assert rays == one.counters()["rays"]
    grp.close()


def test_group_needs_two_devices_for_nccl(rt, house, env_small):
    if _device_count(rt) < 2:
        pytest.skip("one device: the NCCL reduce of rt_group_reduce needs two")
    w, h = 480, 270
    one = rt.State(house, env_small, w, h, max_bounces=8)
    one.render_samples(16)
    want = one.read_accum()
    grp = rt.StateGroup(house, env_small, w, h, [0, 1], max_bounces=8)
    grp.render_samples(16)
    grp.reduce(0)
    grp.sync()
    got = grp.read_accum(0)
    assert (got[..., 3] == 1.0).all()
    assert np.abs(got[..., :3] - want[..., :3]).max() <= 1e-5 * 16 * max(1.0, float(want[..., :3].max()))
    assert grp.counters(0)["paths"] == w * h * 8 and grp.counters(1)["paths"] == w * h * 8
    grp.close()


def test_cli_multi_gpu_render_matches_api(rt, house, env_small, tmp_path):
    if _device_count(rt) < 2:
        pytest.skip("one device: rt_render --gpus 2 needs two")
    exe = os.path.join(ROOT, "rsoderh-raytracing_b200", "rt_render")
    pfm, png = tmp_path / "g2.pfm", tmp_path / "g2.png"
    cmd = [exe, f"--scene={os.path.join(ASSETS, 'scenes', 'house.toml')}", "--width", "320", "--height", "180",
           "--spp", "8", "--bounces", "8", "--gpus", "2", "--out", str(png), "--pfm", str(pfm)]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr
    assert pfm.exists() and pfm.stat().st_size > 320 * 180 * 12


# ---------------------------------------------------------------- memory budget
def test_memory_budget_does_not_change_the_image(rt, house, env_small):
    w, h = 640, 360
    a = rt.State(house, env_small, w, h, max_bounces=8)
    a.render_samples(24)
    want = a.read_accum()
    b = rt.State(house, env_small, w, h, max_bounces=8)
    b.set_memory_budget(staging_bytes=8 << 20, path_slots=64 * 1024)   # 2 samples per batch, small pool
    b.render_samples(24)
    assert b.read_accum().tobytes() == want.tobytes()
    assert b.memory_usage() < a.memory_usage()


# ---------------------------------------------------------------- dev_index views, environment slots
def _two_skies(rt):
    a = rt.Environment.synthetic(64, 32)
    rgb = np.ascontiguousarray(a.rgb[::-1, ::-1] * np.float32(0.5) + np.float32(0.125))  # a different sky
    return rt.EnvironmentMaps([a, rt.Environment(rgb)])


def test_debug_view_environment_texels(rt, house):
    # dev_index 3 (shader.wgsl:1332-1338): saturate(environment texel at the pixel), f16 texture
    envs = _two_skies(rt)
    w, h = 96, 54
    st = rt.State(house, envs, w, h, environment_index=1)
    got = st.debug_view(3)
    rgb = envs.environments[1].rgb
    want = np.zeros((h, w, 4), dtype=np.float32)
    want[:32, :64, :3] = np.clip(rgb, 0.0, 1.0).astype(np.float16).astype(np.float32)
    assert got.tobytes() == want.tobytes()
    with pytest.raises(rt.RtError):
        st.debug_view(1)   # not a view: dev_index 1 renders normally in the reference


def test_debug_view_sampling_distribution(rt, orc, house):
    # dev_index 2 (shader.wgsl:1314-1331) against a numpy walk over the same alias table and RNG
    import ctypes as C
    envs = _two_skies(rt)
    w, h, sample = 48, 27, 5
    st = rt.State(house, envs, w, h)
    got = st.debug_view(2, sample_index=sample)
    env = envs.environments[0]
    prob = env.alias[:, 0].view(np.float32)
    alias = env.alias[:, 1]
    length = env.width * env.height
    lib = orc.lib()
    counts = np.zeros((h, w), dtype=np.int64)
    state = C.c_uint32()
    for y in range(h):
        for x in range(w):
            lib.oracle_rng_seed(y * w + x, sample, C.byref(state))
            for _ in range(20):
                index = min(int(np.float32(lib.oracle_rng_uniform(C.byref(state))) * np.float32(length)), length - 1)
                pick = index if lib.oracle_rng_uniform(C.byref(state)) < prob[index] else int(alias[index])
                ex, ey = pick % env.width, pick // env.width
                if ex < w and ey < h:
                    counts[ey, ex] += 1
    assert counts.sum() > 0
    want = np.zeros((h, w, 4), dtype=np.float32)
    step = np.float32(0.1) / np.float32(20.0)
    for (y, x), k in np.ndenumerate(counts):
        c = np.float32(0.0)
        for _ in range(int(k)):
            nxt = np.float32(np.float16(c + step))
            if nxt == c:
                break
            c = nxt
        want[y, x, :3] = c
    assert got.tobytes() == want.tobytes()


def test_environment_slot_cycling(rt, orc, house):
    # the next_environment key (camera.rs:271-278): slot 0 -> 1 -> back to 0, accumulation restarts
    envs = _two_skies(rt)
    w, h = 128, 72
    st = rt.State(house, envs, w, h, max_bounces=8)
    images = []
    for expect_slot in (0, 1, 0):
        assert st.environment_index == expect_slot and st.sample_count == 0
        st.render_samples(3)
        got = st.read_accum()
        want, _ = orc.Oracle(house, envs, w, h, max_bounces=8, environment_index=expect_slot).render(0, 3)
        assert got.tobytes() == want.tobytes()
        images.append(got)
        st.next_environment()
    assert images[0].tobytes() == images[2].tobytes() != images[1].tobytes()
