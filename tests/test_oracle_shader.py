"""The CPU oracle itself (oracle/shader_oracle.cpp): known answers and invariants — CPU only.

The reference ships no golden vectors (PARITY UNPINNED, see the oracle header), so the
pins here are: the PCG hash against an independent big-int restatement of
shader.wgsl:605-623, analytic hit distances on hand-built scenes, and self-consistency
(literal vs pruned traversal, thread-count independence, frame-by-frame accumulation).
"""
import ctypes as C
import math
import os

import numpy as np
import pytest

from conftest import ASSETS

MISS = 0xFFFFFFFF


def _pcg_py(state):
    state = (state * 747796405 + 2891336453) & 0xFFFFFFFF
    r = (((state >> ((state >> 28) + 4)) ^ state) * 277803737) & 0xFFFFFFFF
    return state, ((r >> 22) ^ r) & 0xFFFFFFFF


def test_rng_known_answers(orc):
    lib = orc.lib()
    for pixel, sample in [(0, 0), (1, 0), (1919 + 1079 * 1920, 255), (12345, 4294967295)]:
        s = 0
        s ^= pixel
        s, _ = _pcg_py(s)
        s ^= sample
        s, _ = _pcg_py(s)
        st = C.c_uint32()
        lib.oracle_rng_seed(pixel, sample, C.byref(st))
        assert st.value == s
        for _ in range(5):
            s, out = _pcg_py(s)
            st2 = C.c_uint32(st.value)
            assert lib.oracle_rng_next(C.byref(st)) == out
            # f32(u32) / 2^32 with round-to-nearest: can be exactly 1.0.
            assert lib.oracle_rng_uniform(C.byref(st2)) == float(np.float32(out) / np.float32(4294967296.0))
    # A state whose output rounds up to 2^32: uniform returns exactly 1.0.
    assert float(np.float32(0xFFFFFFFF) / np.float32(4294967296.0)) == 1.0


def test_transcendentals_accuracy_and_edges(orc):
    lib = orc.lib()
    xs = np.linspace(-7.0, 7.0, 4001).astype(np.float32)
    assert max(abs(lib.oracle_sin(float(x)) - math.sin(float(x))) for x in xs) < 2.5e-7
    assert max(abs(lib.oracle_cos(float(x)) - math.cos(float(x))) for x in xs) < 2.5e-7
    ys = np.linspace(-1.0, 1.0, 2001).astype(np.float32)
    assert max(abs(lib.oracle_asin(float(y)) - math.asin(float(y))) for y in ys) < 4e-7
    rng = np.random.default_rng(3)
    for y, x in rng.standard_normal((2000, 2)).astype(np.float32):
        assert abs(lib.oracle_atan2(float(y), float(x)) - math.atan2(float(y), float(x))) < 6e-7
    assert lib.oracle_sin(0.0) == 0.0 and lib.oracle_cos(0.0) == 1.0
    assert lib.oracle_atan2(0.0, 0.0) == 0.0
    assert lib.oracle_atan2(1.0, 0.0) == float(np.float32(math.pi / 2))
    assert lib.oracle_asin(1.0000001) == lib.oracle_asin(1.0)  # clamped, not NaN


def _one_sphere_scene(rt):
    return (rt.SceneBuilder().add_material((0.8, 0.8, 0.8), 1.0, 0.0)
            .add_sphere((0, 0, -5), 1.0, 0).set_camera((0, 0, 0), 0, 0, math.radians(60)).build())


def test_analytic_first_hit(rt, orc, env_small):
    scene = _one_sphere_scene(rt)
    w, h = 65, 65
    ora = orc.Oracle(scene, env_small, w, h)
    slots, t, counters = ora.first_hit(0)
    assert slots[32, 32] == 0 and abs(t[32, 32] - 4.0) < 2e-2   # jitter < 1 px
    assert slots[0, 0] == MISS and t[0, 0] == 0.0
    assert counters["paths"] == w * h == counters["rays"]
    # Camera ray of the centre pixel points down -z; seed state is reproducible.
    o, d, state = ora.camera_ray(32, 32, 0)
    assert o.tolist() == [0, 0, 0] and d[2] < -0.99 and abs(float(np.linalg.norm(d)) - 1) < 1e-6
    o2, d2, state2 = ora.camera_ray(32, 32, 0)
    assert state == state2 and d.tobytes() == d2.tobytes()


def test_plane_and_triangle_hits(rt, orc, env_small):
    tri = "v -1 -1 -3\nv 1 -1 -3\nv 0 1 -3\nvn 0 0 1\nf 1//1 2//1 3//1\n"
    scene = (rt.SceneBuilder().add_material((1, 1, 1), 1, 0)
             .add_plane((-10, -2, -10), (0, 0, 20), (20, 0, 0), 0)
             .add_mesh_obj(tri, 0).set_camera((0, 0, 0), 0, 0, math.radians(90)).build())
    w, h = 33, 33
    slots, t, _ = orc.Oracle(scene, env_small, w, h).first_hit(0)
    prims = scene.primitives_words().tolist()
    centre = prims[slots[16, 16]]
    assert centre == [2, 0] and abs(t[16, 16] - 3.0) < 0.1
    below = prims[slots[30, 16]]
    assert below == [1, 0]            # looking down: the ground plane at y = -2
    assert slots[1, 16] == MISS       # looking up: sky


@pytest.mark.parametrize("name", ["house", "default"])
def test_pruned_traversal_equals_literal_shader(rt, orc, env_small, name):
    scene = rt.Scene.load_toml(os.path.join(ASSETS, "scenes", f"{name}.toml"))
    w, h = 96, 54
    lit = orc.Oracle(scene, env_small, w, h, max_bounces=8, prune=False)
    pru = orc.Oracle(scene, env_small, w, h, max_bounces=8, prune=True)
    a, ca = lit.render(0, 3)
    b, cb = pru.render(0, 3)
    assert a.tobytes() == b.tobytes()
    for k in ("paths", "rays", "shadow_rays", "hits", "tri_hits", "shaded_bounces", "fallback_rays"):
        assert ca[k] == cb[k]
    assert cb["nodes_tested"] <= ca["nodes_tested"]
    s1, t1, _ = lit.first_hit(2)
    s2, t2, _ = pru.first_hit(2)
    assert np.array_equal(s1, s2) and t1.tobytes() == t2.tobytes()


def test_oracle_is_thread_count_independent_and_accumulates_per_frame(orc, house, env_small):
    w, h = 64, 36
    ora = orc.Oracle(house, env_small, w, h, max_bounces=8)
    a, ca = ora.render(0, 4, threads=1)
    b, cb = ora.render(0, 4, threads=5)
    assert a.tobytes() == b.tobytes() and ca == cb
    # Frame-by-frame accumulation (shader.wgsl:1367-1371): 1 + 3 samples == 4 samples.
    c, _ = ora.render(0, 1)
    c, _ = ora.render(1, 3, accum=c)
    assert c.tobytes() == a.tobytes()
    assert (a[..., 3] == 1).all() and np.isfinite(a).all() and (a[..., :3] >= 0).all()
    # Per-pixel tracer agrees with the image path.
    acc = np.zeros(3, np.float32)
    for s in range(4):
        acc = acc + ora.trace_pixel(40, 20, s)
    assert acc.tobytes() == a[20, 40, :3].tobytes()


def test_ray_budget_bounds(orc, house, env_small):
    # <= 1 closest-hit + 1 shadow ray per bounce.
    w, h, spp, bounces = 48, 27, 4, 8
    _, c = orc.Oracle(house, env_small, w, h, max_bounces=bounces).render(0, spp)
    assert c["paths"] == w * h * spp
    assert c["paths"] <= c["rays"] <= c["paths"] * bounces
    assert c["shadow_rays"] <= c["hits"] == c["shaded_bounces"]
    assert c["env_lookups"] == c["rays"]  # one bilinear fetch per miss + one per NEE sample
    _, c1 = orc.Oracle(house, env_small, w, h, max_bounces=1).render(0, spp)
    assert c1["rays"] == c1["paths"]


def test_sky_only_scene_matches_closed_form(rt, orc, env_small):
    # Camera looking away from a far, tiny sphere: every path escapes at bounce 0 and the
    # pixel equals sky * power_heuristic(1, pdf) (shader.wgsl:1222-1231, last_sample_pdf = 1).
    scene = (rt.SceneBuilder().add_material((1, 1, 1), 1, 0).add_sphere((0, -500, 500), 0.01, 0)
             .set_camera((0, 0, 0), 0, 0.4, math.radians(50)).build())
    w, h = 16, 9
    ora = orc.Oracle(scene, env_small, w, h, max_bounces=8)
    acc, c = ora.render(0, 1)
    assert c["rays"] == c["paths"] == w * h and c["hits"] == 0 and c["fallback_rays"] == w * h
    env = env_small.environments[0]
    f32 = np.float32
    for (x, y) in [(0, 0), (8, 4), (15, 8)]:
        _, d, _ = ora.camera_ray(x, y, 0)
        lib = orc.lib()
        inv_pi = f32(1.0 / 3.14159)
        u = f32(lib.oracle_atan2(float(d[2]), float(d[0]))) * inv_pi * f32(0.5) + f32(0.5)
        v = f32(0.5) - f32(lib.oracle_asin(float(d[1]))) * inv_pi
        # bilinear fetch, clamp-to-edge
        fx, fy = u * f32(env.width) - f32(0.5), v * f32(env.height) - f32(0.5)
        x0, y0 = np.floor(fx), np.floor(fy)
        wx, wy = f32(fx - x0), f32(fy - y0)
        cl = lambda a, n: int(min(max(a, 0), n - 1))
        tex = lambda xx, yy: env.rgb[cl(yy, env.height), cl(xx, env.width)]
        top = tex(x0, y0) * (f32(1) - wx) + tex(x0 + 1, y0) * wx
        bot = tex(x0, y0 + 1) * (f32(1) - wx) + tex(x0 + 1, y0 + 1) * wx
        sky = top * (f32(1) - wy) + bot * wy
        ix = min(int(u * f32(env.width)), env.width - 1)
        iy = min(int(v * f32(env.height)), env.height - 1)
        pmf = env.alias[ix + iy * env.width, 2:3].view(np.float32)[0]
        pi = f32(3.14159)
        sin_t = max(f32(1e-6), f32(lib.oracle_sin(float(pi * v))))
        pdf = pmf / (f32(2.0 * 3.14159) / f32(env.width) * (pi / f32(env.height)) * sin_t)
        weight = f32(1.0) / (f32(1.0) + pdf * pdf)
        want = sky * weight
        assert np.allclose(acc[y, x, :3], want, rtol=1e-6, atol=1e-7), (x, y, acc[y, x], want)


def test_white_furnace_energy(rt, orc):
    # A constant environment of radiance c around one big sphere. Background pixels have a closed
    # form: c / (1 + pdf^2) with the uniform-sphere pdf 1 / (4 pi') that the luminance x sin(theta)
    # alias table yields for a constant map (pi' = the shader's 3.14159). Pixels on the sphere
    # estimate the hemispherical albedo of the Lambert + GGX BSDF times c: with MIS and next-event
    # estimation in place they must not gain energy (<= c within noise), and a white rough
    # dielectric keeps most of it. A wrong MIS weight, pdf or cosine shows up here at once.
    c = np.float32(0.75)
    sky = rt.Environment(np.full((32, 64, 3), c, dtype=np.float32))
    envs = rt.EnvironmentMaps([sky])
    scene = (rt.SceneBuilder().add_material((1, 1, 1), 1, 0).add_sphere((0, 0, -3), 1.0, 0)
             .set_camera((0, 0, 0), 0, 0, math.radians(60)).build())
    w, h, spp = 64, 36, 96
    ora = orc.Oracle(scene, envs, w, h, max_bounces=10)
    acc, counters = ora.render(0, spp)
    img = acc[..., :3] / np.float32(spp)
    slots, _, _ = ora.first_hit(0)
    on_sphere = slots != MISS
    assert 0.1 < on_sphere.mean() < 0.6
    pdf = 1.0 / (4.0 * 3.14159)
    assert np.allclose(img[0, 0], c / (1.0 + pdf * pdf), rtol=2e-3)     # corner pixel: always background
    # the jittered camera rays of edge pixels mix both, so look at the disc's interior only
    ys, xs = np.nonzero(on_sphere)
    cy, cx = ys.mean(), xs.mean()
    r = np.sqrt(on_sphere.sum() / np.pi)
    inner = on_sphere & (((np.arange(h)[:, None] - cy) ** 2 + (np.arange(w)[None, :] - cx) ** 2) < (0.8 * r) ** 2)
    mean = float(img[inner].mean())
    assert 0.80 * c < mean <= 1.005 * c, mean
    assert float(img[inner].max()) < 1.25 * c      # no fireflies above the furnace level
    assert counters["shadow_rays"] > 0 and counters["env_lookups"] > 0


@pytest.mark.parametrize("name,w,h,bounces", [("house", 96, 54, 8), ("default", 80, 45, 10), ("suzanne", 64, 36, 8)])
def test_third_opinion_on_trace_ray(rt, orc, name, w, h, bounces):
    # oracle/third_opinion.py restates main / trace_ray of shader.wgsl once more (numpy float32
    # scalars, written from the WGSL, fed by the numpy front-end). 200 random pixel-samples per scene
    # must equal the C++ oracle's bit for bit: a slip in either transcription would show here.
    import independent
    import third_opinion
    path = os.path.join(ASSETS, "scenes", f"{name}.toml")
    sky = rt.Environment.synthetic(64, 32)
    world = third_opinion.World(path, sky.rgb, w, h, bounces)
    ora = orc.Oracle(independent.IndependentScene(path), independent.IndependentEnvironments([sky.rgb]), w, h,
                     max_bounces=bounces)
    rng = np.random.default_rng(2026)
    values = set()
    for _ in range(200):
        x, y, sample = int(rng.integers(0, w)), int(rng.integers(0, h)), int(rng.integers(0, 1 << 20))
        mine, theirs = world.pixel_sample(x, y, sample), ora.trace_pixel(x, y, sample)
        assert mine.tobytes() == theirs.tobytes(), (name, x, y, sample, mine, theirs)
        values.add(mine.tobytes())
    assert len(values) > 150  # the samples really differ from each other


def test_oracle_command_line(rt, orc, tmp_path):
    # oracle/cli.py writes what the parity checks compare; its outputs equal the API's
    import json
    import subprocess
    import sys
    from conftest import ROOT
    scene = os.path.join(ROOT, "assets", "scenes", "default.toml")
    out = {k: str(tmp_path / k) for k in ("a.pfm", "hit.u32", "bvh.bin", "alias.bin", "c.json")}
    subprocess.run([sys.executable, os.path.join(ROOT, "oracle", "cli.py"), "--scene", scene, "--width", "40",
                    "--height", "24", "--spp", "2", "--bounces", "3", "--pfm", out["a.pfm"], "--first-hit",
                    out["hit.u32"], "--bvh", out["bvh.bin"], "--alias", out["alias.bin"], "--counters",
                    out["c.json"]], check=True, cwd=ROOT)
    sc = rt.Scene.load_toml(scene)
    env = rt.EnvironmentMaps([rt.Environment.synthetic(2048, 1024)])
    o = orc.Oracle(sc, env, 40, 24, max_bounces=3)
    accum, counters = o.render(0, 2)
    slots, _t, _c = o.first_hit(0)
    raw = open(out["a.pfm"], "rb").read()
    data = np.frombuffer(raw[len(b"PF\n40 24\n-1.0\n"):], dtype="<f4").reshape(24, 40, 3)
    assert np.array_equal(data, accum[::-1, :, :3])
    assert np.array_equal(np.fromfile(out["hit.u32"], dtype=np.uint32).reshape(24, 40), slots)
    c = sc.counts()
    assert os.path.getsize(out["bvh.bin"]) == 48 * c["bvh_nodes"] + 8 * c["primitives"]
    assert os.path.getsize(out["alias.bin"]) == 16 * 2048 * 1024
    report = json.load(open(out["c.json"]))
    assert report["render_counters"]["rays"] == counters["rays"] and report["counts"] == c


def test_fan_anchor_report_is_current():
    # wavefront_obj's fan anchoring is restated from memory (last corner). The committed report
    # says what would move under the other reading (first corner): never the BVH topology or the
    # primitive order, only which triangle of a split quad a ray lands in and last-bit distances.
    import subprocess
    import sys
    from conftest import ROOT
    import json
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "fan_anchor_report.py"), "--check"],
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, "tests/golden/fan_anchor_report.json is stale: rerun tools/fan_anchor_report.py"
    with open(os.path.join(ROOT, "tests", "golden", "fan_anchor_report.json")) as fh:
        rep = json.load(fh)
    for name, s in rep["scenes"].items():
        assert s["bvh_topology_identical"] and s["bvh_bounds_identical"] and s["primitive_order_identical"], name
        assert s["first_hit_t_max_rel_diff"] < 1e-5 and s["image_rel_rmse"] < 0.01, name
