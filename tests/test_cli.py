"""The headless command line renderer (rsoderh-raytracing_b200/host/cli_main.cpp): the reference's
flags, messages and exit codes (src/cli.rs:10-46, camera.rs:139-181) on the CPU; a render
through it against the Python API on the GPU."""
import os
import subprocess

import numpy as np
import pytest

from conftest import ASSETS, ROOT

CLI = os.path.join(ROOT, "rsoderh-raytracing_b200", "rt_render")
HOUSE = os.path.join(ASSETS, "scenes", "house.toml")


def _run(*args):
    return subprocess.run([CLI, *args], capture_output=True, text=True, cwd=ROOT)


@pytest.fixture(scope="module", autouse=True)
def _built():
    if not os.path.exists(CLI):
        subprocess.run(["make", "-s"], cwd=ROOT, check=True)


def test_usage_and_required_scene():
    r = _run()
    assert r.returncode == 2 and "--scene <SCENE>" in r.stderr  # clap's missing-argument exit code
    r = _run("--help")
    assert r.returncode == 2 and "Usage: rt_render" in r.stderr
    r = _run("--scene")
    assert r.returncode == 2 and "a value is required" in r.stderr
    r = _run("--scene", HOUSE, "--bogus")
    assert r.returncode == 2 and "unexpected argument '--bogus'" in r.stderr


def test_keyboard_config_errors_keep_the_reference_messages():
    r = _run("--scene", HOUSE, "--movement-keys", "wasd")
    assert r.returncode == 2
    assert "Invalid keyboard config: Invalid keyboard config 'wasd': expected 6 characters." in r.stderr
    r = _run("--scene", HOUSE, "--other-keys=cp")
    assert r.returncode == 2
    assert "Invalid mouse capture config 'cp': expected 3 character." in r.stderr


def test_scene_errors_are_reported_not_crashed(tmp_path):
    r = _run("--scene", str(tmp_path / "missing.toml"), "--width", "8", "--height", "8", "--spp", "1")
    assert r.returncode == 1 and r.stderr.startswith("Error: ")
    bad = tmp_path / "bad.toml"
    bad.write_text("[[objects]]\ntype = \"Sphere\"\n")
    r = _run("--scene", str(bad), "--width", "8", "--height", "8", "--spp", "1")
    assert r.returncode == 1 and r.stderr.startswith("Error: ")
    r = _run("--scene", HOUSE, "--state", "not base64!", "--width", "8", "--height", "8", "--spp", "1")
    assert r.returncode == 1 and r.stderr.startswith("Error: ")
    r = _run("--scene", HOUSE, "--width", "0")
    assert r.returncode == 2


@pytest.mark.gpu
def test_render_to_files_equals_the_python_api(rt, tmp_path):
    w, h, spp, bounces = 96, 54, 3, 6
    pfm, png, hits = (str(tmp_path / n) for n in ("a.pfm", "a.png", "hit.u32"))
    r = _run("--scene", os.path.join(ASSETS, "scenes", "default.toml"), "--scene", HOUSE, "--width", str(w),
             "--height", str(h), "--spp", str(spp), "--bounces", str(bounces), "--pfm", pfm, "--out", png,
             "--first-hit-out", hits)
    assert r.returncode == 0, r.stderr
    scene = rt.Scene.load_toml(HOUSE)  # the last --scene wins (cli.rs:37)
    env = rt.EnvironmentMaps([rt.Environment.synthetic(2048, 1024)])  # what the CLI uses without --hdri
    st = rt.State(scene, env, w, h, max_bounces=bounces)
    st.render_samples(spp)
    accum = st.read_accum()
    raw = open(pfm, "rb").read()
    header = f"PF\n{w} {h}\n-1.0\n".encode()
    assert raw.startswith(header)
    data = np.frombuffer(raw[len(header):], dtype="<f4").reshape(h, w, 3)
    assert np.array_equal(data, accum[::-1, :, :3])
    slots, _t = st.first_hit(0)
    assert np.array_equal(np.fromfile(hits, dtype=np.uint32).reshape(h, w), slots)
    assert open(png, "rb").read(8) == b"\x89PNG\r\n\x1a\n"


@pytest.mark.gpu
def test_device_built_bvh_renders_the_same_file(tmp_path):
    """--bvh device: the BVH rebuilt by rt_bvh_build is the host builder's, so the files are equal."""
    outs = []
    for mode in ("host", "device"):
        pfm = str(tmp_path / f"{mode}.pfm")
        r = _run("--scene", os.path.join(ASSETS, "scenes", "suzanne.toml"), "--width", "80", "--height", "45", "--spp", "2",
                 "--bounces", "4", "--pfm", pfm, "--bvh", mode)
        assert r.returncode == 0, r.stderr
        assert ("BVH built on the device" in r.stderr) == (mode == "device")
        outs.append(open(pfm, "rb").read())
    assert outs[0] == outs[1]
    assert _run("--scene", HOUSE, "--bvh", "gpu").returncode == 2


def _build_c_example(tmp_path):
    exe = str(tmp_path / "render_minimal")
    pkg = os.path.join(ROOT, "rsoderh-raytracing_b200")
    r = subprocess.run(["gcc", "-std=c99", "-Wall", "-Wextra", "-pedantic", "-Werror", "-I" + os.path.join(ROOT, "include"),
                        os.path.join(ROOT, "examples", "render_minimal.c"), "-L" + pkg, "-lrt_host", "-lrt_backend",
                        "-Wl,-rpath," + pkg, "-o", exe], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    return exe


def test_headers_are_plain_c_and_the_c_example_links(tmp_path, cuda_available):
    """include/*.h compile as pedantic C99 and a plain C host links against the two libraries; without
    a GPU the program stops at rt_create with the backend's message (no CPU path)."""
    exe = _build_c_example(tmp_path)
    if not cuda_available:
        r = subprocess.run([exe, HOUSE, str(tmp_path / "o.png"), "64", "36", "2"], capture_output=True, text=True)
        assert r.returncode == 1 and "no CPU path" in r.stderr


@pytest.mark.gpu
def test_c_example_renders(tmp_path):
    exe = _build_c_example(tmp_path)
    out = str(tmp_path / "o.png")
    r = subprocess.run([exe, HOUSE, out, "96", "54", "3"], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    assert "96x54, 3 spp" in r.stdout and open(out, "rb").read(8) == b"\x89PNG\r\n\x1a\n"
