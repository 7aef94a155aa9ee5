import os
import shutil
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle")):
    if p not in sys.path:
        sys.path.insert(0, p)

import _pkg  # noqa: E402

ASSETS = os.path.join(ROOT, "assets")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")
    # Build what is missing. The GPU box receives prebuilt .so files with the snapshot.
    pkg = _pkg.PACKAGE_DIR
    targets = []
    if not os.path.exists(os.path.join(pkg, "librt_host.so")):
        targets.append("host")
    if not os.path.exists(os.path.join(ROOT, "oracle", "liboracle.so")):
        targets.append("oracle")
    if not os.path.exists(os.path.join(pkg, "librt_backend.so")) and (
            shutil.which("nvcc") or os.path.exists("/usr/local/cuda/bin/nvcc")):
        targets.append("backend")
    if targets:
        subprocess.run(["make", "-s"] + targets, cwd=ROOT, check=True)


@pytest.fixture(scope="session")
def rt():
    return _pkg.load()


@pytest.fixture(scope="session")
def orc():
    import oracle
    oracle.lib()
    return oracle


@pytest.fixture(scope="session")
def scene_oracle():
    import scene_oracle
    return scene_oracle


@pytest.fixture(scope="session")
def house(rt):
    return rt.Scene.load_toml(os.path.join(ASSETS, "scenes", "house.toml"))


@pytest.fixture(scope="session")
def default_scene(rt):
    return rt.Scene.load_toml(os.path.join(ASSETS, "scenes", "default.toml"))


@pytest.fixture(scope="session")
def suzanne(rt):
    return rt.Scene.load_toml(os.path.join(ASSETS, "scenes", "suzanne.toml"))


@pytest.fixture(scope="session")
def env_small(rt):
    """A 256x128 synthetic sky: same generator as the 2k one, cheap to build."""
    return rt.EnvironmentMaps([rt.Environment.synthetic(256, 128)])


@pytest.fixture(scope="session")
def env_2k(rt):
    return rt.EnvironmentMaps([rt.Environment.synthetic(2048, 1024)])


@pytest.fixture(scope="session")
def cuda_available(rt):
    import ctypes
    n = ctypes.c_int(0)
    rt._ffi.backend_lib().rt_device_count(ctypes.byref(n))
    return n.value > 0
