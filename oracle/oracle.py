"""ctypes binding of oracle/liboracle.so (shader_oracle.cpp). TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / `--impl reference`
leg import this module. PARITY UNPINNED: see the header of shader_oracle.cpp.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
import sys

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_ROOT = os.path.dirname(_HERE)
LIB_PATH = os.path.join(_HERE, "liboracle.so")

if _ROOT not in sys.path:
    sys.path.insert(0, _ROOT)
import _pkg  # noqa: E402

_rt = _pkg.load()
_ffi = _rt._ffi


class OracleJob(C.Structure):
    _fields_ = [("scene", _ffi.RtSceneDesc), ("env_count", C.c_uint32),
                ("envs", C.POINTER(_ffi.RtEnvironmentDesc)), ("camera", _ffi.RtCameraUniform),
                ("width", C.c_uint32), ("height", C.c_uint32), ("max_bounces", C.c_uint32),
                ("env_index", C.c_uint32), ("prune", C.c_uint32)]


OracleCounters = _ffi.RtCounters

_lib = None


def build():
    subprocess.run(["make", "-s", "oracle"], cwd=_ROOT, check=True)


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            build()
        l = C.CDLL(LIB_PATH)
        l.oracle_render.restype = C.c_int
        l.oracle_render.argtypes = [C.POINTER(OracleJob), C.c_uint32, C.c_uint32, C.c_void_p,
                                    C.POINTER(OracleCounters), C.c_int]
        l.oracle_first_hit.restype = C.c_int
        l.oracle_first_hit.argtypes = [C.POINTER(OracleJob), C.c_uint32, C.c_void_p, C.c_void_p,
                                       C.POINTER(OracleCounters), C.c_int]
        l.oracle_cast_rays.restype = C.c_int
        l.oracle_cast_rays.argtypes = [C.POINTER(OracleJob), C.c_uint32, C.c_void_p, C.c_void_p,
                                       C.c_int, C.c_void_p, C.c_void_p]
        l.oracle_trace_pixel.restype = C.c_int
        l.oracle_trace_pixel.argtypes = [C.POINTER(OracleJob), C.c_uint32, C.c_uint32, C.c_uint32,
                                         C.POINTER(C.c_float)]
        l.oracle_camera_ray.restype = C.c_int
        l.oracle_camera_ray.argtypes = [C.POINTER(OracleJob), C.c_uint32, C.c_uint32, C.c_uint32,
                                        C.POINTER(C.c_float), C.POINTER(C.c_float),
                                        C.POINTER(C.c_uint32)]
        for name in ("oracle_sin", "oracle_cos", "oracle_asin"):
            f = getattr(l, name)
            f.restype = C.c_float
            f.argtypes = [C.c_float]
        l.oracle_atan2.restype = C.c_float
        l.oracle_atan2.argtypes = [C.c_float, C.c_float]
        l.oracle_rng_next.restype = C.c_uint32
        l.oracle_rng_next.argtypes = [C.POINTER(C.c_uint32)]
        l.oracle_rng_uniform.restype = C.c_float
        l.oracle_rng_uniform.argtypes = [C.POINTER(C.c_uint32)]
        l.oracle_rng_seed.restype = None
        l.oracle_rng_seed.argtypes = [C.c_uint32, C.c_uint32, C.POINTER(C.c_uint32)]
        l.oracle_counters_size.restype = C.c_uint32
        l.oracle_job_size.restype = C.c_uint32
        assert l.oracle_counters_size() == C.sizeof(OracleCounters)
        assert l.oracle_job_size() == C.sizeof(OracleJob)
        _lib = l
    return _lib


class Oracle:
    """The reference shader restated on the CPU, over one scene + environments + camera."""

    def __init__(self, scene, environments, width, height, camera=None, max_bounces=10,
                 environment_index=0, prune=False):
        self._scene = scene          # keeps the host buffers alive
        self._envs = environments
        self._env_descs = environments.descs()
        self.job = OracleJob()
        self.job.scene = scene.desc
        self.job.env_count = len(environments.environments)
        self.job.envs = C.cast(self._env_descs, C.POINTER(_ffi.RtEnvironmentDesc))
        self.job.camera = (camera or scene.camera).uniform()
        self.job.width, self.job.height = width, height
        self.job.max_bounces = max_bounces
        self.job.env_index = environment_index
        self.job.prune = int(prune)
        self.width, self.height = width, height

    def render(self, sample_begin, sample_count, accum=None, threads=0):
        """Returns (accum [H, W, 4] float32, counters dict)."""
        if accum is None:
            accum = np.zeros((self.height, self.width, 4), dtype=np.float32)
        counters = OracleCounters()
        st = lib().oracle_render(C.byref(self.job), sample_begin, sample_count, accum.ctypes.data,
                                 C.byref(counters), threads)
        if st:
            raise RuntimeError("oracle_render: invalid job")
        return accum, counters.as_dict()

    def first_hit(self, sample_index=0, threads=0):
        slots = np.empty((self.height, self.width), dtype=np.uint32)
        t = np.empty((self.height, self.width), dtype=np.float32)
        counters = OracleCounters()
        st = lib().oracle_first_hit(C.byref(self.job), sample_index, slots.ctypes.data,
                                    t.ctypes.data, C.byref(counters), threads)
        if st:
            raise RuntimeError("oracle_first_hit: invalid job")
        return slots, t, counters.as_dict()

    def cast_rays(self, origins, dirs, shadow=False):
        """cast_ray (or the shadow-ray occlusion query) for arbitrary rays [n, 3] float32."""
        origins = np.ascontiguousarray(origins, dtype=np.float32)
        dirs = np.ascontiguousarray(dirs, dtype=np.float32)
        n = origins.shape[0]
        slots = np.empty(n, dtype=np.uint32)
        t = np.empty(n, dtype=np.float32)
        st = lib().oracle_cast_rays(C.byref(self.job), n, origins.ctypes.data, dirs.ctypes.data,
                                    int(shadow), slots.ctypes.data, t.ctypes.data)
        if st:
            raise RuntimeError("oracle_cast_rays: invalid job")
        return slots, t

    def trace_pixel(self, x, y, sample_index):
        rgb = (C.c_float * 3)()
        st = lib().oracle_trace_pixel(C.byref(self.job), x, y, sample_index, rgb)
        if st:
            raise RuntimeError("oracle_trace_pixel: invalid arguments")
        return np.array(list(rgb), dtype=np.float32)

    def camera_ray(self, x, y, sample_index):
        o, d = (C.c_float * 3)(), (C.c_float * 3)()
        rng = C.c_uint32()
        lib().oracle_camera_ray(C.byref(self.job), x, y, sample_index, o, d, C.byref(rng))
        return np.array(list(o), np.float32), np.array(list(d), np.float32), int(rng.value)


def tonemap_rgb8(accum: np.ndarray, sample_total: int) -> np.ndarray:
    """Present pass restated: accum / n -> f16 -> ACES fit (src/shaders/hdr.wgsl:3-22)
    -> sRGB encode -> 8 bit."""
    f32 = np.float32
    hdr = (accum[..., :3].astype(f32) / f32(sample_total)).astype(np.float16).astype(f32)
    m1 = np.array([[0.59719, 0.07600, 0.02840], [0.35458, 0.90834, 0.13383],
                   [0.04823, 0.01566, 0.83777]], dtype=f32)  # rows here = WGSL columns
    m2 = np.array([[1.60475, -0.10208, -0.00327], [-0.53108, 1.10813, -0.07276],
                   [-0.07367, -0.00605, 1.07602]], dtype=f32)
    v = (hdr[..., 0:1] * m1[0] + hdr[..., 1:2] * m1[1]) + hdr[..., 2:3] * m1[2]
    a = v * (v + f32(0.0245786)) - f32(0.000090537)
    b = v * (f32(0.983729) * v + f32(0.4329510)) + f32(0.238081)
    r = a / b
    out = (r[..., 0:1] * m2[0] + r[..., 1:2] * m2[1]) + r[..., 2:3] * m2[2]
    out = np.clip(out, 0.0, 1.0).astype(f32)
    negative = (hdr < 0).any(axis=-1)
    out[negative] = np.array([1.0, 0.0, 1.0], dtype=f32)
    srgb = np.where(out <= f32(0.0031308), f32(12.92) * out,
                    f32(1.055) * np.power(out, f32(1.0 / 2.4)) - f32(0.055))
    return (np.clip(srgb, 0.0, 1.0) * 255.0 + 0.5).astype(np.uint8)
