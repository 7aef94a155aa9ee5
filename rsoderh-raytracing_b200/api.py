"""Host-side mirror of the reference's render-step interface, over the C ABI.

Names follow the reference: `Scene.load_toml` (src/scene.rs:235), `Camera`
(src/camera.rs:15), `EnvironmentMaps` (src/environments.rs:9), and `State` with
`update` / `render` / `resize` (src/state.rs:60, 722, 760, 651). `State.render()`
advances one sample per call exactly like one presented frame of the reference;
`State.render_samples()` is the headless batch form.

Everything below the argument marshalling happens in librt_host.so (C++) and
librt_backend.so (CUDA, sm_100a).
"""
from __future__ import annotations

import ctypes as C
import os
from dataclasses import dataclass
from typing import Optional, Sequence

import numpy as np

from . import _ffi
from ._ffi import (RtAliasEntry, RtBvhBuildInput, RtBvhNode, RtCameraUniform, RtCounters, RtEnvironmentDesc,
                   RtMaterial, RtPlane, RtPrimitiveInfo, RtSceneDesc, RtSphere, RtTriangle, RtVec3)

REFERENCE_MAX_BOUNCES = 10  # shader.wgsl:232


class RtError(RuntimeError):
    def __init__(self, status: int, message: str):
        super().__init__(message)
        self.status = status


def _f3(values) -> C.Array:
    v = [float(x) for x in values]
    if len(v) != 3:
        raise ValueError("expected 3 components")
    return (C.c_float * 3)(*v)


def _err_buf():
    return C.create_string_buffer(1024)


# --------------------------------------------------------------------------------------
# Camera (src/camera.rs:14-119)
# --------------------------------------------------------------------------------------
@dataclass
class Camera:
    pos: tuple
    yaw: float    # radians
    pitch: float  # radians
    fov_y: float  # radians

    def uniform(self) -> RtCameraUniform:
        """CameraUniform::new (camera.rs:111-119)."""
        out = RtCameraUniform()
        _ffi.host_lib().rt_camera_uniform(_f3(self.pos), self.yaw, self.pitch, self.fov_y,
                                          C.byref(out))
        return out

    def serialize(self) -> str:
        """Camera::serialize (camera.rs:30-49): the `--state` string."""
        buf = C.create_string_buffer(64)
        st = _ffi.host_lib().rt_camera_serialize(_f3(self.pos), self.yaw, self.pitch, self.fov_y,
                                                 buf, 64)
        if st:
            raise RtError(st, "camera serialisation failed")
        return buf.value.decode()

    @staticmethod
    def deserialize(encoded: str) -> "Camera":
        """Camera::deserialize (camera.rs:51-89)."""
        pos = (C.c_float * 3)()
        yaw, pitch, fov = C.c_float(), C.c_float(), C.c_float()
        err = _err_buf()
        st = _ffi.host_lib().rt_camera_deserialize(encoded.encode(), pos, C.byref(yaw),
                                                   C.byref(pitch), C.byref(fov), err, len(err))
        if st:
            raise RtError(st, err.value.decode())
        return Camera(tuple(pos), yaw.value, pitch.value, fov.value)


# --------------------------------------------------------------------------------------
# Scene (src/scene.rs, src/mesh.rs, src/bvh.rs)
# --------------------------------------------------------------------------------------
class Scene:
    """A loaded scene: the baked reference-layout buffers plus the camera."""

    def __init__(self, handle: C.c_void_p):
        self._handle = handle
        self._desc = RtSceneDesc()
        st = _ffi.host_lib().rt_scene_desc(handle, C.byref(self._desc))
        if st:
            raise RtError(st, "scene is not finalised")
        pos = (C.c_float * 3)()
        yaw, pitch, fov = C.c_float(), C.c_float(), C.c_float()
        _ffi.host_lib().rt_scene_camera(handle, pos, C.byref(yaw), C.byref(pitch), C.byref(fov))
        self.camera = Camera(tuple(pos), yaw.value, pitch.value, fov.value)

    def __del__(self):
        h, self._handle = getattr(self, "_handle", None), None
        if h:
            _ffi.host_lib().rt_scene_free(h)

    @staticmethod
    def load_toml(path: str) -> "Scene":
        """Scene::load_toml (scene.rs:235-252). Raises RtError with the reference's text."""
        handle = C.c_void_p()
        err = _err_buf()
        st = _ffi.host_lib().rt_scene_load_toml(os.fspath(path).encode(), C.byref(handle), err,
                                                len(err))
        if st:
            raise RtError(st, err.value.decode())
        return Scene(handle)

    @staticmethod
    def synthetic_field(assets_dir: str, grid: int, seed: int = 0x5EED, build_bvh: bool = True) -> "Scene":
        """build_bvh=False leaves the BVH to `DeviceBvhBuilder.build(scene)`."""
        handle = C.c_void_p()
        err = _err_buf()
        st = _ffi.host_lib().rt_scene_synthetic_field_ex(os.fspath(assets_dir).encode(), grid, seed,
                                                         1 if build_bvh else 0, C.byref(handle), err, len(err))
        if st:
            raise RtError(st, err.value.decode())
        return Scene(handle)

    def refresh(self):
        """Re-reads the buffer pointers after the BVH was replaced."""
        st = _ffi.host_lib().rt_scene_desc(self._handle, C.byref(self._desc))
        if st:
            raise RtError(st, "scene is not finalised")

    def bvh_input(self) -> RtBvhBuildInput:
        """Input of build_bvh as the device builder takes it (rt_scene_bvh_input)."""
        inp = RtBvhBuildInput()
        st = _ffi.host_lib().rt_scene_bvh_input(self._handle, C.byref(inp))
        if st:
            raise RtError(st, "scene is not finalised")
        return inp

    def plane_bounds(self) -> np.ndarray:
        """Plane::bounds (scene.rs:203-207) of every plane: [min xyz, max xyz]."""
        inp = self.bvh_input()
        if not inp.plane_count:
            return np.zeros((0, 6), dtype=np.float32)
        buf = C.cast(inp.plane_bounds, C.POINTER(C.c_float * (6 * inp.plane_count))).contents
        return np.frombuffer(buf, dtype=np.float32).reshape(inp.plane_count, 6).copy()

    def build_bvh_host(self):
        """build_bvh (bvh.rs:13-28) with the host builder (what load_toml does)."""
        err = _err_buf()
        st = _ffi.host_lib().rt_scene_build_bvh(self._handle, err, len(err))
        if st:
            raise RtError(st, err.value.decode())
        self.refresh()

    @property
    def desc(self) -> RtSceneDesc:
        return self._desc

    @property
    def bvh_depth(self) -> int:
        return int(_ffi.host_lib().rt_scene_bvh_depth(self._handle))

    def _array(self, ptr, count, ctype, dtype, cols):
        if not count:
            return np.zeros((0, cols), dtype=dtype)
        buf = C.cast(ptr, C.POINTER(C.c_uint8 * (count * C.sizeof(ctype)))).contents
        return np.frombuffer(buf, dtype=dtype).reshape(count, cols).copy()

    # Raw views of the reference-layout buffers (copies), as u32/f32 word matrices.
    def bvh_nodes_words(self) -> np.ndarray:
        return self._array(self._desc.bvh_nodes, self._desc.bvh_node_count, RtBvhNode, np.uint32, 12)

    def primitives_words(self) -> np.ndarray:
        return self._array(self._desc.primitives, self._desc.primitive_count, RtPrimitiveInfo,
                           np.uint32, 2)

    def triangles_words(self) -> np.ndarray:
        return self._array(self._desc.triangles, self._desc.triangle_count, RtTriangle, np.uint32, 7)

    def vertices(self) -> np.ndarray:
        return self._array(self._desc.vertices, self._desc.vertex_count, RtVec3, np.float32, 4)[:, :3]

    def normals(self) -> np.ndarray:
        return self._array(self._desc.normals, self._desc.normal_count, RtVec3, np.float32, 4)[:, :3]

    def spheres_words(self) -> np.ndarray:
        return self._array(self._desc.spheres, self._desc.sphere_count, RtSphere, np.float32, 8)

    def planes_words(self) -> np.ndarray:
        return self._array(self._desc.planes, self._desc.plane_count, RtPlane, np.float32, 24)

    def materials_words(self) -> np.ndarray:
        return self._array(self._desc.materials, self._desc.material_count, RtMaterial, np.float32, 12)

    def counts(self) -> dict:
        d = self._desc
        return {"materials": d.material_count, "spheres": d.sphere_count, "planes": d.plane_count,
                "vertices": d.vertex_count, "normals": d.normal_count,
                "triangles": d.triangle_count, "primitives": d.primitive_count,
                "bvh_nodes": d.bvh_node_count}


class DeviceBvhBuilder:
    """build_bvh (bvh.rs:13-28) on the GPU: rt_bvh_build / rt_bvh_build_fetch. The result is
    byte-identical to the host builder's; `build(scene)` installs it into the scene."""

    def __init__(self, device: int = 0):
        self._lib = _ffi.backend_lib()
        self._ctx = C.c_void_p()
        st = self._lib.rt_create(device, C.byref(self._ctx))
        if st:
            raise RtError(st, (self._lib.rt_last_error(None) or b"").decode())

    def close(self):
        ctx, self._ctx = getattr(self, "_ctx", None), None
        if ctx:
            self._lib.rt_destroy(ctx)

    def __del__(self):
        self.close()

    def _check(self, st: int):
        if st:
            raise RtError(st, (self._lib.rt_last_error(self._ctx) or b"").decode())

    def build(self, scene: Scene, install: bool = True) -> dict:
        import time
        inp = RtBvhBuildInput()
        st = _ffi.host_lib().rt_scene_bvh_input(scene._handle, C.byref(inp))
        if st:
            raise RtError(st, "scene is not finalised")
        prims = inp.sphere_count + inp.plane_count + inp.triangle_count
        node_count, depth, ms = C.c_uint32(), C.c_uint32(), C.c_float()
        t0 = time.perf_counter()
        self._check(self._lib.rt_bvh_build(self._ctx, C.byref(inp), C.byref(node_count), C.byref(depth), C.byref(ms)))
        t1 = time.perf_counter()
        if install:
            # straight into the scene's own arrays: one pass over the memory
            nodes_ptr, prims_ptr = C.c_void_p(), C.c_void_p()
            st = _ffi.host_lib().rt_scene_bvh_storage(scene._handle, node_count.value, C.byref(nodes_ptr), C.byref(prims_ptr))
            if st:
                raise RtError(st, "rt_scene_bvh_storage refused the device-built BVH")
            self._check(self._lib.rt_bvh_build_fetch(self._ctx, nodes_ptr, prims_ptr))
            t2 = time.perf_counter()
            _ffi.host_lib().rt_scene_bvh_commit(scene._handle, depth.value)
            scene.refresh()
            nodes = primitives = None
        else:
            nodes = np.empty((node_count.value, 12), dtype=np.uint32)
            primitives = np.empty((prims, 2), dtype=np.uint32)
            self._check(self._lib.rt_bvh_build_fetch(self._ctx, nodes.ctypes.data, primitives.ctypes.data))
            t2 = time.perf_counter()
        return {"nodes": nodes, "primitives": primitives, "node_count": int(node_count.value), "depth": int(depth.value),
                "device_ms": float(ms.value), "build_call_ms": (t1 - t0) * 1e3, "fetch_ms": (t2 - t1) * 1e3,
                "install_ms": (time.perf_counter() - t2) * 1e3}


class SceneBuilder:
    """Programmatic scene construction through the same bake path as the TOML loader."""

    def __init__(self):
        self._handle = C.c_void_p()
        _ffi.host_lib().rt_scene_create(C.byref(self._handle))
        self._done = False

    def __del__(self):
        if not self._done and self._handle:
            _ffi.host_lib().rt_scene_free(self._handle)

    def _check(self, st):
        if st:
            raise RtError(st, "invalid argument")

    def add_material(self, color, roughness, metallic, emission=(0, 0, 0)) -> "SceneBuilder":
        self._check(_ffi.host_lib().rt_scene_add_material(self._handle, _f3(color), roughness,
                                                          metallic, _f3(emission)))
        return self

    def add_sphere(self, pos, radius, material_id) -> "SceneBuilder":
        self._check(_ffi.host_lib().rt_scene_add_sphere(self._handle, _f3(pos), radius, material_id))
        return self

    def add_plane(self, pos, forward, right, material_id) -> "SceneBuilder":
        self._check(_ffi.host_lib().rt_scene_add_plane(self._handle, _f3(pos), _f3(forward),
                                                       _f3(right), material_id))
        return self

    def add_mesh_obj(self, obj_text: str, material_id: int) -> "SceneBuilder":
        err = _err_buf()
        st = _ffi.host_lib().rt_scene_add_mesh_obj(self._handle, obj_text.encode(), material_id,
                                                   err, len(err))
        if st:
            raise RtError(st, err.value.decode())
        return self

    def add_mesh_raw(self, vertices, normals, triangles, material_id) -> "SceneBuilder":
        v = np.ascontiguousarray(vertices, dtype=np.float32).reshape(-1, 3)
        n = np.ascontiguousarray(normals, dtype=np.float32).reshape(-1, 3)
        t = np.ascontiguousarray(triangles, dtype=np.uint32).reshape(-1, 6)
        self._check(_ffi.host_lib().rt_scene_add_mesh_raw(
            self._handle, v.ctypes.data_as(C.POINTER(C.c_float)), len(v),
            n.ctypes.data_as(C.POINTER(C.c_float)), len(n),
            t.ctypes.data_as(C.POINTER(C.c_uint32)), len(t), material_id))
        return self

    def set_camera(self, pos, yaw, pitch, fov_y) -> "SceneBuilder":
        self._check(_ffi.host_lib().rt_scene_set_camera(self._handle, _f3(pos), yaw, pitch, fov_y))
        return self

    def build(self) -> Scene:
        err = _err_buf()
        st = _ffi.host_lib().rt_scene_finalize(self._handle, err, len(err))
        if st:
            raise RtError(st, err.value.decode())
        self._done = True
        return Scene(self._handle)


# --------------------------------------------------------------------------------------
# Environments (src/environments.rs)
# --------------------------------------------------------------------------------------
class Environment:
    """One HDRI: rgb float32 [H, W, 3], its rgba texel buffer and alias table."""

    def __init__(self, rgb: np.ndarray):
        rgb = np.ascontiguousarray(rgb, dtype=np.float32)
        if rgb.ndim != 3 or rgb.shape[2] != 3:
            raise ValueError("rgb must be [H, W, 3]")
        self.height, self.width = int(rgb.shape[0]), int(rgb.shape[1])
        self.rgb = rgb
        n = self.width * self.height
        self.rgba = np.zeros((self.height, self.width, 4), dtype=np.float32)
        _ffi.host_lib().rt_rgb_to_rgba(n, rgb.ctypes.data_as(C.POINTER(C.c_float)),
                                       self.rgba.ctypes.data_as(C.POINTER(C.c_float)))
        # AliasTable::build_by_luminance (environments.rs:97-186)
        self.alias = np.zeros((n, 4), dtype=np.uint32)
        leftover = C.c_uint32()
        st = _ffi.host_lib().rt_alias_build(self.width, self.height,
                                            rgb.ctypes.data_as(C.POINTER(C.c_float)),
                                            self.alias.ctypes.data_as(C.c_void_p),
                                            C.byref(leftover))
        if st:
            raise RtError(st, "alias table build failed")
        self.leftover_count = int(leftover.value)

    @staticmethod
    def synthetic(width: int = 2048, height: int = 1024) -> "Environment":
        rgb = np.zeros((height, width, 3), dtype=np.float32)
        _ffi.host_lib().rt_hdri_synthetic(width, height, rgb.ctypes.data_as(C.POINTER(C.c_float)))
        return Environment(rgb)

    @staticmethod
    def load_hdr(path: str) -> "Environment":
        w, h = C.c_uint32(), C.c_uint32()
        ptr = C.POINTER(C.c_float)()
        err = _err_buf()
        st = _ffi.host_lib().rt_hdr_load(os.fspath(path).encode(), C.byref(w), C.byref(h),
                                         C.byref(ptr), err, len(err))
        if st:
            raise RtError(st, err.value.decode())
        try:
            rgb = np.ctypeslib.as_array(ptr, (h.value, w.value, 3)).copy()
        finally:
            _ffi.host_lib().rt_free(ptr)
        return Environment(rgb)

    def desc(self) -> RtEnvironmentDesc:
        return RtEnvironmentDesc(self.width, self.height, self.rgba.ctypes.data, self.alias.ctypes.data)


class EnvironmentMaps:
    """EnvironmentMaps::new (environments.rs:19-64): the environment slots, in order."""

    def __init__(self, environments: Sequence[Environment]):
        if not environments:
            raise ValueError("at least one environment is required")
        self.environments = list(environments)

    @staticmethod
    def default(assets_dir: Optional[str] = None, width: int = 2048, height: int = 1024) -> "EnvironmentMaps":
        """Slots as in src/state.rs:119-122 when the Poly Haven files are present under
        assets/hdri/, otherwise the deterministic synthetic sky in slot 0."""
        envs = []
        if assets_dir:
            for name in ("winter_lake_01_2k.hdr", "passendorf_snow_2k.hdr"):
                p = os.path.join(assets_dir, "hdri", name)
                if os.path.exists(p):
                    envs.append(Environment.load_hdr(p))
        if not envs:
            envs.append(Environment.synthetic(width, height))
        return EnvironmentMaps(envs)

    def descs(self):
        arr = (RtEnvironmentDesc * len(self.environments))()
        for i, e in enumerate(self.environments):
            arr[i] = e.desc()
        return arr


# --------------------------------------------------------------------------------------
# State (src/state.rs): device context + resources + frame loop
# --------------------------------------------------------------------------------------
class State:
    """The render-step object. Owns one CUDA context on one device."""

    def __init__(self, scene: Scene, environments: EnvironmentMaps, width: int, height: int,
                 device: int = 0, max_bounces: int = REFERENCE_MAX_BOUNCES,
                 environment_index: int = 0):
        self._lib = _ffi.backend_lib()
        self._ctx = C.c_void_p()
        st = self._lib.rt_create(device, C.byref(self._ctx))
        if st:
            raise RtError(st, (self._lib.rt_last_error(None) or b"").decode())
        self.scene = scene
        self.environments = environments
        self.camera = scene.camera
        self.max_bounces = max_bounces
        self.environment_index = environment_index
        self.sample_count = 0  # HdrPipeline.sample_count (hdr.rs:70): next sample index
        self.width = self.height = 0
        self._check(self._lib.rt_upload_scene(self._ctx, C.byref(scene.desc)))
        self._check(self._lib.rt_upload_environments(self._ctx, environments.descs(),
                                                     len(environments.environments)))
        self.resize(width, height)
        self.update()

    # -- plumbing --
    def _check(self, st: int):
        if st:
            raise RtError(st, (self._lib.rt_last_error(self._ctx) or b"").decode())

    def close(self):
        ctx, self._ctx = getattr(self, "_ctx", None), None
        if ctx:
            self._lib.rt_destroy(ctx)

    def __del__(self):
        self.close()

    def upload_scene(self, scene: Scene):
        """Scene bind group upload (state.rs:394-458) from the host buffers."""
        self._check(self._lib.rt_upload_scene(self._ctx, C.byref(scene.desc)))
        self.scene = scene

    # -- State::resize (state.rs:651-700) --
    def resize(self, width: int, height: int):
        self._check(self._lib.rt_resize(self._ctx, width, height))
        self.width, self.height = int(width), int(height)
        self.sample_count = 0

    # -- State::update (state.rs:722-758): push the camera uniform --
    def update(self, camera: Optional[Camera] = None):
        if camera is not None:
            if camera != self.camera:
                self.reset()
            self.camera = camera
        u = self.camera.uniform()
        self._check(self._lib.rt_set_camera(self._ctx, C.byref(u)))

    def reset(self):
        """Scene-hash change (state.rs:775-789): restart accumulation."""
        self._check(self._lib.rt_clear(self._ctx))
        self.sample_count = 0

    # -- State::render (state.rs:760-833): one frame = one sample per pixel --
    def render(self):
        self.render_samples(1)

    def render_samples(self, count: int):
        """`count` consecutive frames in one launch (sample indices continue)."""
        self._check(self._lib.rt_render(self._ctx, self.sample_count, count, self.max_bounces,
                                        self.environment_index))
        self.sample_count += count

    def render_range(self, sample_begin: int, count: int):
        """Explicit sample-index range (multi-GPU sample split)."""
        self._check(self._lib.rt_render(self._ctx, sample_begin, count, self.max_bounces,
                                        self.environment_index))

    def sync(self):
        self._check(self._lib.rt_sync(self._ctx))

    def read_accum(self, out: Optional[np.ndarray] = None) -> np.ndarray:
        if out is None:
            out = np.empty((self.height, self.width, 4), dtype=np.float32)
        self._check(self._lib.rt_read_accum(self._ctx, out.ctypes.data))
        return out

    def write_accum(self, rgba: np.ndarray):
        rgba = np.ascontiguousarray(rgba, dtype=np.float32)
        assert rgba.shape == (self.height, self.width, 4)
        self._check(self._lib.rt_write_accum(self._ctx, rgba.ctypes.data))

    def first_hit(self, sample_index: int = 0, out=None):
        """Primary rays only (rt_first_hit). `out` = (slots, t) arrays to fill — e.g. a pair from
        pinned_array(), which the device copies into at link speed."""
        if out is None:
            slots = np.empty((self.height, self.width), dtype=np.uint32)
            t = np.empty((self.height, self.width), dtype=np.float32)
        else:
            slots, t = out
            assert slots.shape == (self.height, self.width) and slots.dtype == np.uint32
            assert t.shape == (self.height, self.width) and t.dtype == np.float32
        self._check(self._lib.rt_first_hit(self._ctx, sample_index, slots.ctypes.data, t.ctypes.data))
        return slots, t

    def first_hit_kernel_ms(self) -> float:
        """Device time of the last first_hit() kernel, without its device-to-host copies."""
        return float(self._lib.rt_first_hit_kernel_ms(self._ctx))

    def tonemap_rgb8(self, sample_total: Optional[int] = None) -> np.ndarray:
        out = np.empty((self.height, self.width, 3), dtype=np.uint8)
        total = self.sample_count if sample_total is None else sample_total
        self._check(self._lib.rt_tonemap_rgb8(self._ctx, total, out.ctypes.data))
        return out

    # -- counters / timing / interop --
    def debug_view(self, dev_index: int, sample_index: int = 0) -> np.ndarray:
        """out_texture of the reference's dev_index 2 / 3 views (shader.wgsl:1314-1338)."""
        out = np.empty((self.height, self.width, 4), dtype=np.float32)
        self._check(self._lib.rt_debug_view(self._ctx, dev_index, self.environment_index, sample_index,
                                            out.ctypes.data))
        return out

    def next_environment(self):
        """The `next_environment` key (camera.rs:271-278): cycle the slot, restart accumulation."""
        self.environment_index += 1
        if self.environment_index >= len(self.environments.environments):
            self.environment_index = 0
        self.reset()

    def set_memory_budget(self, staging_bytes: int = 0, path_slots: int = 0):
        """rt_set_memory_budget: upper bounds of the staging planes / path-slot arena (0 keeps)."""
        self._check(self._lib.rt_set_memory_budget(self._ctx, staging_bytes, path_slots))

    def memory_usage(self) -> int:
        n = C.c_uint64()
        self._check(self._lib.rt_get_memory_usage(self._ctx, C.byref(n)))
        return int(n.value)

    def set_detailed_counters(self, enabled: bool):
        self._check(self._lib.rt_set_detailed_counters(self._ctx, int(enabled)))

    def set_profiling(self, enabled: bool):
        """Per-kernel CUDA-event timing of the wavefront pipeline (clears the profile)."""
        self._check(self._lib.rt_set_profiling(self._ctx, int(enabled)))

    def profile(self) -> dict:
        prof = _ffi.RtWaveProfile()
        self._check(self._lib.rt_get_profile(self._ctx, C.byref(prof)))
        return prof.as_dict()

    def reset_counters(self):
        self._check(self._lib.rt_reset_counters(self._ctx))

    def counters(self) -> dict:
        c = RtCounters()
        self._check(self._lib.rt_get_counters(self._ctx, C.byref(c)))
        return c.as_dict()

    def set_stream(self, cuda_stream: int):
        """Enqueue on this cudaStream_t handle (0 = the legacy default stream)."""
        self._check(self._lib.rt_set_stream(self._ctx, C.c_void_p(cuda_stream)))

    def reset_stream(self):
        self._check(self._lib.rt_reset_stream(self._ctx))

    def bind_accum(self, device_ptr: int):
        self._check(self._lib.rt_bind_accum(self._ctx, C.c_void_p(device_ptr)))

    def accum_device_ptr(self) -> int:
        p = C.c_void_p()
        self._check(self._lib.rt_accum_device_ptr(self._ctx, C.byref(p)))
        return int(p.value or 0)

    def timer_start(self):
        self._check(self._lib.rt_timer_start(self._ctx))

    def timer_stop(self) -> float:
        ms = C.c_float()
        self._check(self._lib.rt_timer_stop(self._ctx, C.byref(ms)))
        return float(ms.value)

    def launch_count(self) -> int:
        return int(self._lib.rt_launch_count(self._ctx))


class StateGroup:
    """rt_group_*: one host process driving several GPUs — the frame's sample range is split
    into contiguous per-GPU ranges and the accumulation buffers are sum-reduced onto the root
    (SURVEY §8e). After reduce() the root holds the sum with alpha = 1 and the other members
    are zero again, so render() + reduce() rounds can follow each other (progressive render)."""

    def __init__(self, scene: Scene, environments: EnvironmentMaps, width: int, height: int,
                 devices: Sequence[int], max_bounces: int = REFERENCE_MAX_BOUNCES, environment_index: int = 0):
        self._lib = _ffi.backend_lib()
        self._group = C.c_void_p()
        arr = (C.c_int * len(devices))(*devices)
        st = self._lib.rt_group_create(len(devices), arr, C.byref(self._group))
        if st:
            raise RtError(st, (self._lib.rt_last_error(None) or b"").decode())
        self.size = int(self._lib.rt_group_size(self._group))
        self.max_bounces, self.environment_index = max_bounces, environment_index
        self.width, self.height = int(width), int(height)
        self.sample_count = 0
        self._keep = (scene, environments)
        u = scene.camera.uniform()
        for r in range(self.size):
            ctx = self.context(r)
            self._check(r, self._lib.rt_upload_scene(ctx, C.byref(scene.desc)))
            self._check(r, self._lib.rt_upload_environments(ctx, environments.descs(), len(environments.environments)))
            self._check(r, self._lib.rt_resize(ctx, width, height))
            self._check(r, self._lib.rt_set_camera(ctx, C.byref(u)))

    def context(self, rank: int):
        return C.c_void_p(self._lib.rt_group_context(self._group, rank))

    def _check(self, rank, st):
        if st:
            raise RtError(st, (self._lib.rt_last_error(self.context(rank)) or b"").decode())

    def render_samples(self, count: int):
        self._check(0, self._lib.rt_group_render(self._group, self.sample_count, count, self.max_bounces,
                                                 self.environment_index))
        self.sample_count += count

    def reduce(self, root: int = 0):
        self._check(root, self._lib.rt_group_reduce(self._group, root))

    def sync(self):
        self._check(0, self._lib.rt_group_sync(self._group))

    def read_accum(self, rank: int = 0) -> np.ndarray:
        out = np.empty((self.height, self.width, 4), dtype=np.float32)
        self._check(rank, self._lib.rt_read_accum(self.context(rank), out.ctypes.data))
        return out

    def counters(self, rank: int) -> dict:
        c = RtCounters()
        self._check(rank, self._lib.rt_get_counters(self.context(rank), C.byref(c)))
        return c.as_dict()

    def close(self):
        g, self._group = getattr(self, "_group", None), None
        if g:
            self._lib.rt_group_destroy(g)

    def __del__(self):
        self.close()


class _PinnedBlock:
    def __init__(self, nbytes: int):
        self.ptr = C.c_void_p()
        st = _ffi.backend_lib().rt_host_alloc(nbytes, C.byref(self.ptr))
        if st:
            raise RtError(st, (_ffi.backend_lib().rt_last_error(None) or b"").decode())

    def __del__(self):
        ptr, self.ptr = getattr(self, "ptr", None), None
        if ptr:
            _ffi.backend_lib().rt_host_free(ptr)


def pinned_array(shape, dtype) -> np.ndarray:
    """A numpy array over page-locked host memory (rt_host_alloc); freed with the array."""
    dtype = np.dtype(dtype)
    n = int(np.prod(shape))
    block = _PinnedBlock(max(1, n * dtype.itemsize))
    raw = (C.c_uint8 * (n * dtype.itemsize)).from_address(block.ptr.value)
    raw._block = block  # the ctypes view, and through it the array, keep the allocation alive
    return np.frombuffer(raw, dtype=dtype, count=n).reshape(shape)


def write_png(path: str, rgb8: np.ndarray):
    rgb8 = np.ascontiguousarray(rgb8, dtype=np.uint8)
    st = _ffi.host_lib().rt_png_write_rgb8(os.fspath(path).encode(), rgb8.shape[1], rgb8.shape[0],
                                           rgb8.ctypes.data)
    if st:
        raise RtError(st, f"cannot write {path}")


def write_pfm(path: str, rgba: np.ndarray):
    rgba = np.ascontiguousarray(rgba, dtype=np.float32)
    st = _ffi.host_lib().rt_pfm_write_rgb(os.fspath(path).encode(), rgba.shape[1], rgba.shape[0],
                                          rgba.ctypes.data_as(C.POINTER(C.c_float)), rgba.shape[2])
    if st:
        raise RtError(st, f"cannot write {path}")
