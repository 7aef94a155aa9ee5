"""ctypes declarations for include/rt_types.h, rt_host.h and rt_backend.h.

Loads the two in-tree shared libraries. There is no fallback of any kind: if
librt_backend.so is missing, or no CUDA device is usable, the product path raises.
"""
from __future__ import annotations

import ctypes as C
import os

_PKG_DIR = os.path.dirname(os.path.abspath(__file__))
HOST_LIB_PATH = os.path.join(_PKG_DIR, "librt_host.so")
# RT_BACKEND_LIB selects an alternative build of the SAME backend (kernel tuning variants).
BACKEND_LIB_PATH = os.environ.get("RT_BACKEND_LIB") or os.path.join(_PKG_DIR, "librt_backend.so")

# Status codes (rt_types.h).
RT_OK = 0
RT_ERR_INVALID_ARGUMENT = 1
RT_ERR_CUDA = 2
RT_ERR_OUT_OF_MEMORY = 3
RT_ERR_NCCL = 4
RT_ERR_IO = 5
RT_ERR_PARSE = 6
RT_ERR_SCENE = 7
RT_ERR_STATE = 8
RT_MISS_SLOT = 0xFFFFFFFF


class RtMaterial(C.Structure):
    _fields_ = [("color", C.c_float * 3), ("roughness", C.c_float), ("metallic", C.c_float),
                ("_pad0", C.c_float * 3), ("emission", C.c_float * 3), ("_pad1", C.c_float)]


class RtSphere(C.Structure):
    _fields_ = [("pos", C.c_float * 3), ("radius", C.c_float), ("material_id", C.c_uint32),
                ("_pad", C.c_uint32 * 3)]


class RtPlane(C.Structure):
    _fields_ = [("pos", C.c_float * 3), ("_pad0", C.c_float), ("normal", C.c_float * 3),
                ("_pad1", C.c_float), ("base_change", (C.c_float * 4) * 3),
                ("material_id", C.c_uint32), ("_pad2", C.c_uint32 * 3)]


class RtVec3(C.Structure):
    _fields_ = [("x", C.c_float), ("y", C.c_float), ("z", C.c_float), ("_pad", C.c_float)]


class RtTriangle(C.Structure):
    _fields_ = [("vertex_0", C.c_uint32), ("vertex_1", C.c_uint32), ("vertex_2", C.c_uint32),
                ("normal_0", C.c_uint32), ("normal_1", C.c_uint32), ("normal_2", C.c_uint32),
                ("material_id", C.c_uint32)]


class RtPrimitiveInfo(C.Structure):
    _fields_ = [("primitive_type", C.c_uint32), ("index", C.c_uint32)]


class RtBvhNode(C.Structure):
    _fields_ = [("min", C.c_float * 3), ("_pad0", C.c_float), ("max", C.c_float * 3),
                ("_pad1", C.c_float), ("primitives_or_second_child_index", C.c_uint32),
                ("primitives_len", C.c_uint32), ("split_axis", C.c_uint32), ("_pad2", C.c_uint32)]


class RtAliasEntry(C.Structure):
    _fields_ = [("probability", C.c_float), ("alias_index", C.c_uint32), ("pmf", C.c_float),
                ("_pad", C.c_uint32)]


class RtCameraUniform(C.Structure):
    _fields_ = [("pos", C.c_float * 3), ("_pad0", C.c_float), ("rot", (C.c_float * 4) * 3),
                ("fov_y", C.c_float), ("_pad1", C.c_float * 3)]


class RtSceneDesc(C.Structure):
    _fields_ = [("materials", C.c_void_p), ("material_count", C.c_uint32),
                ("spheres", C.c_void_p), ("sphere_count", C.c_uint32),
                ("planes", C.c_void_p), ("plane_count", C.c_uint32),
                ("vertices", C.c_void_p), ("vertex_count", C.c_uint32),
                ("normals", C.c_void_p), ("normal_count", C.c_uint32),
                ("triangles", C.c_void_p), ("triangle_count", C.c_uint32),
                ("primitives", C.c_void_p), ("primitive_count", C.c_uint32),
                ("bvh_nodes", C.c_void_p), ("bvh_node_count", C.c_uint32)]


class RtBvhBuildInput(C.Structure):
    _fields_ = [("spheres", C.c_void_p), ("sphere_count", C.c_uint32),
                ("plane_bounds", C.c_void_p), ("plane_count", C.c_uint32),
                ("vertices", C.c_void_p), ("vertex_count", C.c_uint32),
                ("triangles", C.c_void_p), ("triangle_count", C.c_uint32)]


class RtEnvironmentDesc(C.Structure):
    _fields_ = [("width", C.c_uint32), ("height", C.c_uint32), ("rgba", C.c_void_p),
                ("alias", C.c_void_p)]


class RtCounters(C.Structure):
    _fields_ = [(n, C.c_uint64) for n in (
        "paths", "rays", "shadow_rays", "nodes_tested", "tris_tested", "spheres_tested",
        "planes_tested", "hits", "tri_hits", "shaded_bounces", "env_lookups", "fallback_rays")]

    def as_dict(self) -> dict:
        return {n: int(getattr(self, n)) for n, _ in self._fields_}


assert C.sizeof(RtMaterial) == 48 and C.sizeof(RtSphere) == 32 and C.sizeof(RtPlane) == 96
assert C.sizeof(RtVec3) == 16 and C.sizeof(RtTriangle) == 28 and C.sizeof(RtPrimitiveInfo) == 8
assert C.sizeof(RtBvhNode) == 48 and C.sizeof(RtAliasEntry) == 16 and C.sizeof(RtCameraUniform) == 80

_f3 = C.POINTER(C.c_float)
_vp = C.c_void_p
_u32 = C.c_uint32

# name -> (restype, argtypes); used both to bind and by the symbol-export test.
HOST_SIGNATURES = {
    "rt_scene_load_toml": (C.c_int, [C.c_char_p, C.POINTER(_vp), C.c_char_p, C.c_size_t]),
    "rt_scene_create": (C.c_int, [C.POINTER(_vp)]),
    "rt_scene_add_material": (C.c_int, [_vp, _f3, C.c_float, C.c_float, _f3]),
    "rt_scene_add_sphere": (C.c_int, [_vp, _f3, C.c_float, _u32]),
    "rt_scene_add_plane": (C.c_int, [_vp, _f3, _f3, _f3, _u32]),
    "rt_scene_add_mesh_obj": (C.c_int, [_vp, C.c_char_p, _u32, C.c_char_p, C.c_size_t]),
    "rt_scene_add_mesh_raw": (C.c_int, [_vp, _f3, _u32, _f3, _u32, C.POINTER(_u32), _u32, _u32]),
    "rt_scene_set_camera": (C.c_int, [_vp, _f3, C.c_float, C.c_float, C.c_float]),
    "rt_scene_finalize": (C.c_int, [_vp, C.c_char_p, C.c_size_t]),
    "rt_scene_synthetic_field": (C.c_int, [C.c_char_p, _u32, _u32, C.POINTER(_vp), C.c_char_p, C.c_size_t]),
    "rt_scene_synthetic_field_ex": (C.c_int, [C.c_char_p, _u32, _u32, C.c_int, C.POINTER(_vp), C.c_char_p,
                                              C.c_size_t]),
    "rt_scene_bvh_input": (C.c_int, [_vp, C.POINTER(RtBvhBuildInput)]),
    "rt_scene_set_bvh": (C.c_int, [_vp, _vp, _u32, _vp, _u32, _u32]),
    "rt_scene_build_bvh": (C.c_int, [_vp, C.c_char_p, C.c_size_t]),
    "rt_scene_bvh_storage": (C.c_int, [_vp, _u32, C.POINTER(_vp), C.POINTER(_vp)]),
    "rt_scene_bvh_commit": (C.c_int, [_vp, _u32]),
    "rt_scene_free": (None, [_vp]),
    "rt_scene_desc": (C.c_int, [_vp, C.POINTER(RtSceneDesc)]),
    "rt_scene_camera": (C.c_int, [_vp, _f3, _f3, _f3, _f3]),
    "rt_scene_bvh_depth": (_u32, [_vp]),
    "rt_camera_uniform": (None, [_f3, C.c_float, C.c_float, C.c_float, C.POINTER(RtCameraUniform)]),
    "rt_camera_serialize": (C.c_int, [_f3, C.c_float, C.c_float, C.c_float, C.c_char_p, C.c_size_t]),
    "rt_camera_deserialize": (C.c_int, [C.c_char_p, _f3, _f3, _f3, _f3, C.c_char_p, C.c_size_t]),
    "rt_hdr_load": (C.c_int, [C.c_char_p, C.POINTER(_u32), C.POINTER(_u32), C.POINTER(_f3), C.c_char_p, C.c_size_t]),
    "rt_hdr_decode": (C.c_int, [C.c_char_p, C.c_size_t, C.POINTER(_u32), C.POINTER(_u32), C.POINTER(_f3), C.c_char_p, C.c_size_t]),
    "rt_hdr_write": (C.c_int, [C.c_char_p, _u32, _u32, _f3]),
    "rt_hdri_synthetic": (None, [_u32, _u32, _f3]),
    "rt_alias_build": (C.c_int, [_u32, _u32, _f3, _vp, C.POINTER(_u32)]),
    "rt_alias_build_serial": (C.c_int, [_u32, _u32, _f3, _vp, C.POINTER(_u32)]),
    "rt_rgb_to_rgba": (None, [_u32, _f3, _f3]),
    "rt_png_write_rgb8": (C.c_int, [C.c_char_p, _u32, _u32, _vp]),
    "rt_pfm_write_rgb": (C.c_int, [C.c_char_p, _u32, _u32, _f3, _u32]),
    "rt_free": (None, [_vp]),
}

class RtWaveProfile(C.Structure):
    """include/rt_backend.h RtWaveProfile. Kernel order: shade, finish, pre-pass, walk."""
    KERNELS = ("wave_shade_kernel", "wave_finish_kernel", "wave_prepass_kernel", "wave_walk_kernel")
    _fields_ = [("renders", C.c_uint64), ("waves", C.c_uint64), ("launches", C.c_uint64 * 4),
                ("kernel_ms", C.c_double * 4), ("gap_ms", C.c_double), ("total_ms", C.c_double)]

    def as_dict(self):
        return {"renders": int(self.renders), "waves": int(self.waves), "gap_ms": float(self.gap_ms),
                "total_ms": float(self.total_ms),
                "kernels": {k: {"launches": int(self.launches[i]), "ms": float(self.kernel_ms[i])}
                            for i, k in enumerate(self.KERNELS)}}


BACKEND_SIGNATURES = {
    "rt_backend_version": (C.c_char_p, []),
    "rt_device_count": (C.c_int, [C.POINTER(C.c_int)]),
    "rt_create": (C.c_int, [C.c_int, C.POINTER(_vp)]),
    "rt_destroy": (None, [_vp]),
    "rt_last_error": (C.c_char_p, [_vp]),
    "rt_upload_scene": (C.c_int, [_vp, C.POINTER(RtSceneDesc)]),
    "rt_upload_environments": (C.c_int, [_vp, C.POINTER(RtEnvironmentDesc), _u32]),
    "rt_set_camera": (C.c_int, [_vp, C.POINTER(RtCameraUniform)]),
    "rt_resize": (C.c_int, [_vp, _u32, _u32]),
    "rt_bind_accum": (C.c_int, [_vp, _vp]),
    "rt_accum_device_ptr": (C.c_int, [_vp, C.POINTER(_vp)]),
    "rt_clear": (C.c_int, [_vp]),
    "rt_render": (C.c_int, [_vp, _u32, _u32, _u32, _u32]),
    "rt_first_hit": (C.c_int, [_vp, _u32, _vp, _vp]),
    "rt_first_hit_kernel_ms": (C.c_float, [_vp]),
    "rt_debug_view": (C.c_int, [_vp, _u32, _u32, _u32, _vp]),
    "rt_bvh_build": (C.c_int, [_vp, C.POINTER(RtBvhBuildInput), C.POINTER(_u32), C.POINTER(_u32),
                               C.POINTER(C.c_float)]),
    "rt_bvh_build_fetch": (C.c_int, [_vp, _vp, _vp]),
    "rt_set_memory_budget": (C.c_int, [_vp, C.c_uint64, _u32]),
    "rt_get_memory_usage": (C.c_int, [_vp, C.POINTER(C.c_uint64)]),
    "rt_host_alloc": (C.c_int, [C.c_size_t, C.POINTER(C.c_void_p)]),
    "rt_host_free": (None, [_vp]),
    "rt_sync": (C.c_int, [_vp]),
    "rt_read_accum": (C.c_int, [_vp, _vp]),
    "rt_write_accum": (C.c_int, [_vp, _vp]),
    "rt_tonemap_rgb8": (C.c_int, [_vp, _u32, _vp]),
    "rt_set_profiling": (C.c_int, [_vp, C.c_int]),
    "rt_get_profile": (C.c_int, [_vp, C.POINTER(RtWaveProfile)]),
    "rt_set_detailed_counters": (C.c_int, [_vp, C.c_int]),
    "rt_reset_counters": (C.c_int, [_vp]),
    "rt_get_counters": (C.c_int, [_vp, C.POINTER(RtCounters)]),
    "rt_set_stream": (C.c_int, [_vp, _vp]),
    "rt_reset_stream": (C.c_int, [_vp]),
    "rt_timer_start": (C.c_int, [_vp]),
    "rt_timer_stop": (C.c_int, [_vp, C.POINTER(C.c_float)]),
    "rt_launch_count": (C.c_uint64, [_vp]),
    "rt_group_create": (C.c_int, [C.c_int, C.POINTER(C.c_int), C.POINTER(_vp)]),
    "rt_group_destroy": (None, [_vp]),
    "rt_group_size": (C.c_int, [_vp]),
    "rt_group_context": (_vp, [_vp, C.c_int]),
    "rt_group_render": (C.c_int, [_vp, _u32, _u32, _u32, _u32]),
    "rt_group_reduce": (C.c_int, [_vp, C.c_int]),
    "rt_group_sync": (C.c_int, [_vp]),
}


def _bind(path: str, signatures: dict) -> C.CDLL:
    if not os.path.exists(path):
        raise RuntimeError(
            f"{os.path.basename(path)} is not built ({path}); run `make` or "
            "`python -c 'import __graft_entry__ as g; g.build()'` at the repo root. "
            "There is no fallback path.")
    # RTLD_LOCAL: the backend links cudart statically; its symbols must not leak into
    # other CUDA users in the process (e.g. torch).
    lib = C.CDLL(path, mode=C.RTLD_LOCAL)
    for name, (restype, argtypes) in signatures.items():
        fn = getattr(lib, name)
        fn.restype = restype
        fn.argtypes = argtypes
    return lib


_host = None
_backend = None


def host_lib() -> C.CDLL:
    global _host
    if _host is None:
        _host = _bind(HOST_LIB_PATH, HOST_SIGNATURES)
    return _host


def backend_lib() -> C.CDLL:
    global _backend
    if _backend is None:
        _backend = _bind(BACKEND_LIB_PATH, BACKEND_SIGNATURES)
    return _backend
