"""A scene, camera and environment built WITHOUT the product's host library. TEST INFRASTRUCTURE ONLY.

oracle/scene_oracle.py (tomllib + numpy float32) restates scene.rs / mesh.rs / bvh.rs /
environments.rs / camera.rs; this module packs its results into the byte layouts of
include/rt_types.h, so that liboracle.so (shader_oracle.cpp) can render from buffers that
librt_host.so never touched. An image rendered this way and the CUDA backend's image of the
same TOML then agree only if the BVH order, the packing and the alias table of the product's
front-end agree with the restatement too. PARITY UNPINNED: see shader_oracle.cpp.
"""
from __future__ import annotations

import ctypes as C
import os
import sys

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_ROOT = os.path.dirname(_HERE)
for _p in (_ROOT, _HERE):
    if _p not in sys.path:
        sys.path.insert(0, _p)
import _pkg  # noqa: E402
import scene_oracle  # noqa: E402

_ffi = _pkg.load()._ffi
f32 = np.float32


class _Camera:
    def __init__(self, pos, yaw, pitch, fov_y):
        self.pos, self.yaw, self.pitch, self.fov_y = pos, yaw, pitch, fov_y

    def uniform(self):
        """CameraUniform::new (camera.rs:111-119) from scene_oracle.rot_transform."""
        u = _ffi.RtCameraUniform()
        cols = scene_oracle.rot_transform(self.yaw, self.pitch)
        for k in range(3):
            u.pos[k] = float(self.pos[k])
            for j in range(3):
                u.rot[k][j] = float(cols[k][j])
        u.fov_y = float(self.fov_y)
        return u


class IndependentScene:
    """Duck-types api.Scene for oracle.Oracle: `.desc` (RtSceneDesc) and `.camera`."""

    def __init__(self, toml_path: str):
        ref = scene_oracle.load_scene_toml(toml_path)
        nodes, ordered = scene_oracle.build_bvh(ref)
        u32 = np.uint32

        mats = np.zeros((len(ref.materials), 12), dtype=f32)
        for i, (color, rough, metal, emission) in enumerate(ref.materials):
            mats[i, 0:3], mats[i, 3], mats[i, 4], mats[i, 8:11] = color, rough, metal, emission
        spheres = np.zeros((len(ref.spheres), 8), dtype=f32)
        for i, (pos, r, mat) in enumerate(ref.spheres):
            spheres[i, 0:3], spheres[i, 3] = pos, r
            spheres[i, 4:5].view(u32)[0] = mat
        planes = np.zeros((len(ref.planes), 24), dtype=f32)
        for i, (pos, fwd, right, mat) in enumerate(ref.planes):
            normal, cols = scene_oracle.plane_to_uniform(pos, fwd, right)
            planes[i, 0:3], planes[i, 4:7] = pos, normal
            for k in range(3):
                planes[i, 8 + 4 * k: 11 + 4 * k] = cols[k]
            planes[i, 20:21].view(u32)[0] = mat
        verts = np.zeros((len(ref.vertices), 4), dtype=f32)
        verts[:, 0:3] = ref.vertices
        norms = np.zeros((len(ref.normals), 4), dtype=f32)
        norms[:, 0:3] = ref.normals
        tris = np.array([list(t) for t in ref.triangles], dtype=u32).reshape(-1, 7)
        prims = np.array([list(p) for p in ordered], dtype=u32).reshape(-1, 2)
        bvh = np.zeros((len(nodes), 12), dtype=f32)
        for i, (mn, mx, idx, length, axis) in enumerate(nodes):
            bvh[i, 0:3], bvh[i, 4:7] = mn, mx
            bvh[i, 8:11].view(u32)[:] = (idx, length, axis)
        self._arrays = [np.ascontiguousarray(a) for a in (mats, spheres, planes, verts, norms, tris, prims, bvh)]
        d = _ffi.RtSceneDesc()
        names = ("materials", "spheres", "planes", "vertices", "normals", "triangles", "primitives", "bvh_nodes")
        counts = ("material_count", "sphere_count", "plane_count", "vertex_count", "normal_count",
                  "triangle_count", "primitive_count", "bvh_node_count")
        for a, n, c in zip(self._arrays, names, counts):
            setattr(d, n, a.ctypes.data if len(a) else None)
            setattr(d, c, len(a))
        self.desc = d
        self.camera = _Camera(*ref.camera)


class _Env:
    def __init__(self, rgb):
        self.height, self.width = rgb.shape[0], rgb.shape[1]
        self.rgba = np.zeros((self.height, self.width, 4), dtype=f32)
        self.rgba[..., 0:3] = rgb
        prob, alias, pmf, self.leftover_count = scene_oracle.alias_table(rgb)
        self.alias = np.zeros((self.width * self.height, 4), dtype=np.uint32)
        self.alias[:, 0] = prob.view(np.uint32)
        self.alias[:, 1] = alias
        self.alias[:, 2] = pmf.view(np.uint32)

    def desc(self):
        return _ffi.RtEnvironmentDesc(self.width, self.height, self.rgba.ctypes.data, self.alias.ctypes.data)


class IndependentEnvironments:
    """Duck-types api.EnvironmentMaps for oracle.Oracle, with the alias table built by the pure
    Python loop of scene_oracle.alias_table (keep the maps small: it is O(texels) in Python)."""

    def __init__(self, rgbs):
        self.environments = [_Env(np.ascontiguousarray(r, dtype=f32)) for r in rgbs]

    def descs(self):
        arr = (_ffi.RtEnvironmentDesc * len(self.environments))()
        for i, e in enumerate(self.environments):
            arr[i] = e.desc()
        return arr
