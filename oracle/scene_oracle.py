"""scene_oracle.py — CPU ORACLE (numpy float32) for the host-side builders. TEST
INFRASTRUCTURE ONLY: only tests/ may import this.

Independent restatement of the reference's CPU code that feeds the hot path:
  src/scene.rs:60-141   Bounds3 (from_points / from_bounds / center / max_axis / surface_area)
  src/scene.rs:166-207  Sphere / Plane bounds, Plane::to_uniform
  src/scene.rs:264-441  scene TOML schema + build_scene
  src/mesh.rs:24-113    Mesh::load (over wavefront_obj 11.0.0 semantics) + pack_meshes
  src/bvh.rs:13-337     build_bvh (binned SAH, flatten)
  src/environments.rs:87-186  AliasTable::build_by_luminance
  src/camera.rs:26-28,103-119 rot_transform / CameraUniform

Every arithmetic step is a single np.float32 operation in the reference's order.
PARITY UNPINNED: the reference has no tests or golden vectors and cannot run here; the
only external pins are the BVH dumps of SURVEY.md Appendix B (tests/golden/).
"""
from __future__ import annotations

import math
import os
import tomllib

import numpy as np

f32 = np.float32
F32_MAX = f32(3.4028234663852886e38)


def v3(x, y, z):
    return np.array([x, y, z], dtype=f32)


# ---------------------------------------------------------------- Bounds3 (scene.rs:60-141)
class Bounds3:
    __slots__ = ("min", "max")

    def __init__(self, mn, mx):
        self.min, self.max = mn, mx

    @staticmethod
    def identity():
        return Bounds3(v3(F32_MAX, F32_MAX, F32_MAX), v3(-F32_MAX, -F32_MAX, -F32_MAX))

    @staticmethod
    def from_points(points):
        b = Bounds3.identity()
        for p in points:
            b = Bounds3(np.minimum(b.min, p), np.maximum(b.max, p))
        return b

    def union(self, o):
        return Bounds3(np.minimum(self.min, o.min), np.maximum(self.max, o.max))

    def center(self):
        return self.min * f32(0.5) + self.max * f32(0.5)

    def max_axis(self):
        d = self.max - self.min
        if d[2] > d[0] and d[2] > d[1]:
            return 2
        if d[1] > d[0]:
            return 1
        return 0

    def surface_area(self):
        d = self.max - self.min
        return f32(2.0) * (d[0] * d[1] + d[0] * d[2] + d[1] * d[2])


# ---------------------------------------------------------------- OBJ (mesh.rs + wavefront_obj)
# wavefront_obj 11.0.0 is not under /root/reference; its fan triangulation is restated from memory as
# anchored at the LAST corner. tools/fan_anchor_report.py flips this switch to measure what would
# change if the crate anchored at the first corner instead (tests/golden/fan_anchor_report.json).
FAN_ANCHOR = "last"


def load_obj(text: str):
    """Returns (vertices [n,3] f32, normals [m,3] f32, triangles [(v0,v1,v2,n0,n1,n2)])."""
    verts, norms, faces = [], [], []
    for raw in text.splitlines():
        line = raw.split("#", 1)[0].split()
        if not line:
            continue
        if line[0] == "v":
            verts.append([f32(float(t)) for t in line[1:4]])  # f64 parse, cast to f32
        elif line[0] == "vn":
            norms.append([f32(float(t)) for t in line[1:4]])
        elif line[0] == "f":
            corners = []
            for tok in line[1:]:
                parts = tok.split("/")
                vi = int(parts[0])
                ni = int(parts[2]) if len(parts) > 2 and parts[2] else None
                corners.append((vi, ni))
            faces.append(corners)
    tris = []
    for corners in faces:
        n = len(corners)
        if n < 3:
            continue
        for i in range(n - 2):
            if FAN_ANCHOR == "last":  # fan anchored at the LAST corner (wavefront_obj)
                a, b, c = corners[-1], corners[i], corners[i + 1]
            else:
                a, b, c = corners[0], corners[i + 1], corners[i + 2]
            if a[1] is None or b[1] is None or c[1] is None:
                raise ValueError("Object must include baked normals")

            def res(k, count):
                return k - 1 if k > 0 else count + k
            tris.append((res(a[0], len(verts)), res(b[0], len(verts)), res(c[0], len(verts)),
                         res(a[1], len(norms)), res(b[1], len(norms)), res(c[1], len(norms))))
    return (np.array(verts, dtype=f32).reshape(-1, 3), np.array(norms, dtype=f32).reshape(-1, 3),
            tris)


# ---------------------------------------------------------------- scene TOML (scene.rs:264-441)
class SceneData:
    def __init__(self):
        self.materials = []   # (color[3], roughness, metallic, emission[3]) as f32
        self.spheres = []     # (pos, radius, material_id)
        self.planes = []      # (pos, forward, right, material_id)
        self.vertices = np.zeros((0, 3), dtype=f32)
        self.normals = np.zeros((0, 3), dtype=f32)
        self.triangles = []   # (v0, v1, v2, n0, n1, n2, material_id) with global indices
        self.camera = None    # (pos, yaw, pitch, fov_y) radians


def load_scene_toml(path: str) -> SceneData:
    with open(path, "rb") as fh:
        doc = tomllib.load(fh)
    s = SceneData()
    names = []
    for m in doc["material"]:
        names.append(m["name"])
        s.materials.append((v3(*m["color"]), f32(m["roughness"]), f32(m["metallic"]),
                            v3(*m["emission"])))
    base = os.path.dirname(path) or "."
    for i, obj in enumerate(doc["object"]):
        (kind, body), = obj.items()
        if body["material"] not in names:
            raise ValueError(f"Error in object {i} ({kind}): Material '{body['material']}' does not exist.")
        mat = names.index(body["material"])
        if kind == "Sphere":
            s.spheres.append((v3(*body["pos"]), f32(body["radius"]), mat))
        elif kind == "Plane":
            s.planes.append((v3(*body["pos"]), v3(*body["forward"]), v3(*body["right"]), mat))
        elif kind == "Mesh":
            with open(os.path.join(base, body["path"])) as fh:
                verts, norms, tris = load_obj(fh.read())
            vo, no = len(s.vertices), len(s.normals)  # pack_meshes (mesh.rs:92-113)
            s.triangles += [(a + vo, b + vo, c + vo, x + no, y + no, z + no, mat)
                            for a, b, c, x, y, z in tris]
            s.vertices = np.concatenate([s.vertices, verts])
            s.normals = np.concatenate([s.normals, norms])
        else:
            raise ValueError(f"unknown variant {kind}")
    cam = doc["camera"]
    k = f32(math.pi / 180.0)  # cgmath Deg -> Rad
    s.camera = (v3(*cam["pos"]), f32(cam["yaw"]) * k, f32(cam["pitch"]) * k, f32(cam["fov_y"]) * k)
    return s


# ---------------------------------------------------------------- BVH (bvh.rs)
MAX_PRIMS_PER_LEAF = 5
BUCKETS = 12


def primitive_bounds(scene: SceneData):
    """PrimitiveInfo::build_array order: spheres, planes, triangles (bvh.rs:40-72)."""
    prims = []
    for i, (pos, r, _m) in enumerate(scene.spheres):
        rr = v3(r, r, r)
        prims.append((0, i, Bounds3(pos - rr, pos + rr)))
    for i, (pos, fwd, right, _m) in enumerate(scene.planes):
        prims.append((1, i, Bounds3.from_points([pos, pos + fwd + right])))
    for i, t in enumerate(scene.triangles):
        prims.append((2, i, Bounds3.from_points([scene.vertices[t[0]], scene.vertices[t[1]],
                                                 scene.vertices[t[2]]])))
    return prims


def build_bvh(scene: SceneData):
    """Returns (nodes, ordered) where nodes = [(min, max, idx, len, axis)] in flattened
    pre-order and ordered = [(type, index)] by leaf slot."""
    prims = [(t, i, b, b.center()) for t, i, b in primitive_bounds(scene)]
    nodes, ordered = [], []

    def build(items):
        bounds = Bounds3.identity()
        for it in items:
            bounds = bounds.union(it[2])

        def leaf():
            first = len(ordered)
            ordered.extend((it[0], it[1]) for it in items)
            nodes.append((bounds.min, bounds.max, first, len(items), 0))
            return bounds

        if len(items) <= MAX_PRIMS_PER_LEAF:
            return leaf()
        cb = Bounds3.from_points([it[3] for it in items])
        axis = cb.max_axis()
        cmin, cmax = cb.min[axis], cb.max[axis]
        if cmin == cmax:
            return leaf()

        def bucket_of(it):
            v = f32(BUCKETS) * ((it[3][axis] - cmin) / (cmax - cmin))
            b = int(v) if v > 0 else 0
            return min(b, BUCKETS - 1)

        counts = [0] * BUCKETS
        bbs = [Bounds3.identity() for _ in range(BUCKETS)]
        for it in items:
            b = bucket_of(it)
            counts[b] += 1
            bbs[b] = bbs[b].union(it[2])
        total = bounds.surface_area()
        costs = []
        for split in range(BUCKETS - 1):
            b0, b1 = Bounds3.identity(), Bounds3.identity()
            c0 = c1 = 0
            for b in range(split + 1):
                b0 = b0.union(bbs[b])
                c0 += counts[b]
            for b in range(split + 1, BUCKETS):
                b1 = b1.union(bbs[b])
                c1 += counts[b]
            with np.errstate(over="ignore", invalid="ignore"):
                costs.append(f32(0.125) + (f32(c0) * b0.surface_area() + f32(c1) * b1.surface_area()) / total)
        best = 0
        for i in range(1, len(costs)):
            if costs[i] < costs[best]:
                best = i
        # swap-to-end partition (bvh.rs:304-315)
        items = list(items)
        split, end = 0, len(items)
        while split < end:
            if bucket_of(items[split]) <= best:
                split += 1
            else:
                end -= 1
                items[split], items[end] = items[end], items[split]
        assert 0 < split < len(items), "median-split fallback is unreachable"
        me = len(nodes)
        nodes.append(None)
        lb = build(items[:split])
        second = len(nodes)
        rb = build(items[split:])
        nb = Bounds3.identity().union(lb).union(rb)
        nodes[me] = (nb.min, nb.max, second, 0, axis)
        return nb

    build(prims)
    return nodes, ordered


# ---------------------------------------------------------------- Plane::to_uniform (scene.rs:190-201)
def _cross(a, b):
    return v3(a[1] * b[2] - b[1] * a[2], a[2] * b[0] - b[2] * a[0], a[0] * b[1] - b[0] * a[1])


def _dot(a, b):
    return (a[0] * b[0] + a[1] * b[1]) + a[2] * b[2]


def plane_to_uniform(pos, forward, right):
    """Returns (normal, base_change columns [3][3])."""
    c = _cross(forward, right)
    normal = c * (f32(1.0) / np.sqrt(_dot(c, c)))
    x, y, z = right, normal, forward
    t0, t1, t2 = _cross(y, z), _cross(z, x), _cross(x, y)
    inv_det = f32(1.0) / _dot(z, t2)
    r0, r1, r2 = t0 * inv_det, t1 * inv_det, t2 * inv_det
    cols = [v3(r0[k], r1[k], r2[k]) for k in range(3)]
    return normal, cols


# ---------------------------------------------------------------- alias table (environments.rs:97-186)
def alias_table(rgb: np.ndarray):
    """rgb [H, W, 3] f32 -> (probability f32[N], alias_index u32[N], pmf f32[N], leftover)."""
    h, w, _ = rgb.shape
    n = w * h
    pi = f32(math.pi)
    rows = np.arange(h, dtype=f32)
    angle = pi * (rows + f32(0.5)) / f32(h)
    row_sin = np.array([f32(math.sin(float(a))) for a in angle], dtype=f32)
    lum = (f32(0.2126) * rgb[..., 0] + f32(0.7152) * rgb[..., 1]) + f32(0.0722) * rgb[..., 2]
    weights = (lum * row_sin[:, None]).astype(f32).reshape(-1)
    total = np.cumsum(weights, dtype=f32)[-1]  # sequential f32 sum
    prob = (weights * f32(n)) / total
    alias_prob = prob.copy()
    small = [i for i in range(n) if prob[i] < 1.0]
    large = [i for i in range(n) if prob[i] >= 1.0]
    out_p = np.ones(n, dtype=f32)
    out_a = np.arange(n, dtype=np.uint32)
    out_pmf = np.full(n, f32(1.0) / f32(n), dtype=f32)
    assigned = np.zeros(n, dtype=bool)
    one = f32(1.0)
    while small:
        s = small.pop()
        if not large:
            break
        l = large.pop()
        out_p[s] = alias_prob[s]
        out_a[s] = l
        out_pmf[s] = prob[s] / f32(n)
        assigned[s] = True
        alias_prob[l] = alias_prob[l] - (one - alias_prob[s])
        (small if alias_prob[l] < 1.0 else large).append(l)
    return out_p, out_a, out_pmf, int((~assigned).sum())


# ---------------------------------------------------------------- camera (camera.rs:26-28)
def rot_transform(yaw, pitch):
    """R_y(yaw) * R_x(pitch) as 3 columns, glam Mat3::from_axis_angle; sin/cos in f64->f32."""
    def axis_angle(axis, angle):
        s, c = f32(math.sin(float(angle))), f32(math.cos(float(angle)))
        x, y, z = axis
        omc = f32(1.0) - c
        xy, xz, yz = x * y * omc, x * z * omc, y * z * omc
        return [v3(x * x * omc + c, xy + z * s, xz - y * s),
                v3(xy - z * s, y * y * omc + c, yz + x * s),
                v3(xz + y * s, yz - x * s, z * z * omc + c)]

    def mul_vec(m, v):
        return (m[0] * v[0] + m[1] * v[1]) + m[2] * v[2]

    ry = axis_angle(v3(0, 1, 0), f32(yaw))
    rx = axis_angle(v3(1, 0, 0), f32(pitch))
    return [mul_vec(ry, rx[k]) for k in range(3)]
