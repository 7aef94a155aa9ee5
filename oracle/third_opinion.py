"""A third opinion on `main` / `trace_ray`: shader.wgsl restated once more, in numpy float32 scalars.

TEST INFRASTRUCTURE ONLY. The C++ oracle (shader_oracle.cpp) is what the CUDA backend is compared
with; nothing in the reference pins it (PARITY UNPINNED). This module is a second, independently
written reading of the same WGSL — every function below was written from
/root/reference/src/shaders/shader.wgsl (cited by line), on scene arrays that come from the numpy
front-end (scene_oracle.py), not from the product's host library — so that a transcription slip in
either restatement shows up as a bit difference on a handful of pixel-samples
(tests/test_oracle_shader.py::test_third_opinion_*).

Shared with the oracle on purpose: only the arithmetic contract of DESIGN.md section 2 — one IEEE
binary32 operation per WGSL operator in source order, dot = (x*x' + y*y') + z*z', normalize =
v * (1 / length), reflect = i - (2 * dot(n, i)) * n, abstract-float constants folded in f64 and
rounded once — and the four transcendental polynomials (sin, cos, atan2, asin), which are a choice of
this repository rather than a reading of the shader and are called through liboracle.so.
"""
from __future__ import annotations

import ctypes as C
import math
import os
import sys

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
for _p in (os.path.dirname(_HERE), _HERE):
    if _p not in sys.path:
        sys.path.insert(0, _p)
import oracle as _orc  # noqa: E402  (sin / cos / atan2 / asin only)
import scene_oracle  # noqa: E402

f32 = np.float32
INFINITY = f32(1.70141183460469231732e+38)      # shader.wgsl:235
PI = f32(3.14159)                                # shader.wgsl:240
INV_PI = f32(1.0 / 3.14159)                      # shader.wgsl:241
TWO_PI = f32(2.0 * 3.14159)                      # `2 * PI`, `2. * PI`: constant expressions
ONE_OVER_PI = f32(1.0 / 3.14159)                 # `1 / PI` (shader.wgsl:1082)
ZERO, ONE, TWO, HALF = f32(0.0), f32(1.0), f32(2.0), f32(0.5)


def _lib():
    return _orc.lib()


def sin(x):
    return f32(_lib().oracle_sin(float(x)))


def cos(x):
    return f32(_lib().oracle_cos(float(x)))


def asin(x):
    return f32(_lib().oracle_asin(float(x)))


def atan2(y, x):
    return f32(_lib().oracle_atan2(float(y), float(x)))


def sqrt(x):
    return f32(np.sqrt(f32(x)))


# ---- vec3 as tuples of float32 ----
def v3(x, y, z):
    return (f32(x), f32(y), f32(z))


def add(a, b):
    return (a[0] + b[0], a[1] + b[1], a[2] + b[2])


def sub(a, b):
    return (a[0] - b[0], a[1] - b[1], a[2] - b[2])


def mul(a, b):  # component-wise
    return (a[0] * b[0], a[1] * b[1], a[2] * b[2])


def scale(a, s):
    return (a[0] * s, a[1] * s, a[2] * s)


def div(a, s):
    return (a[0] / s, a[1] / s, a[2] / s)


def neg(a):
    return (-a[0], -a[1], -a[2])


def dot(a, b):
    return (a[0] * b[0] + a[1] * b[1]) + a[2] * b[2]


def cross(a, b):
    return (a[1] * b[2] - a[2] * b[1], a[2] * b[0] - a[0] * b[2], a[0] * b[1] - a[1] * b[0])


def length(a):
    return sqrt(dot(a, a))


def normalize(a):
    return scale(a, ONE / length(a))


def reflect(i, n):
    return sub(i, scale(n, TWO * dot(n, i)))


def saturate(x):
    return min(max(f32(x), ZERO), ONE)


def mat_vec(cols, v):
    """mat3x3 * vec3 with the matrix given by columns."""
    return add(add(scale(cols[0], v[0]), scale(cols[1], v[1])), scale(cols[2], v[2]))


# ---- RNG (shader.wgsl:605-631) ----
class Rng:
    def __init__(self, pixel_index, sample_count):
        self.s = 0
        self.salt(pixel_index)      # shader.wgsl:1311
        self.salt(sample_count)     # shader.wgsl:1312

    def salt(self, v):
        self.s ^= v & 0xFFFFFFFF
        self.next_u32()

    def next_u32(self):
        self.s = (self.s * 747796405 + 2891336453) & 0xFFFFFFFF
        r = (((self.s >> ((self.s >> 28) + 4)) ^ self.s) * 277803737) & 0xFFFFFFFF
        return ((r >> 22) ^ r) & 0xFFFFFFFF

    def uniform(self):
        # f32(u32) / 4294967295.0: the literal rounds to 2^32 in f32, so the division is exact
        return f32(self.next_u32()) / f32(4294967296.0)


class World:
    """Scene, camera and environment 0 of one TOML, built by the numpy front-end."""

    def __init__(self, toml_path, env_rgb, width, height, max_bounces):
        ref = scene_oracle.load_scene_toml(toml_path)
        self.nodes, self.prims = scene_oracle.build_bvh(ref)
        self.spheres = [(tuple(p), f32(r), m) for p, r, m in ref.spheres]
        self.planes = []
        for pos, fwd, right, m in ref.planes:
            normal, cols = scene_oracle.plane_to_uniform(pos, fwd, right)
            self.planes.append((tuple(pos), tuple(normal), [tuple(c) for c in cols], m))
        self.vertices = [tuple(v) for v in ref.vertices]
        self.normals = [tuple(n) for n in ref.normals]
        self.triangles = ref.triangles
        self.materials = [(tuple(c), f32(r), f32(mt), tuple(e)) for c, r, mt, e in ref.materials]
        pos, yaw, pitch, fov = ref.camera
        self.cam_pos, self.cam_rot, self.fov_y = tuple(pos), [tuple(c) for c in scene_oracle.rot_transform(yaw, pitch)], f32(fov)
        self.env = np.ascontiguousarray(env_rgb, dtype=f32)
        self.env_h, self.env_w = self.env.shape[0], self.env.shape[1]
        self.alias_p, self.alias_i, self.alias_pmf, _ = scene_oracle.alias_table(self.env)
        self.width, self.height, self.max_bounces = width, height, max_bounces

    # ---- intersections ----
    def hit_sphere(self, o, d, sphere):            # shader.wgsl:295-360
        pos, radius, mat = sphere
        eps = f32(1.0e-4)
        l = sub(o, pos)
        a = dot(d, d)
        b = TWO * dot(d, l)
        c = dot(l, l) - radius * radius
        disc = b * b - f32(4.0) * a * c
        if disc < 0:
            return None
        if disc == 0:
            t = f32(-0.5) * b / a
        else:
            q = f32(-0.5) * (b + sqrt(disc)) if b > 0 else f32(-0.5) * (b - sqrt(disc))
            t0, t1 = q / a, c / q
            if t0 < eps:
                t = t1
            elif t1 < eps:
                t = t0
            else:
                t = min(t0, t1)
        if t < eps:
            return None
        point = add(o, scale(d, t))
        n = normalize(sub(point, pos))
        back = sub(pos, o)
        if dot(back, back) - radius * radius < f32(1.0e-6):
            n = scale(n, f32(-1.0))
        return (t, point, n, mat)

    def hit_plane(self, o, d, plane):              # shader.wgsl:362-406
        pos, normal, cols, mat = plane
        den = dot(normal, d)
        if abs(den) < f32(0.0001):
            return None
        t = dot(normal, sub(pos, o)) / den
        if t < f32(0.001):
            return None
        inter = add(o, scale(d, t))
        q = mat_vec(cols, sub(inter, pos))
        if q[0] < 0 or ONE < q[0] or q[2] < 0 or ONE < q[2]:
            return None
        n = normal
        if dot(o, n) < 0:
            n = scale(n, f32(-1.0))
        return (t, inter, n, mat)

    def hit_triangle(self, o, d, tri):             # shader.wgsl:409-466
        a, b, c = self.vertices[tri[0]], self.vertices[tri[1]], self.vertices[tri[2]]
        e0, e1 = sub(b, a), sub(c, a)
        p0 = cross(sub(o, a), e0)
        p1 = cross(d, e1)
        det = dot(e0, p1)
        with np.errstate(divide="ignore"):
            inv = ONE / det
        if abs(det) < f32(1.0e-8):
            return None
        u = dot(sub(o, a), p1) * inv
        v = dot(d, p0) * inv
        if u < 0 or ONE < u:
            return None
        if v < 0 or ONE < (u + v):
            return None
        t = dot(e1, p0) * inv
        if t < f32(1.0e-5):
            return None
        n0, n1, n2 = self.normals[tri[3]], self.normals[tri[4]], self.normals[tri[5]]
        n = normalize(add(add(scale(n0, ONE - u - v), scale(n1, u)), scale(n2, v)))
        if dot(n, d) > 0:
            n = scale(n, f32(-1.0))
        return (t, add(o, scale(d, t)), n, tri[6])

    def hit_primitive(self, o, d, kind, index):
        if kind == 0:
            return self.hit_sphere(o, d, self.spheres[index])
        if kind == 1:
            return self.hit_plane(o, d, self.planes[index])
        return self.hit_triangle(o, d, self.triangles[index])

    @staticmethod
    def box_hit(o, inv, mn, mx):                   # shader.wgsl:262-293
        t0, t1 = ZERO, INFINITY
        for axis in range(3):
            with np.errstate(invalid="ignore", over="ignore"):
                tn = (mn[axis] - o[axis]) * inv[axis]
                tf = (mx[axis] - o[axis]) * inv[axis]
            if tn > tf:
                tn, tf = tf, tn
            if tn > t0:
                t0 = tn
            if tf < t1:
                t1 = tf
            if t0 > t1:
                return False
        return True

    def cast_ray_bvh(self, o, d):                  # shader.wgsl:469-564
        with np.errstate(divide="ignore"):
            inv = (ONE / d[0], ONE / d[1], ONE / d[2])
        best, best_t = None, INFINITY
        stack, cur = [], 0
        while True:
            mn, mx, idx, count, axis = self.nodes[cur]
            if self.box_hit(o, inv, mn, mx):
                if count > 0:
                    for i in range(count):
                        kind, index = self.prims[idx + i]
                        h = self.hit_primitive(o, d, kind, index)
                        if h is not None and h[0] < best_t:
                            best, best_t = h, h[0]
                    if not stack:
                        break
                    cur = stack.pop()
                elif inv[axis] < 0:
                    stack.append(cur + 1)
                    cur = idx
                else:
                    stack.append(idx)
                    cur = cur + 1
            else:
                if not stack:
                    break
                cur = stack.pop()
        return best

    def cast_ray(self, o, d):                      # shader.wgsl:567-601
        h = self.cast_ray_bvh(o, d)
        if h is not None:
            return h
        best, best_t = None, INFINITY
        for s in self.spheres:
            h = self.hit_sphere(o, d, s)
            if h is not None and h[0] < best_t:
                best, best_t = h, h[0]
        for p in self.planes:
            h = self.hit_plane(o, d, p)
            if h is not None and h[0] < best_t:
                best, best_t = h, h[0]
        return best

    # ---- environment (shader.wgsl:667-831) ----
    def texel(self, x, y):
        x = min(max(int(x), 0), self.env_w - 1)    # clamp-to-edge (state.rs:134-142)
        y = min(max(int(y), 0), self.env_h - 1)
        t = self.env[y, x]
        return (t[0], t[1], t[2])

    def sample_texture(self, u, v):
        """textureSampleLevel(.., uv, 0) under a clamp-to-edge, linear sampler: bilinear blend."""
        x = u * f32(self.env_w) - HALF
        y = v * f32(self.env_h) - HALF
        x0, y0 = f32(np.floor(x)), f32(np.floor(y))
        wx, wy = x - x0, y - y0
        top = add(scale(self.texel(x0, y0), ONE - wx), scale(self.texel(x0 + ONE, y0), wx))
        bot = add(scale(self.texel(x0, y0 + ONE), ONE - wx), scale(self.texel(x0 + ONE, y0 + ONE), wx))
        return add(scale(top, ONE - wy), scale(bot, wy))

    @staticmethod
    def direction_to_uv(d):                        # shader.wgsl:710-714
        return atan2(d[2], d[0]) * INV_PI * HALF + HALF, HALF - asin(d[1]) * INV_PI

    def pixel_solid_angle(self, v):                # shader.wgsl:739-749
        sin_t = max(f32(1.0e-6), sin(PI * v))
        return (TWO_PI / f32(self.env_w)) * (PI / f32(self.env_h)) * sin_t

    def direction_pdf(self, d):                    # shader.wgsl:753-769
        u, v = self.direction_to_uv(d)
        x = min(int(u * f32(self.env_w)), self.env_w - 1)
        y = min(int(v * f32(self.env_h)), self.env_h - 1)
        return self.alias_pmf[x + y * self.env_w] / self.pixel_solid_angle(v)

    def sample_environment(self, rng):             # shader.wgsl:689-706, 782-820
        n = self.env_w * self.env_h
        index = min(int(rng.uniform() * f32(n)), n - 1)
        pick = index if rng.uniform() < self.alias_p[index] else int(self.alias_i[index])
        x, y = pick % self.env_w, pick // self.env_w
        jx, jy = rng.uniform(), rng.uniform()
        u = (f32(x) + jx) / f32(self.env_w)
        v = (f32(y) + jy) / f32(self.env_h)
        phi = (TWO * u - ONE) * PI                 # shader.wgsl:718-732
        theta = PI * v
        st, ct = sin(theta), cos(theta)
        direction = (st * cos(phi), ct, st * sin(phi))
        radiance = self.sample_texture(u, v)
        pdf = self.alias_pmf[pick] / self.pixel_solid_angle(v)
        return direction, radiance, pdf

    # ---- BSDF (shader.wgsl:834-1210) ----
    @staticmethod
    def make_frame(n):                             # shader.wgsl:55-69
        helper = v3(0, 0, 1) if abs(n[2]) < f32(0.999) else v3(1, 0, 0)
        tangent = normalize(cross(helper, n))
        return tangent, cross(n, tangent), n

    @staticmethod
    def to_local(frame, w):
        return (dot(w, frame[0]), dot(w, frame[1]), dot(w, frame[2]))

    @staticmethod
    def to_world(frame, w):
        return normalize(add(add(scale(frame[0], w[0]), scale(frame[1], w[1])), scale(frame[2], w[2])))

    @staticmethod
    def luminance(c):
        return f32(0.2126) * c[0] + f32(0.7152) * c[1] + f32(0.0722) * c[2]

    @staticmethod
    def bsdf_material(material):                   # shader.wgsl:850-873
        color, roughness, metallic, emission = material
        alpha = max(f32(0.001), roughness * roughness)
        t = saturate(metallic)
        f0 = add(scale(v3(0.04, 0.04, 0.04), ONE - t), scale(color, t))
        return color, metallic, alpha, f0, emission

    @staticmethod
    def d_ggx(ndh, alpha):                         # shader.wgsl:924-928
        a2 = alpha * alpha
        den = (ndh * ndh) * (a2 - ONE) + ONE
        return a2 / (PI * den * den)

    @staticmethod
    def lambda_ggx(ndv, alpha):                    # shader.wgsl:1014-1020
        n2 = ndv * ndv
        return (sqrt(ONE + alpha * alpha * (ONE - n2) / n2) - ONE) / TWO

    def g1_ggx(self, ndv, alpha):
        return ONE / (ONE + self.lambda_ggx(ndv, alpha))

    def bsdf_eval(self, wo, wi, m):                # shader.wgsl:1053-1087
        color, metallic, alpha, f0, _ = m
        if wo[2] <= 0 or wi[2] <= 0:
            return v3(0, 0, 0)
        h = normalize(add(wo, wi))
        ndh = saturate(h[2])
        d = self.d_ggx(ndh, alpha)
        g = self.g1_ggx(wo[2], alpha) * self.g1_ggx(wi[2], alpha)
        x = ONE - saturate(dot(h, wo))            # f_schlick, shader.wgsl:1045-1051
        x2 = x * x
        x5 = x2 * x2 * x
        fres = add(f0, scale(sub(v3(1, 1, 1), f0), x5))
        fs = scale(fres, (d * g) / (f32(4.0) * wo[2] * wi[2]))
        kd0 = scale(color, ONE - saturate(metallic))
        kd = scale(kd0, ONE - max(f0[0], max(f0[1], f0[2])))
        return add(scale(kd, ONE_OVER_PI), fs)

    def bsdf_pdf(self, wo, wi, m):                 # shader.wgsl:1089-1114
        _, _, alpha, f0, _ = m
        if wo[2] <= 0 or wi[2] <= 0:
            return ZERO
        p_spec = saturate(self.luminance(f0))
        p_diff = ONE - p_spec
        pdf_cos = ZERO if wi[2] <= 0 else wi[2] / PI
        # pdf_specular_wi_visible
        h = normalize(add(wo, wi))
        wo_dot_h = abs(dot(wo, h))
        if wo_dot_h <= 0:
            pdf_spec = ZERO
        else:
            if h[2] <= 0:
                half_pdf = ZERO
            else:
                half_pdf = self.d_ggx(h[2], alpha) * self.g1_ggx(wo[2], alpha) * max(ZERO, dot(wo, h)) / wo[2]
            pdf_spec = half_pdf / (f32(4.0) * wo_dot_h)
        return p_diff * pdf_cos + p_spec * pdf_spec

    @staticmethod
    def sample_ggx_half_vector(s0, s1, wo, alpha):  # shader.wgsl:962-1009
        vs = normalize(mul(wo, (alpha, alpha, ONE)))
        len2 = vs[0] * vs[0] + vs[1] * vs[1]
        if len2 > 0:
            tx = scale((-vs[1], vs[0], ZERO), ONE / sqrt(len2))
        else:
            tx = v3(1, 0, 0)
        ty = cross(vs, tx)
        radius = sqrt(s0)                          # sample_uniform_disk
        azimuth = TWO_PI * s1
        px, py = radius * cos(azimuth), radius * sin(azimuth)
        a = sqrt(max(ZERO, ONE - px * px))
        py = (ONE - vs[2]) * a + vs[2] * py        # lerp_f32(a, py, vs.z)
        hs = add(add(scale(tx, px), scale(ty, py)), scale(vs, sqrt(max(ZERO, ONE - px * px - py * py))))
        return normalize((alpha * hs[0], alpha * hs[1], max(ZERO, hs[2])))

    def bsdf_sample(self, d, n, m, rng):           # shader.wgsl:1116-1202 -> (direction, scattering, pdf)
        _, _, alpha, f0, _ = m
        wo_world = neg(d)
        if dot(n, wo_world) <= 0:
            return v3(0, 0, 0), v3(0, 0, 1), ZERO
        frame = self.make_frame(n)
        wo = self.to_local(frame, wo_world)
        if wo[2] <= 0:
            return v3(0, 0, 0), v3(0, 1, 0), ZERO
        p_spec = saturate(self.luminance(f0))
        p_diff = ONE - p_spec
        s = rng.uniform()
        if s < p_diff:
            sx, sy = s / max(p_diff, f32(1.0e-6)), rng.uniform()
            r = sqrt(sx)                           # sample_cosine_hemisphere
            phi = TWO_PI * sy
            x, y = r * cos(phi), r * sin(phi)
            wi = (x, y, sqrt(max(ZERO, ONE - x * x - y * y)))
        else:
            sx, sy = (s - p_diff) / max(p_spec, f32(1.0e-6)), rng.uniform()
            h = self.sample_ggx_half_vector(sx, sy, wo, alpha)
            wi = reflect(neg(wo), h)
            if wi[2] <= 0:
                return v3(1, 0, 0), v3(1, 0, 0), ZERO
        scattering = self.bsdf_eval(wo, wi, m)
        pdf = self.bsdf_pdf(wo, wi, m)
        wi_world = self.to_world(frame, wi)
        if dot(n, wi_world) < 0:
            return v3(0, 0, 0), v3(0, 1, 0), ZERO
        return wi_world, scattering, pdf

    # ---- trace_ray (shader.wgsl:1213-1303) ----
    def trace_ray(self, o, d, rng):
        light, throughput, last_pdf = v3(0, 0, 0), v3(1, 1, 1), ONE
        for _ in range(self.max_bounces):
            info = self.cast_ray(o, d)
            if info is None:
                u, v = self.direction_to_uv(d)
                sky = self.sample_texture(u, v)
                pdf = self.direction_pdf(d)
                w = (last_pdf * last_pdf) / (last_pdf * last_pdf + pdf * pdf)
                light = add(light, scale(mul(throughput, sky), w))
                break
            _t, point, n, mat = info
            m = self.bsdf_material(self.materials[mat])
            light = add(light, mul(throughput, m[4]))
            e_dir, e_rad, e_pdf = self.sample_environment(rng)
            cos_theta = max(ZERO, dot(n, e_dir))
            if cos_theta > 0 and e_pdf > 0 and self.cast_ray_bvh(point, e_dir) is None:
                frame = self.make_frame(n)
                wo, wi = self.to_local(frame, neg(d)), self.to_local(frame, e_dir)
                f = self.bsdf_eval(wo, wi, m)
                pdf_bsdf = self.bsdf_pdf(wo, wi, m)
                w = (e_pdf * e_pdf) / (e_pdf * e_pdf + pdf_bsdf * pdf_bsdf)
                light = add(light, div(scale(mul(mul(scale(throughput, w), e_rad), f), cos_theta), e_pdf))
            s_dir, s_f, s_pdf = self.bsdf_sample(d, n, m, rng)
            if s_dir[0] == 0 and s_dir[1] == 0 and s_dir[2] == 0:
                light = s_f
                break
            if s_pdf <= 0:
                break
            cos_theta = max(ZERO, dot(n, s_dir))
            throughput = mul(throughput, scale(s_f, cos_theta / s_pdf))
            if length(throughput) < f32(0.001):
                break
            last_pdf = s_pdf
            o, d = point, s_dir
        return light

    # ---- main (shader.wgsl:1305-1373), without the accumulation ----
    def camera_ray(self, px, py, sample):
        rng = Rng(py * self.width + px, sample)
        angle = rng.uniform() * TWO * f32(3.1415926)          # random_in_circle_uniform
        cx, sy = cos(angle), sin(angle)
        r = sqrt(rng.uniform())
        jx, jy = f32(px) + cx * r, f32(py) + sy * r
        sx = ((jx / f32(self.width)) * TWO - ONE) * ONE
        sy2 = ((jy / f32(self.height)) * TWO - ONE) * f32(-1.0)
        m = sin(self.fov_y / TWO)
        aspect = f32(self.width) / f32(self.height)
        cam = (sx * m * aspect, sy2 * m, f32(-1.0))
        return self.cam_pos, normalize(mat_vec(self.cam_rot, cam)), rng

    def pixel_sample(self, px, py, sample):
        o, d, rng = self.camera_ray(px, py, sample)
        return np.array(self.trace_ray(tuple(f32(c) for c in o), d, rng), dtype=f32)
