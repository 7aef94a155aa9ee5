// rt_device.cuh — device-side arithmetic of the path tracer for sm_100a.
//
// Semantics follow /root/reference/src/shaders/shader.wgsl; the data layout, traversal
// (closest-hit pruning, any-hit shadow rays, deferred hit reconstruction) and kernel
// structure are this backend's own. Every floating-point expression is written in
// the WGSL source's evaluation order as single IEEE binary32 operations and the file
// is compiled with --fmad=false, so results are reproducible bit for bit (DESIGN.md,
// "arithmetic contract").
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>
#include <string.h>

namespace rtd {

// RT_HOST_EMU (tests only) also compiles the ray-casting routines for the host, so the
// CPU test suite can exercise the kernels' traversal logic without a GPU (tests/emu/).
#ifdef RT_HOST_EMU
#define RT_DEV __host__ __device__ __forceinline__
#define RT_COLD __host__ __device__ __noinline__
#else
#define RT_DEV __device__ __forceinline__
#define RT_COLD __device__ __noinline__
#endif
#define RT_DEVONLY __device__ __forceinline__

RT_DEV uint32_t f_as_u(float f) {
#ifdef __CUDA_ARCH__
    return __float_as_uint(f);
#else
    uint32_t u;
    memcpy(&u, &f, 4);
    return u;
#endif
}
RT_DEV int f_as_i(float f) { return (int)f_as_u(f); }
RT_DEV float u_as_f(uint32_t u) {
#ifdef __CUDA_ARCH__
    return __uint_as_float(u);
#else
    float f;
    memcpy(&f, &u, 4);
    return f;
#endif
}

// ---------------------------------------------------------------------------------
// Small vector helpers (WGSL evaluation order).
// ---------------------------------------------------------------------------------
struct V3 {
    float x, y, z;
};
RT_DEV V3 v3(float x, float y, float z) { return V3{x, y, z}; }
RT_DEV V3 v3s(float s) { return V3{s, s, s}; }
RT_DEV V3 xyz(float4 v) { return V3{v.x, v.y, v.z}; }
RT_DEV V3 operator+(V3 a, V3 b) { return v3(a.x + b.x, a.y + b.y, a.z + b.z); }
RT_DEV V3 operator-(V3 a, V3 b) { return v3(a.x - b.x, a.y - b.y, a.z - b.z); }
RT_DEV V3 operator-(V3 a) { return v3(-a.x, -a.y, -a.z); }
RT_DEV V3 operator*(V3 a, V3 b) { return v3(a.x * b.x, a.y * b.y, a.z * b.z); }
RT_DEV V3 operator*(V3 a, float s) { return v3(a.x * s, a.y * s, a.z * s); }
RT_DEV V3 operator*(float s, V3 a) { return v3(s * a.x, s * a.y, s * a.z); }
RT_DEV V3 operator/(V3 a, float s) { return v3(a.x / s, a.y / s, a.z / s); }
RT_DEV float dot(V3 a, V3 b) { return (a.x * b.x + a.y * b.y) + a.z * b.z; }
RT_DEV V3 cross(V3 a, V3 b) {
    return v3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x);
}
RT_DEV float length(V3 v) { return sqrtf(dot(v, v)); }
RT_DEV V3 normalize(V3 v) { return v * (1.0f / length(v)); }
RT_DEV V3 reflect(V3 i, V3 n) { return i - (2.0f * dot(n, i)) * n; }
RT_DEV float saturate(float x) { return fminf(fmaxf(x, 0.0f), 1.0f); }
RT_DEV float comp(V3 v, uint32_t axis) { return axis == 0 ? v.x : (axis == 1 ? v.y : v.z); }

// ---------------------------------------------------------------------------------
// Constants (shader.wgsl:232-240). Abstract-float expressions fold in f64.
// ---------------------------------------------------------------------------------
// 1: fetch environment texels from linear memory instead of the texture object (measured
// equal on B200; setting L2 aside as persisting for them made the render 2x slower).
#ifndef RT_ENV_LINEAR
#define RT_ENV_LINEAR 0
#endif
#define RT_INFINITY 1.70141183460469231732e+38f
#define RT_PI ((float)3.14159)
#define RT_INV_PI ((float)(1.0 / 3.14159))
#define RT_TWO_PI ((float)(2.0 * 3.14159))

// ---------------------------------------------------------------------------------
// Transcendentals of the arithmetic contract: Cody-Waite reduction by pi/2 and
// Cephes single-precision polynomials, op for op.
// ---------------------------------------------------------------------------------
RT_DEV float sin_poly(float r) {
    float z = r * r;
    return r + (r * z) * ((-1.9515295891e-4f * z + 8.3321608736e-3f) * z - 1.6666654611e-1f);
}
RT_DEV float cos_poly(float r) {
    float z = r * r;
    return (1.0f - 0.5f * z) +
           (z * z) * ((2.443315711809948e-5f * z - 1.388731625493765e-3f) * z + 4.166664568298827e-2f);
}
RT_DEV void rt_sincos(float x, float* s_out, float* c_out) {
    float q = rintf(x * 0.636619772f);
    float r = ((x - q * 1.5703125f) - q * 4.837512969970703125e-4f) - q * 7.54978995489188216e-8f;
    int qi = (int)q;
    float sp = sin_poly(r), cp = cos_poly(r);
    // quadrant 0: (s, c) = (sp, cp); 1: (cp, -sp); 2: (-sp, -cp); 3: (-cp, sp)
    float s = (qi & 1) ? cp : sp;
    float c = (qi & 1) ? sp : cp;
    if (qi & 2) s = -s;
    if ((qi + 1) & 2) c = -c;
    *s_out = s;
    *c_out = c;
}
RT_DEV float rt_sin(float x) {
    float s, c;
    rt_sincos(x, &s, &c);
    return s;
}

RT_DEV float atan_pos(float t) {
    float y0 = 0.0f;
    if (t > 2.41421366f) {
        y0 = 1.57079637f;
        t = -(1.0f / t);
    } else if (t > 0.414213568f) {
        y0 = 0.785398185f;
        t = (t - 1.0f) / (t + 1.0f);
    }
    float z = t * t;
    float p = (((8.05374449538e-2f * z - 1.38776856032e-1f) * z + 1.99777106478e-1f) * z -
               3.33329491539e-1f) * z * t + t;
    return y0 + p;
}
RT_DEV float rt_atan2(float y, float x) {
    if (x == 0.0f) {
        if (y > 0.0f) return 1.57079637f;
        if (y < 0.0f) return -1.57079637f;
        return 0.0f;
    }
    float a = atan_pos(fabsf(y / x));
    if (x < 0.0f) a = 3.14159274f - a;
    return y < 0.0f ? -a : a;
}
RT_DEV float rt_asin(float x) {
    float a = fabsf(x);
    if (a > 1.0f) a = 1.0f;
    bool big = a > 0.5f;
    float z = big ? 0.5f * (1.0f - a) : a * a;
    float w = big ? sqrtf(z) : a;
    float p = ((((4.2163199048e-2f * z + 2.4181311049e-2f) * z + 4.5470025998e-2f) * z +
                7.4953002686e-2f) * z + 1.6666752422e-1f) * z * w + w;
    if (big) p = 1.57079637f - (p + p);
    return x < 0.0f ? -p : p;
}

// ---------------------------------------------------------------------------------
// RNG (shader.wgsl:605-631).
// ---------------------------------------------------------------------------------
RT_DEV uint32_t rng_next(uint32_t& s) {
    s = s * 747796405u + 2891336453u;
    uint32_t r = ((s >> ((s >> 28) + 4u)) ^ s) * 277803737u;
    return (r >> 22) ^ r;
}
RT_DEV float rng_uniform(uint32_t& s) {
    // f32(u32) / 4294967295.0: the divisor rounds to 2^32, so this is an exact scaling.
    return (float)rng_next(s) * 2.3283064365386963e-10f;
}
RT_DEV uint32_t rng_seed(uint32_t pixel_index, uint32_t sample_count) {
    uint32_t s = 0u;
    s ^= pixel_index;
    rng_next(s);
    s ^= sample_count;
    rng_next(s);
    return s;
}

// ---------------------------------------------------------------------------------
// Device scene view. Records are 128-bit aligned and read with vector loads.
//   nodes:   2 x float4 per node : (min.xyz, bits idx) (max.xyz, bits meta); meta = len (29 bits)
//            | bit 29 set when the leaf holds a sphere or plane | split axis << 30
//   records: 4 x float4 per primitive, indexed by leaf slot, followed by the brute-force
//            fallback lists (all spheres, then all planes) in reference order:
//     triangle (type 2): (a.xyz, type) (e0.xyz, mat) (e1.xyz, -) (-)
//     sphere   (type 0): (pos.xyz, type) (radius, radius^2, bits mat, -)
//     plane    (type 1): (pos.xyz, type) (n.xyz, bits mat) (M row x, -) (M row z, -)
//   tri_normals: 3 x float4 per leaf slot (the triangle's three vertex normals)
//   materials:   3 x float4 : (emission, alpha) (f0, p_spec) (kd/PI, p_diff)
// ---------------------------------------------------------------------------------
struct SceneView {
    const float4* __restrict__ nodes;
    const float4* __restrict__ records;
    const float4* __restrict__ tri_normals;
    const float4* __restrict__ materials;
    uint32_t slot_count;      // leaf slots; fallback spheres start here
    uint32_t sphere_count;
    uint32_t plane_count;
    uint32_t prune;           // 0 when a plane may stick out of its leaf box (rt_fast_host.hpp)
};

struct EnvView {
    cudaTextureObject_t texture;  // float4 texels, point sampling, unnormalised coords
    const float4* __restrict__ texels;  // the same texels, row-major, in linear memory
    const float4* __restrict__ alias;  // (probability, bits alias_index, pmf, -)
    uint32_t width, height;
};

struct CameraParams {
    float pos[3];
    float rot[9];  // column-major
    float fov_y;
};

struct Counts {
    uint32_t rays, shadow_rays, paths;
    uint32_t nodes, tris, spheres, planes, hits, tri_hits, shaded, env_lookups, fallback;
};

RT_DEV float4 ld4(const float4* p) {
#ifdef __CUDA_ARCH__
    return __ldg(p);
#else
    return *p;
#endif
}

// Two neighbouring float4 (32-byte aligned) with ONE 256-bit access (sm_100), for the
// wavefront's own state. Only 256-bit accesses can carry an L2 eviction priority.
// (Measured: the path records of shade / finish gain ~1 % this way; the walk's node and
// triangle fetches got 3 % slower with 256-bit loads and stay 128-bit.)
// POLICY 1: evict first from L2 (data that streams through once per wave).
template <int POLICY = 0>
RT_DEVONLY void ld8w(const float4* p, float4& a, float4& b) {
    if (POLICY == 1)
        asm volatile("ld.global.L2::evict_first.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                     : "=f"(a.x), "=f"(a.y), "=f"(a.z), "=f"(a.w), "=f"(b.x), "=f"(b.y), "=f"(b.z), "=f"(b.w)
                     : "l"(p)
                     : "memory");
    else
        asm volatile("ld.global.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                     : "=f"(a.x), "=f"(a.y), "=f"(a.z), "=f"(a.w), "=f"(b.x), "=f"(b.y), "=f"(b.z), "=f"(b.w)
                     : "l"(p)
                     : "memory");
}
template <int POLICY = 0>
RT_DEVONLY void st8(float4* p, float4 a, float4 b) {
    if (POLICY == 1)
        asm volatile("st.global.L2::evict_first.v8.f32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(p), "f"(a.x), "f"(a.y),
                     "f"(a.z), "f"(a.w), "f"(b.x), "f"(b.y), "f"(b.z), "f"(b.w)
                     : "memory");
    else
        asm volatile("st.global.v8.f32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(p), "f"(a.x), "f"(a.y), "f"(a.z), "f"(a.w),
                     "f"(b.x), "f"(b.y), "f"(b.z), "f"(b.w)
                     : "memory");
}

// ---------------------------------------------------------------------------------
// Primitive tests. Each returns the candidate distance or leaves `hit` false; the
// comparison forms are the reference's so NaNs fall the same way.
// ---------------------------------------------------------------------------------

// shader.wgsl:409-444 (normal interpolation deferred to hit reconstruction).
RT_DEV bool test_triangle(V3 o, V3 d, float4 r0, float4 r1, float4 r2, float& t, float& u,
                          float& v) {
    V3 a = xyz(r0), e0 = xyz(r1), e1 = xyz(r2);
    V3 oa = o - a;
    V3 p0 = cross(oa, e0);
    V3 p1 = cross(d, e1);
    float det = dot(e0, p1);
    float inv_det = 1.0f / det;
    if (fabsf(det) < 1.0e-8f) return false;
    u = dot(oa, p1) * inv_det;
    v = dot(d, p0) * inv_det;
    if (u < 0.0f || 1.0f < u) return false;
    if (v < 0.0f || 1.0f < (u + v)) return false;
    t = dot(e1, p0) * inv_det;
    if (t < 1.0e-5f) return false;
    return true;
}

// shader.wgsl:295-333.
RT_DEV bool test_sphere(V3 o, V3 d, float4 r0, float4 r1, float& t) {
    const float EPSILON = 1.0e-4f;
    V3 l = o - xyz(r0);
    float a = dot(d, d);
    float b = 2.0f * dot(d, l);
    float c = dot(l, l) - r1.y;  // r1.y = radius * radius
    float disc = b * b - 4.0f * a * c;
    if (disc < 0.0f) return false;
    if (disc == 0.0f) {
        t = -0.5f * b / a;
    } else {
        float sq = sqrtf(disc);
        float q = b > 0.0f ? -0.5f * (b + sq) : -0.5f * (b - sq);
        float t_0 = q / a;
        float t_1 = c / q;
        if (t_0 < EPSILON) t = t_1;
        else if (t_1 < EPSILON) t = t_0;
        else t = fminf(t_0, t_1);
    }
    if (t < EPSILON) return false;
    return true;
}

// shader.wgsl:362-391.
RT_DEV bool test_plane(V3 o, V3 d, float4 r0, float4 r1, float4 r2, float4 r3, float& t) {
    V3 n = xyz(r1), pos = xyz(r0);
    float den = dot(n, d);
    if (fabsf(den) < 0.0001f) return false;
    t = dot(n, pos - o) / den;
    if (t < 0.001f) return false;
    V3 il = (o + d * t) - pos;
    // x and z rows of base_change_matrix * inter_local (the y row is never read).
    float px = (r2.x * il.x + r2.y * il.y) + r2.z * il.z;
    float pz = (r3.x * il.x + r3.y * il.y) + r3.z * il.z;
    if (px < 0.0f || 1.0f < px || pz < 0.0f || 1.0f < pz) return false;
    return true;
}

struct Hit {
    float t;
    uint32_t rec;  // record index; 0xFFFFFFFF = none
    float u, v;
};

template <bool COUNT>
RT_DEV void test_record(const SceneView& sc, uint32_t rec, V3 o, V3 d, Hit& best, Counts& cn) {
    const float4* p = sc.records + 4u * (size_t)rec;
    float4 r0 = ld4(p), r1 = ld4(p + 1);
    int type = f_as_i(r0.w);
    float t = 0.0f, u = 0.0f, v = 0.0f;
    bool hit;
    if (type == 2) {
        float4 r2 = ld4(p + 2);
        if (COUNT) cn.tris++;
        hit = test_triangle(o, d, r0, r1, r2, t, u, v);
    } else if (type == 0) {
        if (COUNT) cn.spheres++;
        hit = test_sphere(o, d, r0, r1, t);
    } else {
        float4 r2 = ld4(p + 2), r3 = ld4(p + 3);
        if (COUNT) cn.planes++;
        hit = test_plane(o, d, r0, r1, r2, r3, t);
    }
    if (hit && t < best.t) {  // strict <: the first candidate met wins ties (shader.wgsl:500)
        best.t = t;
        best.rec = rec;
        best.u = u;
        best.v = v;
    }
}

// ---------------------------------------------------------------------------------
// Ray casting: cast_ray_bvh (shader.wgsl:469-564) in the reference's visiting order
// (near child chosen by the sign of inv_dir[split_axis]) plus cast_ray's brute-force
// fallback (shader.wgsl:583-598), as ONE routine so the intersection code exists once
// per call site.
//
// Structure ("while-while"): every lane first walks interior nodes until it owns a
// pending leaf, then all lanes test their leaf primitives together; the fallback list
// (all spheres, then all planes, stored behind the leaf slots) is handed to the same
// primitive loop as a last virtual leaf when the BVH found nothing.
//
// Result-preserving changes versus the literal shader:
//   * the slab test's far bound starts at 1.001 x the closest hit so far instead of
//     INFINITY (closest-hit pruning);
//   * any_hit (shadow rays) returns at the first accepted hit;
//   * when no component of 1/d is infinite a slab can never produce NaN, so the
//     swap/compare form reduces to min/max; rays with an infinite component take the
//     literal select form, which keeps NaN slabs inert exactly like the WGSL.
// `stack` is caller-provided storage with capacity >= the tree depth.
// ---------------------------------------------------------------------------------
#define RT_NONE 0xFFFFFFFFu

#ifndef RT_TRACE_QUAL
#define RT_TRACE_QUAL __device__ __noinline__
#endif
template <bool COUNT>
RT_TRACE_QUAL void trace(const SceneView& sc, V3 o, V3 d, bool any_hit, bool with_fallback,
                         Hit& best, uint32_t* stack, Counts& cn) {
    const V3 inv = v3(1.0f / d.x, 1.0f / d.y, 1.0f / d.z);
    // "finite" is false for +-inf and NaN.
    const bool plain = fabsf(inv.x) < 3.0e38f && fabsf(inv.y) < 3.0e38f && fabsf(inv.z) < 3.0e38f;
    best.t = RT_INFINITY;
    best.rec = RT_NONE;
    best.u = 0.0f;
    best.v = 0.0f;
    float limit = RT_INFINITY;
    uint32_t sp = 0u;
    uint32_t cur = 0u;
    uint32_t leaf_begin = 0u, leaf_end = 0u;
    bool fallback = with_fallback && (sc.sphere_count + sc.plane_count) > 0u;

    for (;;) {
        // ---- phase A: descend until this lane holds a leaf or has run out of nodes ----
        while (cur != RT_NONE && leaf_begin == leaf_end) {
            const float4 n0 = ld4(sc.nodes + 2u * (size_t)cur);
            const float4 n1 = ld4(sc.nodes + 2u * (size_t)cur + 1);
            if (COUNT) cn.nodes++;
            // ray_intersects_bounds (shader.wgsl:262-293)
            float tnx = (n0.x - o.x) * inv.x, tfx = (n1.x - o.x) * inv.x;
            float tny = (n0.y - o.y) * inv.y, tfy = (n1.y - o.y) * inv.y;
            float tnz = (n0.z - o.z) * inv.z, tfz = (n1.z - o.z) * inv.z;
            float t0, t1;
            if (plain) {
                t0 = fmaxf(fmaxf(0.0f, fminf(tnx, tfx)), fmaxf(fminf(tny, tfy), fminf(tnz, tfz)));
                t1 = fminf(fminf(limit, fmaxf(tnx, tfx)), fminf(fmaxf(tny, tfy), fmaxf(tnz, tfz)));
            } else {
                t0 = 0.0f;
                t1 = limit;
                bool sw = tnx > tfx;
                float nr = sw ? tfx : tnx, fr = sw ? tnx : tfx;
                t0 = nr > t0 ? nr : t0;
                t1 = fr < t1 ? fr : t1;
                sw = tny > tfy;
                nr = sw ? tfy : tny, fr = sw ? tny : tfy;
                t0 = nr > t0 ? nr : t0;
                t1 = fr < t1 ? fr : t1;
                sw = tnz > tfz;
                nr = sw ? tfz : tnz, fr = sw ? tnz : tfz;
                t0 = nr > t0 ? nr : t0;
                t1 = fr < t1 ? fr : t1;
            }
            uint32_t next = RT_NONE;  // where to go if this node yields nothing to descend into
            if (!(t0 > t1)) {
                const uint32_t idx = f_as_u(n0.w);
                const uint32_t meta = f_as_u(n1.w);
                const uint32_t len = meta & 0x1FFFFFFFu;
                if (len > 0u) {
                    leaf_begin = idx;
                    leaf_end = idx + len;
                } else {
                    const bool second_first = comp(inv, meta >> 30) < 0.0f;
                    stack[sp++] = second_first ? cur + 1u : idx;
                    next = second_first ? idx : cur + 1u;
                }
            }
            if (next == RT_NONE && sp > 0u) next = stack[--sp];
            cur = next;
        }
        if (leaf_begin == leaf_end) {
            // BVH exhausted. cast_ray falls back to brute force only when nothing was hit.
            if (!fallback || best.rec != RT_NONE) break;
            fallback = false;
            if (COUNT) cn.fallback++;
            leaf_begin = sc.slot_count;
            leaf_end = sc.slot_count + sc.sphere_count + sc.plane_count;
        }
        // ---- phase B: test the pending primitives ----
        while (leaf_begin < leaf_end) {
            test_record<COUNT>(sc, leaf_begin, o, d, best, cn);
            leaf_begin++;
            if (any_hit && best.rec != RT_NONE) return;
        }
        leaf_begin = leaf_end = 0u;
        if (sc.prune) limit = fminf(RT_INFINITY, best.t * 1.001f);
    }
}

// Deferred hit reconstruction: hit_point, shading normal, material row.
RT_DEV void reconstruct_hit(const SceneView& sc, V3 o, V3 d, const Hit& h, V3& point, V3& normal,
                            uint32_t& material) {
    const float4* p = sc.records + 4u * (size_t)h.rec;
    float4 r0 = ld4(p), r1 = ld4(p + 1);
    int type = f_as_i(r0.w);
    point = o + d * h.t;
    if (type == 2) {
        // shader.wgsl:446-457
        const float4* np = sc.tri_normals + 3u * (size_t)h.rec;
        V3 n0 = xyz(ld4(np)), n1 = xyz(ld4(np + 1)), n2 = xyz(ld4(np + 2));
        V3 n = normalize(((1.0f - h.u - h.v) * n0 + h.u * n1) + h.v * n2);
        if (dot(n, d) > 0.0f) n = n * -1.0f;
        normal = n;
        material = f_as_u(r1.w);
    } else if (type == 0) {
        // shader.wgsl:335-351; length_squared(pos - origin) - r^2 equals `c` bit for bit.
        V3 pos = xyz(r0);
        V3 n = normalize(point - pos);
        V3 l = pos - o;
        if (dot(l, l) - r1.y < 1.0e-6f) n = n * -1.0f;
        normal = n;
        material = f_as_u(r1.z);
    } else {
        // shader.wgsl:393-397 (sic: the flip tests the ray origin against the world origin)
        V3 n = xyz(r1);
        if (dot(o, n) < 0.0f) n = n * -1.0f;
        normal = n;
        material = f_as_u(r1.w);
    }
}

// ---------------------------------------------------------------------------------
// Environment (shader.wgsl:667-831).
// ---------------------------------------------------------------------------------
RT_DEV void direction_to_uv(V3 d, float& u, float& v) {
    u = rt_atan2(d.z, d.x) * RT_INV_PI * 0.5f + 0.5f;
    v = 0.5f - rt_asin(d.y) * RT_INV_PI;
}

RT_DEV float pixel_solid_angle(float v, const EnvView& env) {
    float theta = RT_PI * v;
    float sin_t = fmaxf(1.0e-6f, rt_sin(theta));
    float d_phi = RT_TWO_PI / (float)env.width;
    float d_theta = RT_PI / (float)env.height;
    return d_phi * d_theta * sin_t;
}

// textureSampleLevel(.., uv, 0) under the reference sampler (clamp-to-edge, linear):
// four point fetches through the texture unit and an f32 bilinear blend.
RT_DEVONLY V3 env_bilinear(const EnvView& env, float u, float v) {
    float fx = u * (float)env.width - 0.5f;
    float fy = v * (float)env.height - 0.5f;
    float x0f = floorf(fx), y0f = floorf(fy);
    float wx = fx - x0f, wy = fy - y0f;
    float xmax = (float)(env.width - 1u), ymax = (float)(env.height - 1u);
    // clamp to [0, size-1]; the !(f > 0) form sends NaN to 0 (arithmetic contract).
    float x0 = !(x0f > 0.0f) ? 0.0f : (x0f >= xmax ? xmax : x0f);
    float x1f = x0f + 1.0f;
    float x1 = !(x1f > 0.0f) ? 0.0f : (x1f >= xmax ? xmax : x1f);
    float y0 = !(y0f > 0.0f) ? 0.0f : (y0f >= ymax ? ymax : y0f);
    float y1f = y0f + 1.0f;
    float y1 = !(y1f > 0.0f) ? 0.0f : (y1f >= ymax ? ymax : y1f);
#if RT_ENV_LINEAR
    // The coordinates are already clamped, so a point fetch is a plain indexed load.
    const uint32_t ix0 = (uint32_t)x0, ix1 = (uint32_t)x1;
    const size_t row0 = (size_t)(uint32_t)y0 * env.width, row1 = (size_t)(uint32_t)y1 * env.width;
    float4 t00 = ld4(env.texels + row0 + ix0);
    float4 t10 = ld4(env.texels + row0 + ix1);
    float4 t01 = ld4(env.texels + row1 + ix0);
    float4 t11 = ld4(env.texels + row1 + ix1);
#else
    float4 t00 = tex2D<float4>(env.texture, x0 + 0.5f, y0 + 0.5f);
    float4 t10 = tex2D<float4>(env.texture, x1 + 0.5f, y0 + 0.5f);
    float4 t01 = tex2D<float4>(env.texture, x0 + 0.5f, y1 + 0.5f);
    float4 t11 = tex2D<float4>(env.texture, x1 + 0.5f, y1 + 0.5f);
#endif
    V3 top = xyz(t00) * (1.0f - wx) + xyz(t10) * wx;
    V3 bottom = xyz(t01) * (1.0f - wx) + xyz(t11) * wx;
    return top * (1.0f - wy) + bottom * wy;
}

// environment_direction_pdf (shader.wgsl:753-769) given the uv of the direction.
RT_DEV float env_direction_pdf(const EnvView& env, float u, float v) {
    uint32_t x = min((uint32_t)(u * (float)env.width), env.width - 1u);
    uint32_t y = min((uint32_t)(v * (float)env.height), env.height - 1u);
    float pmf = ld4(env.alias + (size_t)(x + y * env.width)).z;
    return pmf / pixel_solid_angle(v, env);
}

// ---------------------------------------------------------------------------------
// BSDF (shader.wgsl:834-1210). Material rows are pre-baked by bake_materials().
// ---------------------------------------------------------------------------------
struct Material {
    V3 emission, f0, fd;  // fd = surface_kd * (1 / PI)
    float alpha, p_spec, p_diff;
};

RT_DEV Material load_material(const SceneView& sc, uint32_t id) {
    const float4* p = sc.materials + 3u * (size_t)id;
    float4 m0 = ld4(p), m1 = ld4(p + 1), m2 = ld4(p + 2);
    Material m;
    m.emission = xyz(m0);
    m.alpha = m0.w;
    m.f0 = xyz(m1);
    m.p_spec = m1.w;
    m.fd = xyz(m2);
    m.p_diff = m2.w;
    return m;
}

RT_DEV float luminance(V3 c) { return (0.2126f * c.x + 0.7152f * c.y) + 0.0722f * c.z; }

// make_bsdf_material + surface_f0 + surface_kd + lobe probabilities (shader.wgsl:850-881,
// 1109-1110) from one reference-layout Material (48 B = 12 floats).
RT_DEV void bake_material(const float* src, float4* dst) {
    V3 color = v3(src[0], src[1], src[2]);
    float roughness = src[3], metallic = src[4];
    V3 emission = v3(src[8], src[9], src[10]);
    float alpha = fmaxf(0.001f, roughness * roughness);
    float tm = saturate(metallic);
    V3 f0 = (1.0f - tm) * v3s(0.04f) + tm * color;  // lerp_vec3f(DIELECTRIC_F0, color, t)
    V3 kd0 = color * (1.0f - saturate(metallic));
    V3 kd = kd0 * (1.0f - fmaxf(f0.x, fmaxf(f0.y, f0.z)));
    V3 fd = kd * RT_INV_PI;
    float ps = saturate(luminance(f0));
    float pd = 1.0f - ps;
    dst[0] = make_float4(emission.x, emission.y, emission.z, alpha);
    dst[1] = make_float4(f0.x, f0.y, f0.z, ps);
    dst[2] = make_float4(fd.x, fd.y, fd.z, pd);
}

struct Frame {
    V3 tangent, bitangent, normal;
};
RT_DEV Frame make_frame(V3 n) {
    V3 helper = fabsf(n.z) < 0.999f ? v3(0.0f, 0.0f, 1.0f) : v3(1.0f, 0.0f, 0.0f);
    Frame f;
    f.tangent = normalize(cross(helper, n));
    f.bitangent = cross(n, f.tangent);
    f.normal = n;
    return f;
}
RT_DEV V3 to_local(const Frame& f, V3 w) {
    return v3(dot(w, f.tangent), dot(w, f.bitangent), dot(w, f.normal));
}
RT_DEV V3 to_world(const Frame& f, V3 l) {
    return normalize((f.tangent * l.x + f.bitangent * l.y) + f.normal * l.z);
}

RT_DEV float d_ggx(float ndh, float alpha) {
    float a2 = alpha * alpha;
    float den = (ndh * ndh) * (a2 - 1.0f) + 1.0f;
    return a2 / (RT_PI * den * den);
}
RT_DEV float lambda_ggx(float ndv, float alpha) {
    float n2 = ndv * ndv;
    return (sqrtf(1.0f + alpha * alpha * (1.0f - n2) / n2) - 1.0f) / 2.0f;
}
RT_DEV float g1_ggx(float ndv, float alpha) { return 1.0f / (1.0f + lambda_ggx(ndv, alpha)); }

// bsdf_eval_local and bsdf_pdf_local (shader.wgsl:1053-1114) evaluated together: they
// share h = normalize(wo + wi) and G1(wo).
RT_DEV void bsdf_eval_pdf(V3 wo, V3 wi, const Material& m, V3& f, float& pdf) {
    if (wo.z <= 0.0f || wi.z <= 0.0f) {
        f = v3s(0.0f);
        pdf = 0.0f;
        return;
    }
    float ndo = wo.z, ndi = wi.z;
    V3 h = normalize(wo + wi);
    float ndh = saturate(h.z);
    float D = d_ggx(ndh, m.alpha);
    float g1o = g1_ggx(ndo, m.alpha);
    float G = g1o * g1_ggx(ndi, m.alpha);
    float hdo = dot(h, wo);
    // f_schlick (shader.wgsl:1045-1051)
    float x = 1.0f - saturate(hdo);
    float x2 = x * x;
    float x5 = x2 * x2 * x;
    V3 F = m.f0 + (v3s(1.0f) - m.f0) * x5;
    V3 fs = ((D * G) / (4.0f * ndo * ndi)) * F;
    f = m.fd + fs;

    // pdf_cosine_hemisphere, pdf_specular_wi_visible, pdf_ggx_half_vector_visible
    float pdf_diff = ndi / RT_PI;
    float pdf_spec = 0.0f;
    float wo_dot_h = fabsf(dot(wo, h));
    if (!(wo_dot_h <= 0.0f)) {
        float pdf_h = 0.0f;
        if (!(h.z <= 0.0f))
            pdf_h = d_ggx(h.z, m.alpha) * g1o * fmaxf(0.0f, dot(wo, h)) / ndo;
        pdf_spec = pdf_h / (4.0f * wo_dot_h);
    }
    pdf = m.p_diff * pdf_diff + m.p_spec * pdf_spec;
}

RT_DEV float power_heuristic(float a, float b) {
    float a2 = a * a, b2 = b * b;
    return a2 / (a2 + b2);
}

// sample_ggx_visible_half_vector (shader.wgsl:962-1009).
RT_DEV V3 sample_ggx_vndf(float sx, float sy, V3 wo, float alpha) {
    V3 vs = normalize(wo * v3(alpha, alpha, 1.0f));
    float len2 = vs.x * vs.x + vs.y * vs.y;
    V3 tx = len2 > 0.0f ? v3(-vs.y, vs.x, 0.0f) * (1.0f / sqrtf(len2)) : v3(1.0f, 0.0f, 0.0f);
    V3 ty = cross(vs, tx);
    float radius = sqrtf(sx);
    float s, c;
    rt_sincos(RT_TWO_PI * sy, &s, &c);
    float dx = radius * c, dy = radius * s;
    float a = sqrtf(fmaxf(0.0f, 1.0f - dx * dx));
    dy = (1.0f - vs.z) * a + vs.z * dy;  // lerp_f32(a, dy, vs.z)
    V3 hs = (dx * tx + dy * ty) + sqrtf(fmaxf(0.0f, 1.0f - dx * dx - dy * dy)) * vs;
    return normalize(v3(alpha * hs.x, alpha * hs.y, fmaxf(0.0f, hs.z)));
}

// Camera ray + RNG seeding for one pixel-sample (shader.wgsl:1307-1362).
RT_DEV void camera_ray(const CameraParams& cam, float sin_half_fov, uint32_t px, uint32_t py,
                       uint32_t width, uint32_t height, uint32_t sample, uint32_t& rng, V3& o,
                       V3& d) {
    rng = rng_seed(py * width + px, sample);
    float angle = rng_uniform(rng) * 2.0f * 3.1415926f;
    float s, c;
    rt_sincos(angle, &s, &c);
    float r = sqrtf(rng_uniform(rng));
    float jx = (float)px + c * r, jy = (float)py + s * r;
    float sx = ((jx / (float)width) * 2.0f - 1.0f) * 1.0f;
    float sy = ((jy / (float)height) * 2.0f - 1.0f) * -1.0f;
    float aspect = (float)width / (float)height;
    float cx = sx * sin_half_fov * aspect, cy = sy * sin_half_fov, cz = -1.0f;
    V3 c0 = v3(cam.rot[0], cam.rot[1], cam.rot[2]);
    V3 c1 = v3(cam.rot[3], cam.rot[4], cam.rot[5]);
    V3 c2 = v3(cam.rot[6], cam.rot[7], cam.rot[8]);
    V3 w = (c0 * cx + c1 * cy) + c2 * cz;
    o = v3(cam.pos[0], cam.pos[1], cam.pos[2]);
    d = normalize(w);
}

}  // namespace rtd
