// shader_oracle.cpp — CPU ORACLE. TEST INFRASTRUCTURE ONLY.
//
// A literal, scalar C++ restatement of the reference's per-pixel hot path,
// /root/reference/src/shaders/shader.wgsl (main 1305-1373 and everything it calls).
// It consumes the reference's own buffer layouts (include/rt_types.h mirrors the
// encase/std430 bytes) and walks them exactly like the WGSL does: AoS nodes, the
// triangle -> vertices/normals double indirection, normals computed on every
// candidate hit, the un-pruned box test, full closest-hit shadow rays.
//
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
// leg may load this library. The product (librt_backend.so) never links or calls it.
//
// PARITY UNPINNED: the reference ships no tests, golden vectors or images, and cannot
// be built or run in this image (no rustc/cargo, no Vulkan ICD, missing HDRI blobs),
// so this oracle is pinned only by (a) the reference's input fixtures and (b) the
// BVH dumps in SURVEY.md Appendix B (an independent restatement), not by reference
// outputs.
//
// Arithmetic contract (what "bit-exact" between this oracle and the CUDA kernels
// means). WGSL leaves rounding/contraction of most builtins implementation-defined;
// the contract fixes one realisation:
//   * every +,-,*,/ and sqrt is a single correctly-rounded IEEE-754 binary32
//     operation, evaluated in the WGSL source's left-to-right order; no FMA
//     contraction (build with -ffp-contract=off; the kernels use --fmad=false);
//   * dot(a,b) = (a.x*b.x + a.y*b.y) + a.z*b.z; cross() as written below;
//     length(v) = sqrt(dot(v,v)); normalize(v) = v * (1 / length(v));
//     inverseSqrt(x) = 1 / sqrt(x); reflect(i,n) = i - (2*dot(n,i))*n;
//   * sin, cos, atan2, asin are the polynomial routines below (Cody-Waite reduction +
//     Cephes single-precision minimax polynomials), spelled out op by op;
//   * abstract-float constant expressions are folded in f64 then rounded to f32 once
//     (PI, INV_PI, 2*PI, 1/PI — shader.wgsl:239-240, 746, 894, 1084);
//   * textureSampleLevel(.., uv, 0) with the reference sampler (clamp-to-edge, linear
//     magnification; src/state.rs:134-142) = bilinear filter in f32 over 4 texels.
#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <thread>
#include <vector>

#include "../include/rt_types.h"

namespace {

struct vec3 {
    float x, y, z;
};
struct vec2 {
    float x, y;
};
inline vec3 V(float x, float y, float z) { return vec3{x, y, z}; }
inline vec3 V(float s) { return vec3{s, s, s}; }
inline vec3 operator+(vec3 a, vec3 b) { return V(a.x + b.x, a.y + b.y, a.z + b.z); }
inline vec3 operator-(vec3 a, vec3 b) { return V(a.x - b.x, a.y - b.y, a.z - b.z); }
inline vec3 operator-(vec3 a) { return V(-a.x, -a.y, -a.z); }
inline vec3 operator*(vec3 a, vec3 b) { return V(a.x * b.x, a.y * b.y, a.z * b.z); }
inline vec3 operator*(vec3 a, float s) { return V(a.x * s, a.y * s, a.z * s); }
inline vec3 operator*(float s, vec3 a) { return V(s * a.x, s * a.y, s * a.z); }
inline vec3 operator/(vec3 a, float s) { return V(a.x / s, a.y / s, a.z / s); }
inline float dot(vec3 a, vec3 b) { return (a.x * b.x + a.y * b.y) + a.z * b.z; }
inline float dot2(vec2 a, vec2 b) { return a.x * b.x + a.y * b.y; }
inline vec3 cross(vec3 a, vec3 b) {
    return V(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x);
}
inline float length(vec3 v) { return std::sqrt(dot(v, v)); }
inline vec3 normalize(vec3 v) { return v * (1.0f / length(v)); }
inline float inverse_sqrt(float x) { return 1.0f / std::sqrt(x); }
inline vec3 reflect(vec3 i, vec3 n) { return i - (2.0f * dot(n, i)) * n; }
inline float saturate(float x) { return std::fmin(std::fmax(x, 0.0f), 1.0f); }
inline float comp(vec3 v, uint32_t axis) { return axis == 0 ? v.x : (axis == 1 ? v.y : v.z); }

// ---- transcendental routines of the arithmetic contract --------------------------

// pi/2 split for Cody-Waite reduction: kPio2A + kPio2B + kPio2C ~= pi/2, kPio2A has
// few mantissa bits so q * kPio2A is exact for the small quadrant counts seen here.
const float kTwoOverPi = 0.636619772f;
const float kPio2A = 1.5703125f;
const float kPio2B = 4.837512969970703125e-4f;
const float kPio2C = 7.54978995489188216e-8f;

// Reduces x to r in about [-pi/4, pi/4] and returns the quadrant (x ~= q*pi/2 + r).
inline float reduce_pio2(float x, int* quadrant) {
    float q = std::rint(x * kTwoOverPi);
    float r = ((x - q * kPio2A) - q * kPio2B) - q * kPio2C;
    *quadrant = (int)q;
    return r;
}
inline float sin_poly(float r) {
    float z = r * r;
    return r + (r * z) * ((-1.9515295891e-4f * z + 8.3321608736e-3f) * z - 1.6666654611e-1f);
}
inline float cos_poly(float r) {
    float z = r * r;
    return (1.0f - 0.5f * z) +
           (z * z) * ((2.443315711809948e-5f * z - 1.388731625493765e-3f) * z + 4.166664568298827e-2f);
}
float rt_sin(float x) {
    int q;
    float r = reduce_pio2(x, &q);
    switch (q & 3) {
        case 0: return sin_poly(r);
        case 1: return cos_poly(r);
        case 2: return -sin_poly(r);
        default: return -cos_poly(r);
    }
}
float rt_cos(float x) {
    int q;
    float r = reduce_pio2(x, &q);
    switch (q & 3) {
        case 0: return cos_poly(r);
        case 1: return -sin_poly(r);
        case 2: return -cos_poly(r);
        default: return sin_poly(r);
    }
}

const float kPiF = 3.14159274f;      // true pi rounded to f32 (range constants only)
const float kPio2F = 1.57079637f;
const float kPio4F = 0.785398185f;

// atan for t >= 0 (Cephes atanf).
inline float atan_pos(float t) {
    float y0;
    if (t > 2.41421366f) {  // tan(3 pi / 8)
        y0 = kPio2F;
        t = -(1.0f / t);
    } else if (t > 0.414213568f) {  // tan(pi / 8)
        y0 = kPio4F;
        t = (t - 1.0f) / (t + 1.0f);
    } else {
        y0 = 0.0f;
    }
    float z = t * t;
    float p = (((8.05374449538e-2f * z - 1.38776856032e-1f) * z + 1.99777106478e-1f) * z -
               3.33329491539e-1f) * z * t + t;
    return y0 + p;
}
float rt_atan2(float y, float x) {
    if (x == 0.0f) {
        if (y > 0.0f) return kPio2F;
        if (y < 0.0f) return -kPio2F;
        return 0.0f;
    }
    float a = atan_pos(std::fabs(y / x));
    if (x < 0.0f) a = kPiF - a;
    return y < 0.0f ? -a : a;
}
float rt_asin(float x) {
    float a = std::fabs(x);
    if (a > 1.0f) a = 1.0f;  // normalised directions can exceed 1 by an ulp
    float z, w;
    bool big = a > 0.5f;
    if (big) {
        z = 0.5f * (1.0f - a);
        w = std::sqrt(z);
    } else {
        w = a;
        z = a * a;
    }
    float p = ((((4.2163199048e-2f * z + 2.4181311049e-2f) * z + 4.5470025998e-2f) * z +
                7.4953002686e-2f) * z + 1.6666752422e-1f) * z * w + w;
    if (big) p = kPio2F - (p + p);
    return x < 0.0f ? -p : p;
}

// ---- shader.wgsl constants --------------------------------------------------------
const float INFINITY_F = 1.70141183460469231732e+38f;  // shader.wgsl:235
const float PI = (float)3.14159;                        // 239
const float INV_PI = (float)(1.0 / 3.14159);            // 240 (abstract-float fold)
const float TWO_PI = (float)(2.0 * 3.14159);            // "2 * PI", "2. * PI" folds

struct Ray {
    vec3 origin, direction;
};
struct HitInfo {
    bool did_hit;
    float distance;
    vec3 hit_point, normal;
    uint32_t material_id;
};
const HitInfo NO_HIT = {false, 0.0f, {0, 0, 0}, {0, 0, 0}, 0u};

struct Counters {
    uint64_t paths, rays, shadow_rays, nodes_tested, tris_tested, spheres_tested, planes_tested,
        hits, tri_hits, shaded_bounces, env_lookups, fallback_rays;
};

struct Job {
    RtSceneDesc scene;
    uint32_t env_count;
    const RtEnvironmentDesc* envs;
    RtCameraUniform camera;
    uint32_t width, height;
    uint32_t max_bounces;
    uint32_t env_index;
    uint32_t prune;  // 0: literal reference traversal; 1: closest-hit pruning + any-hit shadows
};

struct Ctx {
    const Job* job;
    Counters c;
};

inline vec3 ld3(const float* p) { return V(p[0], p[1], p[2]); }
inline vec3 ldv(const RtVec3& v) { return V(v.x, v.y, v.z); }

// shader.wgsl:55-84
struct Frame {
    vec3 tangent, bitangent, normal;
};
Frame make_frame(vec3 normal) {
    vec3 helper = std::fabs(normal.z) < 0.999f ? V(0, 0, 1) : V(1, 0, 0);  // select(f, t, cond)
    vec3 tangent = normalize(cross(helper, normal));
    vec3 bitangent = cross(normal, tangent);
    return Frame{tangent, bitangent, normal};
}
vec3 to_frame_local(const Frame& f, vec3 w) {
    return V(dot(w, f.tangent), dot(w, f.bitangent), dot(w, f.normal));
}
vec3 to_frame_world(const Frame& f, vec3 l) {
    return normalize((f.tangent * l.x + f.bitangent * l.y) + f.normal * l.z);
}

inline vec3 lerp_vec3f(vec3 a, vec3 b, float t) { return (1.0f - t) * a + t * b; }  // 242-244
inline float lerp_f32(float a, float b, float t) { return (1.0f - t) * a + t * b; }  // 246-248
inline float length_squared(vec3 v) { return dot(v, v); }
inline float max_component(vec3 v) { return std::fmax(v.x, std::fmax(v.y, v.z)); }

// shader.wgsl:262-293. t_1_init is INFINITY in the reference; the pruned variant
// (job.prune) passes the current closest distance times 1.001 instead.
bool ray_intersects_bounds(const Ray& ray, const RtBvhNode& node, vec3 inv_direction,
                           float t_1_init) {
    float t_0 = 0.0f;
    float t_1 = t_1_init;
    for (uint32_t axis = 0; axis < 3; axis++) {
        float t_near = (node.min[axis] - comp(ray.origin, axis)) * comp(inv_direction, axis);
        float t_far = (node.max[axis] - comp(ray.origin, axis)) * comp(inv_direction, axis);
        if (t_near > t_far) {
            float t_near_old = t_near;
            t_near = t_far;
            t_far = t_near_old;
        }
        if (t_near > t_0) t_0 = t_near;
        if (t_far < t_1) t_1 = t_far;
        if (t_0 > t_1) return false;
    }
    return true;
}

// shader.wgsl:295-360
HitInfo cast_ray_sphere(const Ray& ray, const RtSphere& sphere) {
    const float EPSILON = 1.0e-4f;
    vec3 pos = ld3(sphere.pos);
    vec3 l = ray.origin - pos;
    float a = dot(ray.direction, ray.direction);
    float b = 2.0f * dot(ray.direction, l);
    float c = dot(l, l) - sphere.radius * sphere.radius;

    float t;
    float discriminant = b * b - 4.0f * a * c;
    if (discriminant < 0.0f) {
        return NO_HIT;
    } else if (discriminant == 0.0f) {
        t = -0.5f * b / a;
    } else {
        float q = b > 0.0f ? -0.5f * (b + std::sqrt(discriminant))
                           : -0.5f * (b - std::sqrt(discriminant));
        float t_0 = q / a;
        float t_1 = c / q;
        if (t_0 < EPSILON) {
            t = t_1;
        } else if (t_1 < EPSILON) {
            t = t_0;
        } else {
            t = std::fmin(t_0, t_1);
        }
    }
    // `t < EPSILON` is false for NaN, as in WGSL.
    if (t < EPSILON) return NO_HIT;

    vec3 hit_point = ray.origin + ray.direction * t;
    vec3 normal = normalize(hit_point - pos);
    if (length_squared(pos - ray.origin) - sphere.radius * sphere.radius < 1.0e-6f)
        normal = normal * -1.0f;
    return HitInfo{true, t, hit_point, normal, sphere.material_id};
}

// shader.wgsl:362-406
HitInfo cast_ray_plane(const Ray& ray, const RtPlane& plane) {
    vec3 n = ld3(plane.normal), pos = ld3(plane.pos);
    float denominator = dot(n, ray.direction);
    if (std::fabs(denominator) < 0.0001f) return NO_HIT;
    float t = dot(n, pos - ray.origin) / denominator;
    if (t < 0.001f) return NO_HIT;

    vec3 inter = ray.origin + ray.direction * t;
    vec3 inter_local = inter - pos;
    // mat3x3 * vec3 = col0 * v.x + col1 * v.y + col2 * v.z
    vec3 c0 = ld3(plane.base_change[0]), c1 = ld3(plane.base_change[1]),
         c2 = ld3(plane.base_change[2]);
    vec3 ips = (c0 * inter_local.x + c1 * inter_local.y) + c2 * inter_local.z;
    if (ips.x < 0.0f || 1.0f < ips.x || ips.z < 0.0f || 1.0f < ips.z) return NO_HIT;

    vec3 normal = n;
    if (dot(ray.origin, normal) < 0.0f) normal = normal * -1.0f;
    return HitInfo{true, t, inter, normal, plane.material_id};
}

// shader.wgsl:409-466
HitInfo cast_ray_triangle(const Job& job, const Ray& ray, const RtTriangle& tri) {
    vec3 a = ldv(job.scene.vertices[tri.vertex_0]);
    vec3 b = ldv(job.scene.vertices[tri.vertex_1]);
    vec3 c = ldv(job.scene.vertices[tri.vertex_2]);
    vec3 pos = a;
    vec3 edge_0 = b - a;
    vec3 edge_1 = c - a;
    vec3 perp_to_edge_0 = cross(ray.origin - pos, edge_0);
    vec3 perp_to_edge_1 = cross(ray.direction, edge_1);
    float determinant = dot(edge_0, perp_to_edge_1);
    float inverse_determinant = 1.0f / determinant;
    if (std::fabs(determinant) < 1.0e-8f) return NO_HIT;

    float u = dot(ray.origin - pos, perp_to_edge_1) * inverse_determinant;
    float v = dot(ray.direction, perp_to_edge_0) * inverse_determinant;
    if (u < 0.0f || 1.0f < u) return NO_HIT;
    if (v < 0.0f || 1.0f < (u + v)) return NO_HIT;

    float t = dot(edge_1, perp_to_edge_0) * inverse_determinant;
    if (t < 1.0e-5f) return NO_HIT;

    vec3 normal_0 = ldv(job.scene.normals[tri.normal_0]);
    vec3 normal_1 = ldv(job.scene.normals[tri.normal_1]);
    vec3 normal_2 = ldv(job.scene.normals[tri.normal_2]);
    vec3 normal = normalize(((1.0f - u - v) * normal_0 + u * normal_1) + v * normal_2);
    if (dot(normal, ray.direction) > 0.0f) normal = normal * -1.0f;
    return HitInfo{true, t, ray.origin + ray.direction * t, normal, tri.material_id};
}

// shader.wgsl:469-564. *slot_out receives the index into `primitives` of the returned
// hit (RT_MISS_SLOT when none). any_hit is only set for shadow rays of the pruned
// variant (the reference runs the full traversal and looks at did_hit alone).
HitInfo cast_ray_bvh(Ctx& ctx, const Ray& ray, uint32_t* slot_out, bool any_hit) {
    const Job& job = *ctx.job;
    vec3 inv = V(1.0f / ray.direction.x, 1.0f / ray.direction.y, 1.0f / ray.direction.z);
    HitInfo result = {false, INFINITY_F, {0, 0, 0}, {0, 0, 0}, 0u};
    uint32_t result_slot = RT_MISS_SLOT;

    uint32_t nodes_to_visit[64];
    uint32_t stack_length = 0;
    uint32_t current = 0;
    for (;;) {
        const RtBvhNode& node = job.scene.bvh_nodes[current];
        float t_1_init = INFINITY_F;
        if (job.prune) t_1_init = std::fmin(INFINITY_F, result.distance * 1.001f);
        ctx.c.nodes_tested++;
        if (ray_intersects_bounds(ray, node, inv, t_1_init)) {
            if (node.primitives_len > 0) {
                for (uint32_t i = 0; i < node.primitives_len; i++) {
                    uint32_t slot = node.primitives_or_second_child_index + i;
                    const RtPrimitiveInfo& info = job.scene.primitives[slot];
                    HitInfo h = NO_HIT;
                    switch (info.primitive_type) {
                        case 0:
                            ctx.c.spheres_tested++;
                            h = cast_ray_sphere(ray, job.scene.spheres[info.index]);
                            break;
                        case 1:
                            ctx.c.planes_tested++;
                            h = cast_ray_plane(ray, job.scene.planes[info.index]);
                            break;
                        case 2:
                            ctx.c.tris_tested++;
                            h = cast_ray_triangle(job, ray, job.scene.triangles[info.index]);
                            break;
                        default: break;
                    }
                    if (h.did_hit && h.distance < result.distance) {
                        result = h;
                        result_slot = slot;
                        if (any_hit) {
                            *slot_out = result_slot;
                            return result;
                        }
                    }
                }
                if (stack_length == 0) break;
                stack_length--;
                current = nodes_to_visit[stack_length];
            } else {
                if (comp(inv, node.split_axis) < 0.0f) {
                    nodes_to_visit[stack_length] = current + 1;
                    stack_length++;
                    current = node.primitives_or_second_child_index;
                } else {
                    nodes_to_visit[stack_length] = node.primitives_or_second_child_index;
                    stack_length++;
                    current = current + 1;
                }
            }
        } else {
            if (stack_length == 0) break;
            stack_length--;
            current = nodes_to_visit[stack_length];
        }
    }
    *slot_out = result_slot;
    if (!result.did_hit) return NO_HIT;
    return result;
}

// Encoding of hits found only by the brute-force fallback (shader.wgsl:583-598).
inline uint32_t fallback_slot(uint32_t type, uint32_t index) {
    return 0x80000000u | (type << 28) | index;
}

// shader.wgsl:567-601
HitInfo cast_ray(Ctx& ctx, const Ray& ray, uint32_t* slot_out) {
    const Job& job = *ctx.job;
    HitInfo result = {false, INFINITY_F, {0, 0, 0}, {0, 0, 0}, 0u};
    {
        HitInfo r = cast_ray_bvh(ctx, ray, slot_out, false);
        if (r.did_hit) return r;
    }
    ctx.c.fallback_rays++;
    *slot_out = RT_MISS_SLOT;
    for (uint32_t i = 0; i < job.scene.sphere_count; i++) {
        ctx.c.spheres_tested++;
        HitInfo h = cast_ray_sphere(ray, job.scene.spheres[i]);
        if (h.did_hit && h.distance < result.distance) {
            result = h;
            *slot_out = fallback_slot(0, i);
        }
    }
    for (uint32_t i = 0; i < job.scene.plane_count; i++) {
        ctx.c.planes_tested++;
        HitInfo h = cast_ray_plane(ray, job.scene.planes[i]);
        if (h.did_hit && h.distance < result.distance) {
            result = h;
            *slot_out = fallback_slot(1, i);
        }
    }
    return result;
}

// ---- RNG: shader.wgsl:605-631 ------------------------------------------------------
inline uint32_t random_u32_uniform(uint32_t* s) {
    *s = *s * 747796405u + 2891336453u;
    uint32_t result = ((*s >> ((*s >> 28) + 4)) ^ *s) * 277803737u;
    result = (result >> 22) ^ result;
    return result;
}
inline void salt_rng(uint32_t* s, uint32_t salt) {
    *s = *s ^ salt;
    random_u32_uniform(s);
}
inline float random_uniform(uint32_t* s) {
    // f32(u32) rounds to nearest; 4294967295.0 rounds to 2^32 as an f32 literal.
    return (float)random_u32_uniform(s) / 4294967296.0f;
}
inline vec2 random_in_circle_uniform(uint32_t* s) {
    float angle = random_uniform(s) * 2.0f * 3.1415926f;
    vec2 p = {rt_cos(angle), rt_sin(angle)};
    float r = std::sqrt(random_uniform(s));
    return vec2{p.x * r, p.y * r};
}

// ---- environment: shader.wgsl:667-831 ----------------------------------------------
inline const RtAliasEntry& environment_alias_entry(const Job& job, uint32_t index, uint32_t env) {
    return job.envs[env].alias[index];
}

uint32_t random_index_in_environment(const Job& job, uint32_t env, uint32_t* rng) {
    const RtEnvironmentDesc& m = job.envs[env];
    uint32_t length = m.width * m.height;
    uint32_t index = std::min((uint32_t)(random_uniform(rng) * (float)length), length - 1);
    const RtAliasEntry& entry = environment_alias_entry(job, index, env);
    return random_uniform(rng) < entry.probability ? index : entry.alias_index;
}

vec2 direction_to_equirectangular_uv(vec3 d) {
    float u = rt_atan2(d.z, d.x) * INV_PI * 0.5f + 0.5f;
    float v = 0.5f - rt_asin(d.y) * INV_PI;
    return vec2{u, v};
}

vec3 equirectangular_uv_to_direction(vec2 uv) {
    float phi = (2.0f * uv.x - 1.0f) * PI;
    float theta = PI * uv.y;
    float sin_theta = rt_sin(theta);
    float cos_theta = rt_cos(theta);
    return V(sin_theta * rt_cos(phi), cos_theta, sin_theta * rt_sin(phi));
}

float environment_pixel_solid_angle(float v, const RtEnvironmentDesc& env) {
    float theta = PI * v;
    float sin_t = std::fmax(1.0e-6f, rt_sin(theta));
    float d_phi = TWO_PI / (float)env.width;
    float d_theta = PI / (float)env.height;
    return d_phi * d_theta * sin_t;
}

// u32(f32) of WGSL for the values met here: truncation, negatives/NaN -> 0.
inline uint32_t f2u(float f) { return f > 0.0f ? (f >= 4294967296.0f ? 0xFFFFFFFFu : (uint32_t)f) : 0u; }

float environment_direction_pdf(const Job& job, vec3 direction, uint32_t env) {
    const RtEnvironmentDesc& m = job.envs[env];
    vec2 uv = direction_to_equirectangular_uv(direction);
    uint32_t x = std::min(f2u(uv.x * (float)m.width), m.width - 1);
    uint32_t y = std::min(f2u(uv.y * (float)m.height), m.height - 1);
    uint32_t index = x + y * m.width;
    float pmf = environment_alias_entry(job, index, env).pmf;
    float delta_solid_angle = environment_pixel_solid_angle(uv.y, m);
    return pmf / delta_solid_angle;
}

// textureSampleLevel(tex, sampler, uv, 0).xyz with clamp-to-edge + linear filtering.
vec3 texture_sample_bilinear(Ctx& ctx, const RtEnvironmentDesc& env, vec2 uv) {
    ctx.c.env_lookups++;
    float fx = uv.x * (float)env.width - 0.5f;
    float fy = uv.y * (float)env.height - 0.5f;
    float x0f = std::floor(fx), y0f = std::floor(fy);
    float wx = fx - x0f, wy = fy - y0f;
    auto clampi = [](float f, uint32_t n) -> uint32_t {
        if (!(f > 0.0f)) return 0u;  // also NaN
        if (f >= (float)(n - 1)) return n - 1;
        return (uint32_t)f;
    };
    uint32_t x0 = clampi(x0f, env.width), x1 = clampi(x0f + 1.0f, env.width);
    uint32_t y0 = clampi(y0f, env.height), y1 = clampi(y0f + 1.0f, env.height);
    auto texel = [&](uint32_t x, uint32_t y) {
        const float* p = env.rgba + ((size_t)y * env.width + x) * 4;
        return V(p[0], p[1], p[2]);
    };
    vec3 top = texel(x0, y0) * (1.0f - wx) + texel(x1, y0) * wx;
    vec3 bottom = texel(x0, y1) * (1.0f - wx) + texel(x1, y1) * wx;
    return top * (1.0f - wy) + bottom * wy;
}

struct EnvironmentSample {
    vec3 direction, radiance;
    float pdf;
};

EnvironmentSample sample_environment(Ctx& ctx, uint32_t env, uint32_t* rng) {
    const Job& job = *ctx.job;
    const RtEnvironmentDesc& m = job.envs[env];
    uint32_t index = random_index_in_environment(job, env, rng);
    uint32_t x = index % m.width;
    uint32_t y = index / m.width;
    float jitter_x = random_uniform(rng);
    float jitter_y = random_uniform(rng);
    vec2 uv = {((float)x + jitter_x) / (float)m.width, ((float)y + jitter_y) / (float)m.height};
    vec3 direction = equirectangular_uv_to_direction(uv);
    vec3 radiance = texture_sample_bilinear(ctx, job.envs[job.env_index], uv);
    float pmf = environment_alias_entry(job, index, env).pmf;
    float delta_solid_angle = environment_pixel_solid_angle(uv.y, m);
    float pdf = pmf / delta_solid_angle;
    return EnvironmentSample{direction, radiance, pdf};
}

vec3 sky_light(Ctx& ctx, vec3 d) {
    vec2 uv = direction_to_equirectangular_uv(d);
    return texture_sample_bilinear(ctx, ctx.job->envs[ctx.job->env_index], uv);
}

// ---- BSDF: shader.wgsl:834-1210 ----------------------------------------------------
struct BsdfMaterial {
    vec3 color;
    float metallic, alpha;
    vec3 f0, emission;
};

vec3 surface_f0(const RtMaterial& m) {
    return lerp_vec3f(V(0.04f, 0.04f, 0.04f), ld3(m.color), saturate(m.metallic));
}
BsdfMaterial make_bsdf_material(const RtMaterial& m) {
    float alpha = std::fmax(0.001f, m.roughness * m.roughness);
    return BsdfMaterial{ld3(m.color), m.metallic, alpha, surface_f0(m), ld3(m.emission)};
}
vec3 surface_kd(const BsdfMaterial& m) {
    vec3 kd0 = m.color * (1.0f - saturate(m.metallic));
    return kd0 * (1.0f - max_component(m.f0));
}
float luminance(vec3 c) { return (0.2126f * c.x + 0.7152f * c.y) + 0.0722f * c.z; }

vec3 sample_cosine_hemisphere(vec2 s) {
    float r = std::sqrt(s.x);
    float phi = TWO_PI * s.y;
    float x = r * rt_cos(phi);
    float y = r * rt_sin(phi);
    float z = std::sqrt(std::fmax(0.0f, 1.0f - x * x - y * y));
    return V(x, y, z);
}
vec2 sample_uniform_disk(vec2 s) {
    float radius = std::sqrt(s.x);
    float azimuth = TWO_PI * s.y;
    return vec2{radius * rt_cos(azimuth), radius * rt_sin(azimuth)};
}
float pdf_cosine_hemisphere(vec3 wi) {
    if (wi.z <= 0.0f) return 0.0f;
    return wi.z / PI;
}
float d_ggx(float ndh, float alpha) {
    float alpha_2 = alpha * alpha;
    float denominator = (ndh * ndh) * (alpha_2 - 1.0f) + 1.0f;
    return alpha_2 / (PI * denominator * denominator);
}
float lambda_ggx(float ndv, float alpha) {
    float ndv2 = ndv * ndv;
    return (std::sqrt(1.0f + alpha * alpha * (1.0f - ndv2) / ndv2) - 1.0f) / 2.0f;
}
float g1_ggx(float ndv, float alpha) { return 1.0f / (1.0f + lambda_ggx(ndv, alpha)); }
float g_smith_ggx(float ndo, float ndi, float alpha) { return g1_ggx(ndo, alpha) * g1_ggx(ndi, alpha); }
float pdf_ggx_half_vector_visible(vec3 h, vec3 wo, float alpha) {
    float ndh = h.z, ndo = wo.z;
    if (ndh <= 0.0f) return 0.0f;
    return d_ggx(ndh, alpha) * g1_ggx(ndo, alpha) * std::fmax(0.0f, dot(wo, h)) / ndo;
}
vec3 sample_ggx_visible_half_vector(vec2 s, vec3 wo, float alpha) {
    vec3 vs = normalize(wo * V(alpha, alpha, 1.0f));
    float len2 = dot2(vec2{vs.x, vs.y}, vec2{vs.x, vs.y});
    vec3 tangent_x = len2 > 0.0f ? V(-vs.y, vs.x, 0.0f) * inverse_sqrt(len2) : V(1.0f, 0.0f, 0.0f);
    vec3 tangent_y = cross(vs, tangent_x);
    vec2 disk = sample_uniform_disk(s);
    disk.y = lerp_f32(std::sqrt(std::fmax(0.0f, 1.0f - disk.x * disk.x)), disk.y, vs.z);
    vec3 hs = (disk.x * tangent_x + disk.y * tangent_y) +
              std::sqrt(std::fmax(0.0f, 1.0f - disk.x * disk.x - disk.y * disk.y)) * vs;
    return normalize(V(alpha * hs.x, alpha * hs.y, std::fmax(0.0f, hs.z)));
}
vec3 f_schlick(vec3 f0, float cos_theta) {
    float x = 1.0f - saturate(cos_theta);
    float x_2 = x * x;
    float x_5 = x_2 * x_2 * x;
    return f0 + (V(1.0f) - f0) * x_5;
}
vec3 bsdf_eval_local(vec3 wo, vec3 wi, const BsdfMaterial& m) {
    if (wo.z <= 0.0f || wi.z <= 0.0f) return V(0.0f);
    float ndo = wo.z, ndi = wi.z;
    vec3 h = normalize(wo + wi);
    float ndh = saturate(h.z);
    float D = d_ggx(ndh, m.alpha);
    float G = g_smith_ggx(ndo, ndi, m.alpha);
    vec3 F = f_schlick(m.f0, dot(h, wo));
    vec3 fs = ((D * G) / (4.0f * ndo * ndi)) * F;
    vec3 kd = surface_kd(m);
    vec3 fd = kd * INV_PI;  // kd * (1 / PI): abstract-float fold
    return fd + fs;
}
float pdf_specular_wi_visible(vec3 wo, vec3 wi, float alpha) {
    if (wo.z <= 0.0f || wi.z <= 0.0f) return 0.0f;
    vec3 h = normalize(wo + wi);
    float wo_dot_h = std::fabs(dot(wo, h));
    if (wo_dot_h <= 0.0f) return 0.0f;
    return pdf_ggx_half_vector_visible(h, wo, alpha) / (4.0f * wo_dot_h);
}
float bsdf_pdf_local(vec3 wo, vec3 wi, const BsdfMaterial& m) {
    if (wo.z <= 0.0f || wi.z <= 0.0f) return 0.0f;
    float ps = saturate(luminance(m.f0));
    float pd = 1.0f - ps;
    return pd * pdf_cosine_hemisphere(wi) + ps * pdf_specular_wi_visible(wo, wi, m.alpha);
}

struct BsdfSample {
    vec3 ray_direction, scattering;
    float pdf;
};
BsdfSample bsdf_sample(const Ray& ray, vec3 n, const BsdfMaterial& m, uint32_t* rng) {
    vec3 wo_world = -ray.direction;
    if (dot(n, wo_world) <= 0.0f) return BsdfSample{V(0, 0, 0), V(0, 0, 1), 0.0f};
    Frame frame = make_frame(n);
    vec3 wo = to_frame_local(frame, wo_world);
    if (wo.z <= 0.0f) return BsdfSample{V(0, 0, 0), V(0, 1, 0), 0.0f};

    float ps = saturate(luminance(m.f0));
    float pd = 1.0f - ps;
    vec3 wi;
    float sample = random_uniform(rng);
    if (sample < pd) {
        float sx = sample / std::fmax(pd, 1.e-6f);
        float sy = random_uniform(rng);
        wi = sample_cosine_hemisphere(vec2{sx, sy});
    } else {
        float sx = (sample - pd) / std::fmax(ps, 1.e-6f);
        float sy = random_uniform(rng);
        vec3 h = sample_ggx_visible_half_vector(vec2{sx, sy}, wo, m.alpha);
        wi = reflect(-wo, h);
        if (wi.z <= 0.0f) return BsdfSample{V(1, 0, 0), V(1, 0, 0), 0.0f};
    }
    vec3 scattering = bsdf_eval_local(wo, wi, m);
    float pdf = bsdf_pdf_local(wo, wi, m);
    vec3 wi_world = to_frame_world(frame, wi);
    if (dot(n, wi_world) < 0.0f) return BsdfSample{V(0, 0, 0), V(0, 1, 0), 0.0f};
    return BsdfSample{wi_world, scattering, pdf};
}

float power_heuristic(float a, float b) {
    float a2 = a * a, b2 = b * b;
    return a2 / (a2 + b2);
}

// shader.wgsl:1213-1303
vec3 trace_ray(Ctx& ctx, Ray ray, uint32_t* rng) {
    const Job& job = *ctx.job;
    vec3 incoming_light = V(0.0f);
    vec3 throughput = V(1.0f);
    float last_sample_pdf = 1.0f;
    for (uint32_t bounce = 0; bounce < job.max_bounces; bounce++) {
        uint32_t slot;
        ctx.c.rays++;
        HitInfo info = cast_ray(ctx, ray, &slot);
        if (!info.did_hit) {
            vec3 environment_light = sky_light(ctx, ray.direction);
            float pdf = environment_direction_pdf(job, ray.direction, job.env_index);
            float weight = power_heuristic(last_sample_pdf, pdf);
            incoming_light = incoming_light + throughput * environment_light * weight;
            break;
        }
        ctx.c.hits++;
        if (slot < 0x80000000u && job.scene.primitives[slot].primitive_type == 2) ctx.c.tri_hits++;
        BsdfMaterial material = make_bsdf_material(job.scene.materials[info.material_id]);
        incoming_light = incoming_light + throughput * material.emission;
        {
            ctx.c.shaded_bounces++;
            EnvironmentSample environment = sample_environment(ctx, job.env_index, rng);
            vec3 wo_world = -ray.direction;
            vec3 wi_world = environment.direction;
            float cos_theta = std::fmax(0.0f, dot(info.normal, wi_world));
            bool lit = false;
            if (cos_theta > 0.0f && environment.pdf > 0.0f) {
                uint32_t shadow_slot;
                ctx.c.shadow_rays++;
                lit = !cast_ray_bvh(ctx, Ray{info.hit_point, environment.direction}, &shadow_slot,
                                    job.prune != 0)
                           .did_hit;
            }
            if (lit) {
                Frame frame = make_frame(info.normal);
                vec3 wo = to_frame_local(frame, wo_world);
                vec3 wi = to_frame_local(frame, wi_world);
                vec3 scattering = bsdf_eval_local(wo, wi, material);
                float pdf_bsdf = bsdf_pdf_local(wo, wi, material);
                float weight = power_heuristic(environment.pdf, pdf_bsdf);
                incoming_light = incoming_light + throughput * weight * environment.radiance *
                                                      scattering * cos_theta / environment.pdf;
            }
        }
        {
            BsdfSample sample = bsdf_sample(ray, info.normal, material, rng);
            if (sample.ray_direction.x == 0.0f && sample.ray_direction.y == 0.0f &&
                sample.ray_direction.z == 0.0f) {
                incoming_light = sample.scattering;
                break;
            }
            if (sample.pdf <= 0.0f) break;
            float cos_theta = std::fmax(0.0f, dot(info.normal, sample.ray_direction));
            throughput = throughput * (sample.scattering * (cos_theta / sample.pdf));
            if (length(throughput) < 0.001f) break;
            last_sample_pdf = sample.pdf;
            ray = Ray{info.hit_point, sample.ray_direction};
        }
    }
    return incoming_light;
}

// shader.wgsl:1305-1362 — RNG seeding and camera ray for one pixel-sample.
Ray camera_ray(const Job& job, uint32_t px, uint32_t py, uint32_t sample_count, uint32_t* rng) {
    uint32_t pixel_index = py * job.width + px;
    *rng = 0;
    salt_rng(rng, pixel_index);
    salt_rng(rng, sample_count);

    vec2 jitter = random_in_circle_uniform(rng);
    vec2 jp = {(float)px + jitter.x, (float)py + jitter.y};
    vec2 screen = {((jp.x / (float)job.width) * 2.0f - 1.0f) * 1.0f,
                   ((jp.y / (float)job.height) * 2.0f - 1.0f) * -1.0f};
    float max_y_component = rt_sin(job.camera.fov_y / 2.0f);
    float aspect_ratio = (float)job.width / (float)job.height;
    vec3 rcs = V(screen.x * max_y_component * aspect_ratio, screen.y * max_y_component, -1.0f);
    vec3 c0 = ld3(job.camera.rot[0]), c1 = ld3(job.camera.rot[1]), c2 = ld3(job.camera.rot[2]);
    vec3 world = (c0 * rcs.x + c1 * rcs.y) + c2 * rcs.z;
    return Ray{ld3(job.camera.pos), normalize(world)};
}

void add_counters(Counters& dst, const Counters& src) {
    uint64_t* d = (uint64_t*)&dst;
    const uint64_t* s = (const uint64_t*)&src;
    for (size_t i = 0; i < sizeof(Counters) / sizeof(uint64_t); i++) d[i] += s[i];
}

template <typename RowFn>
void parallel_rows(uint32_t height, int threads, RowFn&& fn) {
    if (threads <= 0) threads = (int)std::thread::hardware_concurrency();
    if (threads <= 0) threads = 1;
    std::atomic<uint32_t> next(0);
    std::vector<std::thread> pool;
    auto worker = [&](int tid) {
        for (;;) {
            uint32_t y = next.fetch_add(1);
            if (y >= height) break;
            fn(y, tid);
        }
    };
    for (int t = 1; t < threads; t++) pool.emplace_back(worker, t);
    worker(0);
    for (auto& th : pool) th.join();
}

bool job_ok(const Job* job) {
    return job && job->width && job->height && job->scene.bvh_node_count &&
           job->env_index < job->env_count;
}

}  // namespace

extern "C" {

typedef Job OracleJob;
typedef Counters OracleCounters;

// Accumulates samples [sample_begin, sample_begin + sample_count) into accum_rgba
// (W*H*4 floats) exactly like shader.wgsl:1364-1371: total += sample, alpha = 1.
int oracle_render(const OracleJob* job, uint32_t sample_begin, uint32_t sample_count,
                  float* accum_rgba, OracleCounters* counters, int threads) {
    if (!job_ok(job) || !accum_rgba) return 1;
    if (threads <= 0) threads = (int)std::thread::hardware_concurrency();
    if (threads <= 0) threads = 1;
    std::vector<Counters> per_thread((size_t)threads);
    for (auto& c : per_thread) std::memset(&c, 0, sizeof c);
    parallel_rows(job->height, threads, [&](uint32_t y, int tid) {
        Ctx ctx{job, Counters{}};
        std::memset(&ctx.c, 0, sizeof ctx.c);
        for (uint32_t x = 0; x < job->width; x++) {
            float* px = accum_rgba + ((size_t)y * job->width + x) * 4;
            for (uint32_t s = 0; s < sample_count; s++) {
                uint32_t rng;
                Ray ray = camera_ray(*job, x, y, sample_begin + s, &rng);
                ctx.c.paths++;
                vec3 light = trace_ray(ctx, ray, &rng);
                px[0] = px[0] + light.x;
                px[1] = px[1] + light.y;
                px[2] = px[2] + light.z;
                px[3] = 1.0f;
            }
        }
        add_counters(per_thread[(size_t)tid], ctx.c);
    });
    if (counters) {
        std::memset(counters, 0, sizeof *counters);
        for (auto& c : per_thread) add_counters(*counters, c);
    }
    return 0;
}

// Primary rays only: slot in `primitives` of the closest hit of cast_ray (or the
// fallback encoding / RT_MISS_SLOT) and its distance (0 on a miss).
int oracle_first_hit(const OracleJob* job, uint32_t sample_index, uint32_t* slot_out,
                     float* t_out, OracleCounters* counters, int threads) {
    if (!job_ok(job) || !slot_out) return 1;
    if (threads <= 0) threads = (int)std::thread::hardware_concurrency();
    if (threads <= 0) threads = 1;
    std::vector<Counters> per_thread((size_t)threads);
    for (auto& c : per_thread) std::memset(&c, 0, sizeof c);
    parallel_rows(job->height, threads, [&](uint32_t y, int tid) {
        Ctx ctx{job, Counters{}};
        std::memset(&ctx.c, 0, sizeof ctx.c);
        for (uint32_t x = 0; x < job->width; x++) {
            uint32_t rng, slot;
            Ray ray = camera_ray(*job, x, y, sample_index, &rng);
            ctx.c.paths++;
            ctx.c.rays++;
            HitInfo h = cast_ray(ctx, ray, &slot);
            size_t i = (size_t)y * job->width + x;
            slot_out[i] = h.did_hit ? slot : RT_MISS_SLOT;
            if (t_out) t_out[i] = h.did_hit ? h.distance : 0.0f;
        }
        add_counters(per_thread[(size_t)tid], ctx.c);
    });
    if (counters) {
        std::memset(counters, 0, sizeof *counters);
        for (auto& c : per_thread) add_counters(*counters, c);
    }
    return 0;
}

// Arbitrary rays (xyz triples): cast_ray for closest-hit rays (slot/fallback encoding as
// in oracle_first_hit, distance 0 on a miss) or, with shadow != 0, the occlusion query
// trace_ray makes (cast_ray_bvh; slot_out = 1 when anything is hit, else 0).
int oracle_cast_rays(const OracleJob* job, uint32_t n, const float* origins, const float* dirs,
                     int shadow, uint32_t* slot_out, float* t_out) {
    if (!job || !job->scene.bvh_node_count || !slot_out) return 1;
    Ctx ctx{job, Counters{}};
    std::memset(&ctx.c, 0, sizeof ctx.c);
    for (uint32_t i = 0; i < n; i++) {
        Ray ray{ld3(origins + 3 * i), ld3(dirs + 3 * i)};
        uint32_t slot = RT_MISS_SLOT;
        if (shadow) {
            HitInfo h = cast_ray_bvh(ctx, ray, &slot, job->prune != 0);
            slot_out[i] = h.did_hit ? 1u : 0u;
            if (t_out) t_out[i] = 0.0f;
        } else {
            HitInfo h = cast_ray(ctx, ray, &slot);
            slot_out[i] = h.did_hit ? slot : RT_MISS_SLOT;
            if (t_out) t_out[i] = h.did_hit ? h.distance : 0.0f;
        }
    }
    return 0;
}

// One pixel-sample's radiance (debug / unit tests).
int oracle_trace_pixel(const OracleJob* job, uint32_t x, uint32_t y, uint32_t sample_index,
                       float rgb[3]) {
    if (!job_ok(job) || !rgb || x >= job->width || y >= job->height) return 1;
    Ctx ctx{job, Counters{}};
    std::memset(&ctx.c, 0, sizeof ctx.c);
    uint32_t rng;
    Ray ray = camera_ray(*job, x, y, sample_index, &rng);
    vec3 light = trace_ray(ctx, ray, &rng);
    rgb[0] = light.x; rgb[1] = light.y; rgb[2] = light.z;
    return 0;
}

// Camera ray (origin, direction) and RNG state after ray generation.
int oracle_camera_ray(const OracleJob* job, uint32_t x, uint32_t y, uint32_t sample_index,
                      float origin[3], float direction[3], uint32_t* rng_state) {
    if (!job || !job->width || !job->height) return 1;
    uint32_t rng;
    Ray r = camera_ray(*job, x, y, sample_index, &rng);
    origin[0] = r.origin.x; origin[1] = r.origin.y; origin[2] = r.origin.z;
    direction[0] = r.direction.x; direction[1] = r.direction.y; direction[2] = r.direction.z;
    if (rng_state) *rng_state = rng;
    return 0;
}

// Known-answer hooks for the arithmetic contract.
float oracle_sin(float x) { return rt_sin(x); }
float oracle_cos(float x) { return rt_cos(x); }
float oracle_atan2(float y, float x) { return rt_atan2(y, x); }
float oracle_asin(float x) { return rt_asin(x); }
uint32_t oracle_rng_next(uint32_t* state) { return random_u32_uniform(state); }
float oracle_rng_uniform(uint32_t* state) { return random_uniform(state); }
void oracle_rng_seed(uint32_t pixel_index, uint32_t sample_count, uint32_t* state) {
    *state = 0;
    salt_rng(state, pixel_index);
    salt_rng(state, sample_count);
}
uint32_t oracle_counters_size(void) { return (uint32_t)sizeof(Counters); }
uint32_t oracle_job_size(void) { return (uint32_t)sizeof(Job); }

}  // extern "C"
