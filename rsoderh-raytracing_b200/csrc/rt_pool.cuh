// rt_pool.cuh — the render kernel: an SM-resident path pool scheduled as a wavefront
// inside each thread block.
//
// Why: in a thread-per-pixel loop a warp instruction had ~9 of 32 lanes active (ncu,
// profiles/r1_v2_*): rays of one warp need very different numbers of node / primitive
// tests, and shading, sky lookup and traversal alternate per lane. Here every persistent
// block keeps kSlots paths in shared memory (B200: 227 KB per CTA) and its warps pull
// 32-slot chunks from per-stage ring queues, so the lanes of a warp always execute the
// same stage:
//
//   FINISH   add the finished sample to its pixel, start the next sample (camera ray) or
//            fetch the next pixel from the global counter, or retire the slot
//   SHADE    closest hit found: material, next-event estimation set-up, BSDF sample
//   SKY      ray escaped: environment radiance with MIS weight
//   CONTINUE apply the shadow-ray result and start the next bounce
//   TRAVERSE resume the ray's BVH walk for up to kNodeBudget node tests (while-while:
//            descend to a leaf, test its primitives, repeat), brute-force fallback included
//
// A slot is in exactly one queue (or inside one warp) at any time. A pixel is owned by one
// slot for all of its samples, which are traced in sample order, so the accumulation order
// — and therefore every bit of the result — is the same as the reference's frame-by-frame
// accumulation, whatever the scheduling. The arithmetic is rt_device.cuh's, unchanged.
#pragma once

#include "rt_device.cuh"

namespace rtp {
using namespace rtd;

#ifndef RT_POOL_SLOTS
#define RT_POOL_SLOTS 1024
#endif
#ifndef RT_POOL_THREADS
#define RT_POOL_THREADS 512
#endif
#ifndef RT_NODE_BUDGET
#define RT_NODE_BUDGET 12
#endif

constexpr uint32_t kSlots = RT_POOL_SLOTS;  // power of two, <= 32768
constexpr uint32_t kThreads = RT_POOL_THREADS;
constexpr uint32_t kNodeBudget = RT_NODE_BUDGET;  // node tests per TRAVERSE visit
constexpr uint32_t kSmemStack = 8;  // deeper trees keep their stacks in a global scratch buffer
constexpr uint32_t kGlobalStack = 64;

// Queue order = scan priority (rarely-full stages first; TRAVERSE is the reservoir).
enum Stage : int { ST_FINISH = 0, ST_SHADE, ST_SKY, ST_CONTINUE, ST_TRAVERSE, ST_COUNT };

enum Field : uint32_t {
    F_OX, F_OY, F_OZ,   // ray origin
    F_DX, F_DY, F_DZ,   // direction of the ray being traversed
    F_IX, F_IY, F_IZ,   // 1 / direction
    F_NX, F_NY, F_NZ,   // direction of the next bounce while a shadow ray is in flight
    F_TX, F_TY, F_TZ,   // throughput
    F_LX, F_LY, F_LZ,   // radiance of the current sample
    F_PX, F_PY, F_PZ,   // pending next-event contribution (added if the shadow ray is free)
    F_RNG, F_PDF, F_PIXEL, F_SAMPLE, F_FLAGS,
    F_CUR, F_SP,                  // traversal cursor, stack size
    F_BT, F_BR, F_BU, F_BV,       // closest hit so far
    F_STACK,
    F_COUNT = F_STACK + kSmemStack
};

constexpr uint32_t FLAG_SHADOW = 1u;    // the ray in flight is a shadow ray
constexpr uint32_t FLAG_ENDS = 2u;      // the path ends after this bounce
constexpr uint32_t FLAG_FB_DONE = 4u;   // brute-force fallback already run for this ray
constexpr uint32_t FLAG_PLAIN = 8u;     // no component of 1/d is infinite
constexpr uint32_t FLAG_OCCLUDED = 16u; // shadow ray hit something
constexpr uint32_t FLAG_BOUNCE_SHIFT = 8u;

// Ring queues: 2 * kSlots entries of 16 bits = slot (10 bits) | lap tag (6 bits). Producers
// reserve a position with one atomicAdd and then store unconditionally (wait-free);
// a consumer that reserved position q waits until the entry carries q's lap tag, i.e. only
// for a producer that is between its reserve and its store. With twice as many entries as
// slots an entry cannot be overwritten before its consumer has read it.
constexpr uint32_t kRing = 2u * kSlots;
constexpr uint32_t kSlotBits = 10u;
static_assert(kSlots <= (1u << kSlotBits), "slot index must fit the ring entry");
constexpr uint32_t kSpinBudget = 1u << 22;  // waits last tens of cycles; this is ~0.1 s
constexpr uint32_t kNoPixel = 0xFFFFFFFFu;

struct PoolParams {
    SceneView scene;
    EnvView env;
    CameraParams camera;
    float4* accum;
    unsigned long long* counters;
    uint32_t* next_pixel;   // [0] global work counter, [1] watchdog status; zeroed before the launch
    uint32_t* stack_scratch;  // gridDim.x * kSlots * kGlobalStack words, or nullptr
    uint32_t width, height;
    uint32_t tiles_x, pixel_ids;  // 8x8-tile-major pixel enumeration
    uint32_t sample_begin, sample_count;
    uint32_t max_bounces;
    uint32_t node_count, record_count, material_count;  // bounds for RT_POOL_DEBUG checks
};

constexpr size_t pool_smem_bytes() {
    return (size_t)F_COUNT * kSlots * 4 + (size_t)ST_COUNT * kRing * 2 + (size_t)(2 * ST_COUNT + 2) * 4;
}

struct Pool {
    uint32_t* f;
    volatile uint16_t* ring;
    volatile uint32_t* wr;
    volatile uint32_t* rd;
    volatile uint32_t* live;
    volatile uint32_t* abort_flag;  // block-wide: a watchdog tripped, every warp leaves
    uint32_t* gstack;  // this block's scratch, or nullptr
    uint32_t* status;  // global watchdog word: non-zero = a wait exceeded its budget, everyone leaves

    RT_DEV float getf(uint32_t field, uint32_t slot) const { return __uint_as_float(f[field * kSlots + slot]); }
    RT_DEV uint32_t getu(uint32_t field, uint32_t slot) const { return f[field * kSlots + slot]; }
    RT_DEV void setf(uint32_t field, uint32_t slot, float v) { f[field * kSlots + slot] = __float_as_uint(v); }
    RT_DEV void setu(uint32_t field, uint32_t slot, uint32_t v) { f[field * kSlots + slot] = v; }
    RT_DEV V3 get3(uint32_t field, uint32_t slot) const {
        return v3(getf(field, slot), getf(field + 1, slot), getf(field + 2, slot));
    }
    RT_DEV void set3(uint32_t field, uint32_t slot, V3 v) {
        setf(field, slot, v.x);
        setf(field + 1, slot, v.y);
        setf(field + 2, slot, v.z);
    }
    RT_DEV void stack_push(uint32_t slot, uint32_t sp, uint32_t v) {
        if (gstack) gstack[(size_t)slot * kGlobalStack + sp] = v;
        else f[(F_STACK + sp) * kSlots + slot] = v;
    }
    RT_DEV uint32_t stack_pop(uint32_t slot, uint32_t sp) const {
        return gstack ? gstack[(size_t)slot * kGlobalStack + sp] : f[(F_STACK + sp) * kSlots + slot];
    }
};

// Starts the traversal of the ray (origin already stored) along direction d.
RT_DEV void begin_ray(Pool& pl, uint32_t slot, V3 d, uint32_t& flags) {
    V3 inv = v3(1.0f / d.x, 1.0f / d.y, 1.0f / d.z);
    pl.set3(F_DX, slot, d);
    pl.set3(F_IX, slot, inv);
    bool plain = fabsf(inv.x) < 3.0e38f && fabsf(inv.y) < 3.0e38f && fabsf(inv.z) < 3.0e38f;
    flags = (flags & ~(FLAG_FB_DONE | FLAG_PLAIN | FLAG_OCCLUDED)) | (plain ? FLAG_PLAIN : 0u);
    pl.setu(F_CUR, slot, 0u);
    pl.setu(F_SP, slot, 0u);
    pl.setf(F_BT, slot, RT_INFINITY);
    pl.setu(F_BR, slot, RT_NONE);
    pl.setf(F_BU, slot, 0.0f);
    pl.setf(F_BV, slot, 0.0f);
}

// ---- TRAVERSE: cast_ray_bvh (shader.wgsl:469-564) + cast_ray's fallback (583-598) ----
// Same walk as rtd::trace (reference visiting order, closest-hit pruning, any-hit exit for
// shadow rays, min/max slabs when 1/d is finite), made resumable: the cursor, stack and
// closest hit live in the slot, and a visit ends after kNodeBudget node tests. Leaves are
// always finished inside the visit that reached them.
template <bool COUNT>
RT_DEV int stage_traverse(const PoolParams& p, Pool& pl, uint32_t slot, Counts& cn) {
    const V3 o = pl.get3(F_OX, slot);
    const V3 d = pl.get3(F_DX, slot);
    const V3 inv = pl.get3(F_IX, slot);
    uint32_t flags = pl.getu(F_FLAGS, slot);
    const bool plain = (flags & FLAG_PLAIN) != 0u;
    const bool any_hit = (flags & FLAG_SHADOW) != 0u;
    uint32_t cur = pl.getu(F_CUR, slot);
    uint32_t sp = pl.getu(F_SP, slot);
    Hit best;
    best.t = pl.getf(F_BT, slot);
    best.rec = pl.getu(F_BR, slot);
    best.u = pl.getf(F_BU, slot);
    best.v = pl.getf(F_BV, slot);
    float limit = p.scene.prune ? fminf(RT_INFINITY, best.t * 1.001f) : RT_INFINITY;
    uint32_t budget = kNodeBudget;
    uint32_t leaf_begin = 0u, leaf_end = 0u;
    bool leaf_mixed = false;

    for (;;) {
        // ---- phase A: descend until this lane holds a leaf, runs out of nodes or budget ----
        while (cur != RT_NONE && leaf_begin == leaf_end && budget > 0u) {
            budget--;
            const float4 n0 = ld4(p.scene.nodes + 2u * (size_t)cur);
            const float4 n1 = ld4(p.scene.nodes + 2u * (size_t)cur + 1);
            if (COUNT) cn.nodes++;
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
            uint32_t next = RT_NONE;
            if (!(t0 > t1)) {
                const uint32_t idx = __float_as_uint(n0.w);
                const uint32_t meta = __float_as_uint(n1.w);
                const uint32_t len = meta & 0x1FFFFFFFu;
                if (len > 0u) {
                    leaf_begin = idx;
                    leaf_end = idx + len;
                    leaf_mixed = (meta & 0x20000000u) != 0u;
                } else {
                    const bool second_first = comp(inv, meta >> 30) < 0.0f;
                    pl.stack_push(slot, sp, second_first ? cur + 1u : idx);
                    sp++;
                    next = second_first ? idx : cur + 1u;
                }
            }
            if (next == RT_NONE && sp > 0u) {
                sp--;
                next = pl.stack_pop(slot, sp);
            }
            cur = next;
        }
        if (leaf_begin == leaf_end) {
            if (cur != RT_NONE) break;  // budget spent: suspend
            // BVH exhausted. cast_ray falls back to brute force only when nothing was hit.
            uint32_t fb = p.scene.sphere_count + p.scene.plane_count;
            if (any_hit || best.rec != RT_NONE || (flags & FLAG_FB_DONE) || fb == 0u) break;
            flags |= FLAG_FB_DONE;
            if (COUNT) cn.fallback++;
            leaf_begin = p.scene.slot_count;
            leaf_end = p.scene.slot_count + fb;
            leaf_mixed = true;
        }
        // ---- phase B: test the pending primitives ----
        if (leaf_mixed) {
            while (leaf_begin < leaf_end) {
                test_record<COUNT>(p.scene, leaf_begin, o, d, best, cn);
                leaf_begin++;
                if (any_hit && best.rec != RT_NONE) break;
            }
        } else {
            while (leaf_begin < leaf_end) {
                const float4* r = p.scene.records + 4u * (size_t)leaf_begin;
                float4 r0 = ld4(r), r1 = ld4(r + 1), r2 = ld4(r + 2);
                float t = 0.0f, u = 0.0f, v = 0.0f;
                if (COUNT) cn.tris++;
                if (test_triangle(o, d, r0, r1, r2, t, u, v) && t < best.t) {
                    best.t = t;
                    best.rec = leaf_begin;
                    best.u = u;
                    best.v = v;
                }
                leaf_begin++;
                if (any_hit && best.rec != RT_NONE) break;
            }
        }
        if (any_hit && best.rec != RT_NONE) {
            // Occluded: the rest of the walk cannot change the answer.
            pl.setu(F_FLAGS, slot, flags | FLAG_OCCLUDED);
            return (flags & FLAG_ENDS) ? ST_FINISH : ST_CONTINUE;
        }
        leaf_begin = leaf_end = 0u;
        if (p.scene.prune) limit = fminf(RT_INFINITY, best.t * 1.001f);
    }

    if (cur != RT_NONE) {
        // Suspended: park the walk in the slot and come back.
        pl.setu(F_CUR, slot, cur);
        pl.setu(F_SP, slot, sp);
        pl.setf(F_BT, slot, best.t);
        pl.setu(F_BR, slot, best.rec);
        pl.setf(F_BU, slot, best.u);
        pl.setf(F_BV, slot, best.v);
        pl.setu(F_FLAGS, slot, flags);
        return ST_TRAVERSE;
    }
    if (any_hit) return (flags & FLAG_ENDS) ? ST_FINISH : ST_CONTINUE;  // shadow ray is free
    if (best.rec == RT_NONE) return ST_SKY;
    pl.setf(F_BT, slot, best.t);
    pl.setu(F_BR, slot, best.rec);
    pl.setf(F_BU, slot, best.u);
    pl.setf(F_BV, slot, best.v);
    return ST_SHADE;
}

// ---- SKY: the ray escaped (shader.wgsl:1222-1231) ----
template <bool COUNT>
RT_DEV int stage_sky(const PoolParams& p, Pool& pl, uint32_t slot, Counts& cn) {
    const V3 d = pl.get3(F_DX, slot);
    float u, v;
    direction_to_uv(d, u, v);
    V3 sky = env_bilinear(p.env, u, v);
    if (COUNT) cn.env_lookups++;
    float pdf = env_direction_pdf(p.env, u, v);
    float w = power_heuristic(pl.getf(F_PDF, slot), pdf);
    V3 L = pl.get3(F_LX, slot) + pl.get3(F_TX, slot) * sky * w;
    pl.set3(F_LX, slot, L);
    return ST_FINISH;
}

// ---- SHADE: closest hit found (shader.wgsl:1233-1298) ----
template <bool COUNT>
RT_DEV int stage_shade(const PoolParams& p, Pool& pl, uint32_t slot, Counts& cn) {
    const V3 o = pl.get3(F_OX, slot);
    const V3 d = pl.get3(F_DX, slot);
    V3 T = pl.get3(F_TX, slot);
    V3 L = pl.get3(F_LX, slot);
    uint32_t rng = pl.getu(F_RNG, slot);
    uint32_t flags = pl.getu(F_FLAGS, slot);
    Hit hit;
    hit.t = pl.getf(F_BT, slot);
    hit.rec = pl.getu(F_BR, slot);
    hit.u = pl.getf(F_BU, slot);
    hit.v = pl.getf(F_BV, slot);
    V3 point, n;
    uint32_t mat_id;
    reconstruct_hit(p.scene, o, d, hit, point, n, mat_id);
    if (COUNT) {
        cn.hits++;
        if (hit.rec < p.scene.slot_count &&
            __float_as_int(ld4(p.scene.records + 4u * (size_t)hit.rec).w) == 2)
            cn.tri_hits++;
    }
    const Material m = load_material(p.scene, mat_id);
    L = L + T * m.emission;  // shader.wgsl:1236

    const V3 wo_world = -d;
    const Frame frame = make_frame(n);
    const V3 wo = to_local(frame, wo_world);

    // ---- next-event estimation: sample_environment (shader.wgsl:782-820) ----
    bool shadow = false;
    V3 shadow_dir = v3s(0.0f), pending = v3s(0.0f);
    {
        if (COUNT) cn.shaded++;
        uint32_t length = p.env.width * p.env.height;
        uint32_t index = min((uint32_t)(rng_uniform(rng) * (float)length), length - 1u);
        float4 entry = ld4(p.env.alias + index);
        uint32_t pick = rng_uniform(rng) < entry.x ? index : __float_as_uint(entry.y);
        uint32_t ex = pick % p.env.width, ey = pick / p.env.width;
        float jx = rng_uniform(rng), jy = rng_uniform(rng);
        float eu = ((float)ex + jx) / (float)p.env.width;
        float ev = ((float)ey + jy) / (float)p.env.height;
        float phi = (2.0f * eu - 1.0f) * RT_PI;
        float theta = RT_PI * ev;
        float st, ct, sph, cph;
        rt_sincos(theta, &st, &ct);
        rt_sincos(phi, &sph, &cph);
        V3 wi_world = v3(st * cph, ct, st * sph);
        float pmf = pick == index ? entry.z : ld4(p.env.alias + pick).z;
        float d_phi = RT_TWO_PI / (float)p.env.width;
        float d_theta = RT_PI / (float)p.env.height;
        float pdf_env = pmf / (d_phi * d_theta * fmaxf(1.0e-6f, st));
        float cos_theta = fmaxf(0.0f, dot(n, wi_world));
        if (COUNT) cn.env_lookups++;
        if (cos_theta > 0.0f && pdf_env > 0.0f) {
            // The contribution is computed now and added only if the shadow ray is free.
            shadow = true;
            shadow_dir = wi_world;
            V3 radiance = env_bilinear(p.env, eu, ev);
            V3 wi = to_local(frame, wi_world);
            V3 f;
            float pdf_bsdf;
            bsdf_eval_pdf(wo, wi, m, f, pdf_bsdf);
            float w = power_heuristic(pdf_env, pdf_bsdf);
            pending = T * w * radiance * f * cos_theta / pdf_env;
        }
    }

    // ---- bsdf_sample (shader.wgsl:1116-1202) + path update (1270-1298) ----
    bool done = false;
    V3 next_dir = v3s(0.0f);
    float last_pdf = 0.0f;
    if (dot(n, wo_world) <= 0.0f) {
        // The debug colour replaces everything gathered so far (shader.wgsl:1271-1276); the
        // shadow ray is still traced (the shader traces it before bsdf_sample) but adds 0.
        L = v3(0.f, 0.f, 1.f);
        pending = v3s(0.0f);
        done = true;
    } else if (wo.z <= 0.0f) {
        L = v3(0.f, 1.f, 0.f);
        pending = v3s(0.0f);
        done = true;
    } else {
        float sample_u = rng_uniform(rng);
        V3 wi;
        bool blocked = false;
        if (sample_u < m.p_diff) {
            float sx = sample_u / fmaxf(m.p_diff, 1.e-6f);
            float sy = rng_uniform(rng);
            float r = sqrtf(sx);
            float s, c;
            rt_sincos(RT_TWO_PI * sy, &s, &c);
            float x = r * c, y = r * s;
            wi = v3(x, y, sqrtf(fmaxf(0.0f, 1.0f - x * x - y * y)));
        } else {
            float sx = (sample_u - m.p_diff) / fmaxf(m.p_spec, 1.e-6f);
            float sy = rng_uniform(rng);
            V3 h = sample_ggx_vndf(sx, sy, wo, m.alpha);
            wi = reflect(-wo, h);
            blocked = wi.z <= 0.0f;
        }
        if (blocked) {
            done = true;
        } else {
            V3 f;
            float pdf;
            bsdf_eval_pdf(wo, wi, m, f, pdf);
            V3 wi_world = to_world(frame, wi);
            if (dot(n, wi_world) < 0.0f) {
                L = v3(0.f, 1.f, 0.f);
                pending = v3s(0.0f);
                done = true;
            } else if (pdf <= 0.0f) {
                done = true;
            } else {
                float cos_theta = fmaxf(0.0f, dot(n, wi_world));
                T = T * (f * (cos_theta / pdf));
                if (length(T) < 0.001f) {
                    done = true;
                } else {
                    last_pdf = pdf;
                    next_dir = wi_world;
                }
            }
        }
    }

    uint32_t bounce = (flags >> FLAG_BOUNCE_SHIFT) + 1u;
    if (bounce >= p.max_bounces) done = true;
    flags = (flags & 0xFFu & ~(FLAG_ENDS | FLAG_SHADOW)) | (bounce << FLAG_BOUNCE_SHIFT) |
            (done ? FLAG_ENDS : 0u);

    pl.set3(F_OX, slot, point);  // origin of both the shadow ray and the next bounce
    pl.set3(F_TX, slot, T);
    pl.set3(F_LX, slot, L);
    pl.setu(F_RNG, slot, rng);
    if (!done) {
        pl.setf(F_PDF, slot, last_pdf);
        pl.set3(F_NX, slot, next_dir);
    }
    int dest;
    if (shadow) {
        cn.shadow_rays++;
        flags |= FLAG_SHADOW;
        pl.set3(F_PX, slot, pending);
        begin_ray(pl, slot, shadow_dir, flags);
        dest = ST_TRAVERSE;
    } else {
        dest = done ? ST_FINISH : ST_CONTINUE;
    }
    pl.setu(F_FLAGS, slot, flags);
    return dest;
}

// Adds the pending next-event contribution if a shadow ray was traced and found nothing.
RT_DEV void settle_shadow(Pool& pl, uint32_t slot, uint32_t& flags) {
    if ((flags & FLAG_SHADOW) && !(flags & FLAG_OCCLUDED)) {
        V3 L = pl.get3(F_LX, slot) + pl.get3(F_PX, slot);
        pl.set3(F_LX, slot, L);
    }
    flags &= ~(FLAG_SHADOW | FLAG_OCCLUDED);
}

// ---- CONTINUE: next bounce ----
RT_DEV int stage_continue(Pool& pl, uint32_t slot, Counts& cn) {
    uint32_t flags = pl.getu(F_FLAGS, slot);
    settle_shadow(pl, slot, flags);
    cn.rays++;
    begin_ray(pl, slot, pl.get3(F_NX, slot), flags);
    pl.setu(F_FLAGS, slot, flags);
    return ST_TRAVERSE;
}

// 8x8-tile-major pixel enumeration: consecutive ids are spatially close.
RT_DEV bool pixel_of_id(const PoolParams& p, uint32_t id, uint32_t& px, uint32_t& py) {
    uint32_t tile = id >> 6, in = id & 63u;
    px = (tile % p.tiles_x) * 8u + (in & 7u);
    py = (tile / p.tiles_x) * 8u + (in >> 3);
    return px < p.width && py < p.height;
}

// ---- FINISH: accumulate (shader.wgsl:1367-1371), then next sample / next pixel / retire ----
RT_DEV int stage_finish(const PoolParams& p, Pool& pl, uint32_t slot, float sin_half_fov, Counts& cn) {
    uint32_t flags = pl.getu(F_FLAGS, slot);
    settle_shadow(pl, slot, flags);
    uint32_t pixel = pl.getu(F_PIXEL, slot);
    uint32_t sample = pl.getu(F_SAMPLE, slot);
    const uint32_t sample_end = p.sample_begin + p.sample_count;
    if (pixel != kNoPixel) {
        V3 L = pl.get3(F_LX, slot);
        float4 a = __ldcg(p.accum + pixel);
        a.x = a.x + L.x;
        a.y = a.y + L.y;
        a.z = a.z + L.z;
        a.w = 1.0f;
        __stcg(p.accum + pixel, a);
        sample++;
    }
    uint32_t px = 0, py = 0;
    if (pixel == kNoPixel || sample >= sample_end) {
        // This pixel is complete: take the next one from the frame-wide counter.
        uint32_t id = atomicAdd(p.next_pixel, 1u);
        if (id >= p.pixel_ids) {
            atomicSub((uint32_t*)pl.live, 1u);  // nothing left: retire the slot
            return -1;
        }
        if (!pixel_of_id(p, id, px, py)) {
            // Padding of a partial tile: come back for another id.
            pl.setu(F_PIXEL, slot, kNoPixel);
            pl.setu(F_FLAGS, slot, 0u);
            return ST_FINISH;
        }
        pixel = py * p.width + px;
        sample = p.sample_begin;
        pl.setu(F_PIXEL, slot, pixel);
    } else {
        px = pixel % p.width;
        py = pixel / p.width;
    }
    pl.setu(F_SAMPLE, slot, sample);
    uint32_t rng;
    V3 o, d;
    camera_ray(p.camera, sin_half_fov, px, py, p.width, p.height, sample, rng, o, d);
    cn.paths++;
    cn.rays++;
    pl.set3(F_OX, slot, o);
    pl.set3(F_TX, slot, v3s(1.0f));
    pl.set3(F_LX, slot, v3s(0.0f));
    pl.setu(F_RNG, slot, rng);
    pl.setf(F_PDF, slot, 1.0f);
    flags = 0u;
    begin_ray(pl, slot, d, flags);
    pl.setu(F_FLAGS, slot, flags);
    return ST_TRAVERSE;
}

RT_DEV uint32_t ring_entry(uint32_t position, uint32_t slot) {
    return slot | (((position / kRing) & 63u) << kSlotBits);
}

// Publishes every lane's slot to its destination queue (dest < 0: nothing to publish).
// DEST_MASK lists the stages the stage just run can route to.
template <uint32_t DEST_MASK>
RT_DEV void publish(Pool& pl, int dest, uint32_t slot, uint32_t lane) {
    __threadfence_block();  // slot state before queue entry
#pragma unroll
    for (int s = 0; s < ST_COUNT; s++) {
        if (!(DEST_MASK & (1u << s))) continue;
        unsigned m = __ballot_sync(0xFFFFFFFFu, dest == s);
        if (m) {
            int leader = __ffs(m) - 1;
            uint32_t base = 0;
            if ((int)lane == leader) base = atomicAdd((uint32_t*)&pl.wr[s], (uint32_t)__popc(m));
            base = __shfl_sync(0xFFFFFFFFu, base, leader);
            if (dest == s) {
                uint32_t pos = base + __popc(m & ((1u << lane) - 1u));
                pl.ring[(size_t)s * kRing + (pos & (kRing - 1u))] = (uint16_t)ring_entry(pos, slot);
            }
        }
    }
}

template <bool COUNT>
__global__ void __launch_bounds__(kThreads, 1) render_pool_kernel(const PoolParams p) {
    extern __shared__ __align__(16) uint32_t smem[];
    Pool pl;
    pl.f = smem;
    pl.ring = (volatile uint16_t*)(smem + (size_t)F_COUNT * kSlots);
    pl.wr = (volatile uint32_t*)(pl.ring + (size_t)ST_COUNT * kRing);
    pl.rd = pl.wr + ST_COUNT;
    pl.live = pl.rd + ST_COUNT;
    pl.abort_flag = pl.live + 1;
    pl.gstack = p.stack_scratch ? p.stack_scratch + (size_t)blockIdx.x * kSlots * kGlobalStack : nullptr;
    pl.status = p.next_pixel + 1;

    const uint32_t lane = threadIdx.x & 31u;
    // Tag 63 is not lap 0's tag, so untouched entries read as "not yet written".
    for (uint32_t i = threadIdx.x; i < ST_COUNT * kRing; i += kThreads) pl.ring[i] = (uint16_t)0xFFFFu;
    __syncthreads();
    // Every slot starts as "pixel finished": FINISH hands it its first pixel.
    for (uint32_t s = threadIdx.x; s < kSlots; s += kThreads) {
        pl.setu(F_PIXEL, s, kNoPixel);
        pl.setu(F_FLAGS, s, 0u);
        pl.setu(F_SAMPLE, s, 0u);
        pl.ring[(size_t)ST_FINISH * kRing + s] = (uint16_t)ring_entry(s, s);
    }
    if (threadIdx.x == 0) {
        for (int s = 0; s < ST_COUNT; s++) {
            pl.wr[s] = 0u;
            pl.rd[s] = 0u;
        }
        pl.wr[ST_FINISH] = kSlots;
        *pl.live = kSlots;
        *pl.abort_flag = 0u;
    }
    __syncthreads();

    const float sin_half_fov = rt_sin(p.camera.fov_y / 2.0f);
    Counts cn;
    memset(&cn, 0, sizeof cn);
    uint32_t idle = 0;

    int sticky = ST_TRAVERSE;  // the stage this warp ran last: its code is hot in the I-cache
    for (;;) {
        // ---- lane 0 decides for the warp: reserve a chunk, idle, or leave. Every lane acts
        //      on the broadcast decision only: after a *_sync the lanes of a warp may still
        //      run as separate groups, so each reading `live` itself could split the warp
        //      between "leave" and "stay".
        //      Choice: a full chunk (32) of the stage run last, else a full chunk of the first
        //      queue in priority order, else the largest partial chunk. ----
        int stage = -1;  // >= 0 stage to run, -1 idle, -2 all slots retired, -3 abort
        uint32_t base = 0, n = 0;
        if (lane == 0) {
            if (*pl.abort_flag != 0u) {
                stage = -3;
            } else {
                for (;;) {
                    int pick = -1;
                    uint32_t pick_avail = 0, pick_rd = 0;
                    {
                        uint32_t r = pl.rd[sticky];
                        uint32_t avail = pl.wr[sticky] - r;
                        if (avail >= 32u) {
                            pick = sticky;
                            pick_avail = 32u;
                            pick_rd = r;
                        }
                    }
                    if (pick < 0) {
                        for (int s = 0; s < ST_COUNT; s++) {
                            uint32_t r = pl.rd[s];
                            uint32_t avail = pl.wr[s] - r;
                            if (avail >= 32u) {
                                pick = s;
                                pick_avail = 32u;
                                pick_rd = r;
                                break;
                            }
                            if (avail > pick_avail) {
                                pick = s;
                                pick_avail = avail;
                                pick_rd = r;
                            }
                        }
                    }
                    if (pick < 0) break;
                    if (atomicCAS((uint32_t*)&pl.rd[pick], pick_rd, pick_rd + pick_avail) == pick_rd) {
                        stage = pick;
                        base = pick_rd;
                        n = pick_avail;
                        break;
                    }
                }
                if (stage >= 0) {
                    idle = 0;
                } else if (*pl.live == 0u) {
                    stage = -2;
                } else {
                    ++idle;
                    if (idle > (kSpinBudget >> 2)) {  // ~0.3 s without work while slots are live
                        atomicCAS(pl.status, 0u, 0x200u);
                        *pl.abort_flag = 1u;
                        stage = -3;
                    } else if ((idle & 1023u) == 0u && *(volatile uint32_t*)pl.status != 0u) {
                        *pl.abort_flag = 1u;  // another block tripped its watchdog
                        stage = -3;
                    }
                }
            }
        }
        stage = __shfl_sync(0xFFFFFFFFu, stage, 0);
        base = __shfl_sync(0xFFFFFFFFu, base, 0);
        n = __shfl_sync(0xFFFFFFFFu, n, 0);
        if (stage <= -2) break;
        if (stage == -1) {
            __nanosleep(200);
            continue;
        }
        sticky = stage;

        uint32_t slot = 0;
        const bool has = lane < n;
        if (has) {
            const uint32_t pos = base + lane;
            volatile uint16_t* e = pl.ring + (size_t)stage * kRing + (pos & (kRing - 1u));
            const uint32_t tag = (pos / kRing) & 63u;
            uint32_t v;
            // Reserved by its producer, which is about to store it.
            for (uint32_t spin = 0; ((v = *e) >> kSlotBits) != tag; spin++)
                if (spin > kSpinBudget) {
                    atomicCAS(pl.status, 0u, 0x100u + (uint32_t)stage);
                    *pl.abort_flag = 1u;
                    break;
                }
            slot = v & ((1u << kSlotBits) - 1u);
        }
        if (__any_sync(0xFFFFFFFFu, *pl.abort_flag != 0u)) break;
        __threadfence_block();  // queue entry before slot state

        int dest = -1;
        constexpr uint32_t B_FIN = 1u << ST_FINISH, B_SHA = 1u << ST_SHADE, B_SKY = 1u << ST_SKY,
                           B_CON = 1u << ST_CONTINUE, B_TRA = 1u << ST_TRAVERSE;
        switch (stage) {  // warp-uniform
            case ST_TRAVERSE:
                if (has) dest = stage_traverse<COUNT>(p, pl, slot, cn);
                __syncwarp();
                publish<B_TRA | B_SHA | B_SKY | B_CON | B_FIN>(pl, dest, slot, lane);
                break;
            case ST_SHADE:
                if (has) dest = stage_shade<COUNT>(p, pl, slot, cn);
                __syncwarp();
                publish<B_TRA | B_CON | B_FIN>(pl, dest, slot, lane);
                break;
            case ST_SKY:
                if (has) dest = stage_sky<COUNT>(p, pl, slot, cn);
                __syncwarp();
                publish<B_FIN>(pl, dest, slot, lane);
                break;
            case ST_CONTINUE:
                if (has) dest = stage_continue(pl, slot, cn);
                __syncwarp();
                publish<B_TRA>(pl, dest, slot, lane);
                break;
            default:
                if (has) dest = stage_finish(p, pl, slot, sin_half_fov, cn);
                __syncwarp();
                publish<B_TRA | B_FIN>(pl, dest, slot, lane);
                break;
        }
    }

    // ---- counters ----
    const uint32_t vals[12] = {cn.paths, cn.rays,  cn.shadow_rays, cn.nodes,  cn.tris,        cn.spheres,
                               cn.planes, cn.hits, cn.tri_hits,    cn.shaded, cn.env_lookups, cn.fallback};
    const int nvals = COUNT ? 12 : 3;
    for (int i = 0; i < nvals; i++) {
        uint32_t v = vals[i];
        for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(0xFFFFFFFFu, v, off);
        if (lane == 0u && v) atomicAdd(p.counters + i, (unsigned long long)v);
    }
}

}  // namespace rtp
