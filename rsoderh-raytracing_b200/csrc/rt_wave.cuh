// rt_wave.cuh — the wavefront render pipeline: the path loop of shader.wgsl:1213-1303 cut at
// its ray casts and run as two small kernels per "wave".
//
// Why: as one kernel (rt_fast.cuh's render_fast_kernel) the loop body is ~60 KB of SASS — far
// beyond the 32 KB instruction cache level — and a warp's lanes sit in different parts of
// it: ncu showed 40 % issue utilisation with "no instruction" as the top stall and 8 to 10
// of 32 lanes active. Here
//   * `wave_logic_kernel` (one thread per path slot) consumes last wave's results — settles
//     the shadow ray, adds sky light or shades the hit, accumulates finished samples and
//     starts the pixel's next one — then runs the analytic pre-pass for the rays it emits,
//     all lanes together, and queues them;
//   * `wave_trace_kernel` is the triangle walk alone (a few KB of code): persistent warps
//     whose lanes each pull the next ray from the queue the moment their own is finished,
//     so closest-hit and shadow rays of the whole frame share the lanes.
// The next bounce's ray and the shadow ray of the same hit do not depend on each other and
// travel in the same wave. Path state lives in global memory, sized so that one wave's
// working set (~170 B per slot) stays inside the 126 MB L2.
//
// A slot owns one pixel for all of that pixel's samples and traces them in sample order,
// then moves to its next pixel (slot + k * slot_count in an 8x8-tile-major numbering), so
// the accumulation order — and with it every bit of the result — is the reference's
// frame-by-frame order (shader.wgsl:1367-1371) regardless of scheduling.
#pragma once

#include "rt_fast.cuh"

namespace rtw {
using namespace rtd;
using namespace rtf;

constexpr uint32_t kShadowBit = 0x80000000u;  // queue entry: the slot's shadow ray
constexpr int kLogicThreads = 128;
constexpr int kTraceThreads = 256;

enum SlotState : uint32_t { SLOT_RETIRED = 0, SLOT_NEW = 1, SLOT_TRACED = 2, SLOT_ENDING = 3 };
constexpr uint32_t kFlagShadow = 1u;  // a shadow ray was cast for the last shaded hit

// Control block, one per wave parity: the logic kernel of wave w fills ctrl[w & 1] and
// clears ctrl[(w + 1) & 1] for the next wave.
struct WaveCtrl {
    uint32_t queued;  // rays appended to the queue
    uint32_t head;    // rays handed out by the trace kernel
    uint32_t live;    // slots still working after this wave's logic pass
    uint32_t pad;
};

struct WaveBuffers {
    float4* ray_o;    // [S] origin of the slot's rays, .w = closest analytic distance from the pre-pass
    float4* ray_d;    // [S] direction of the closest-hit ray, .w = bits rec of that analytic hit
    float4* sray_d;   // [S] direction of the shadow ray
    float2* fb;       // [S] brute-force fallback winner of the closest-hit ray (t, bits index)
    float4* hit;      // [S] (t, bits rec, u, v) written by the trace kernel
    uint32_t* occl;   // [S] 1 when the shadow ray is blocked
    float4* pT;       // [S] throughput, last_pdf
    float4* pL;       // [S] radiance of the current sample, bits rng
    float4* pP;       // [S] pending next-event light, bits flags
    uint4* pI;        // [S] pixel id, sample, bounce, state
    uint32_t* queue;  // [2S]
    WaveCtrl* ctrl;   // [2]
};

struct WaveParams {
    FastScene scene;
    EnvView env;
    CameraParams camera;
    WaveBuffers b;
    float4* accum;
    unsigned long long* counters;
    uint32_t width, height;
    uint32_t tiles_x, pixel_ids;  // 8x8-tile-major pixel numbering (ids of partial tiles included)
    uint32_t slots;
    uint32_t sample_begin, sample_count;
    uint32_t max_bounces;
    uint32_t wave;
};

RT_DEVONLY bool pixel_of_id(const WaveParams& p, uint32_t id, uint32_t& px, uint32_t& py) {
    uint32_t tile = id >> 6, in = id & 63u;
    px = (tile % p.tiles_x) * 8u + (in & 7u);
    py = (tile / p.tiles_x) * 8u + (in >> 3);
    return px < p.width && py < p.height;
}

// First valid pixel id at or after `id` in this slot's sequence, or kNone.
RT_DEVONLY uint32_t next_pixel_id(const WaveParams& p, uint32_t id, uint32_t& px, uint32_t& py) {
    while (id < p.pixel_ids) {
        if (pixel_of_id(p, id, px, py)) return id;
        id += p.slots;
    }
    return kNone;
}

__global__ void wave_init_kernel(const WaveParams p) {
    uint32_t slot = blockIdx.x * blockDim.x + threadIdx.x;
    if (slot < p.slots) p.b.pI[slot] = make_uint4(slot, p.sample_begin, 0u, SLOT_NEW);
    if (slot < 2) {
        p.b.ctrl[slot].queued = 0u;
        p.b.ctrl[slot].head = 0u;
        p.b.ctrl[slot].live = 0u;
    }
}

// Appends one entry per requesting lane with a single atomic per warp.
RT_DEVONLY void queue_push(const WaveParams& p, WaveCtrl* ctrl, bool want, uint32_t entry) {
    const unsigned m = __ballot_sync(0xFFFFFFFFu, want);
    if (!m) return;
    const uint32_t lane = threadIdx.x & 31u;
    const int leader = __ffs(m) - 1;
    uint32_t base = 0;
    if ((int)lane == leader) base = atomicAdd(&ctrl->queued, (uint32_t)__popc(m));
    base = __shfl_sync(0xFFFFFFFFu, base, leader);
    if (want) p.b.queue[base + __popc(m & ((1u << lane) - 1u))] = entry;
}

#ifndef RT_WAVE_LOGIC_MIN_BLOCKS
#define RT_WAVE_LOGIC_MIN_BLOCKS 4
#endif

__global__ void __launch_bounds__(kLogicThreads, RT_WAVE_LOGIC_MIN_BLOCKS) wave_logic_kernel(const WaveParams p) {
    __shared__ float4 spre[kPreMaxFloat4];
    stage_pre(p.scene, spre);

    WaveCtrl* ctrl = p.b.ctrl + (p.wave & 1u);
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        WaveCtrl* other = p.b.ctrl + ((p.wave + 1u) & 1u);
        other->queued = 0u;
        other->head = 0u;
        other->live = 0u;
    }

    const uint32_t slot = blockIdx.x * kLogicThreads + threadIdx.x;
    uint4 I = slot < p.slots ? p.b.pI[slot] : make_uint4(0u, 0u, 0u, SLOT_RETIRED);
    uint32_t state = I.w;
    uint32_t n_paths = 0, n_rays = 0, n_shadow = 0;
    bool emit_closest = false, emit_shadow = false, shadow_blocked = false;
    V3 o = v3s(0.f), d = v3s(0.f), sdir = v3s(0.f);

    if (state != SLOT_RETIRED) {
        const uint32_t sample_end = p.sample_begin + p.sample_count;
        V3 T = v3s(1.f), L = v3s(0.f), pending = v3s(0.f);
        float last_pdf = 1.0f;
        uint32_t rng = 0u, flags = 0u, bounce = I.z;
        bool finish = false, fresh = state == SLOT_NEW;
        if (!fresh) {
            const float4 t4 = p.b.pT[slot], l4 = p.b.pL[slot], p4 = p.b.pP[slot];
            T = xyz(t4);
            last_pdf = t4.w;
            L = xyz(l4);
            rng = f_as_u(l4.w);
            pending = xyz(p4);
            flags = f_as_u(p4.w);
            // shader.wgsl:1249-1264: the next-event sample counts iff nothing blocks it.
            if ((flags & kFlagShadow) && p.b.occl[slot] == 0u) L = L + pending;
            flags = 0u;
            if (state == SLOT_ENDING) {
                finish = true;
            } else {
                const float4 o4 = p.b.ray_o[slot], d4 = p.b.ray_d[slot], h4 = p.b.hit[slot];
                o = xyz(o4);
                d = xyz(d4);
                FHit hit;
                hit.t = h4.x;
                hit.rec = f_as_u(h4.y);
                hit.u = h4.z;
                hit.v = h4.w;
                if (hit.rec == kNone) {
                    // shader.wgsl:1222-1231 — escaped into the environment.
                    float u, v;
                    direction_to_uv(d, u, v);
                    V3 sky = env_bilinear(p.env, u, v);
                    float pdf = env_direction_pdf(p.env, u, v);
                    float w = power_heuristic(last_pdf, pdf);
                    L = L + T * sky * w;
                    finish = true;
                } else {
                    ShadeOut so;
                    shade_hit(p.scene, p.env, o, d, hit, rng, T, L, so);
                    bounce++;
                    const bool done = so.done || bounce >= p.max_bounces;
                    o = so.point;
                    d = so.next_dir;
                    last_pdf = so.next_pdf;
                    pending = so.pending;
                    if (so.want_shadow) {
                        emit_shadow = true;
                        sdir = so.shadow_dir;
                        flags |= kFlagShadow;
                        n_shadow++;
                    }
                    if (done) {
                        if (emit_shadow) state = SLOT_ENDING;  // wait for the shadow ray, then finish
                        else finish = true;
                    } else {
                        emit_closest = true;
                    }
                }
            }
        }
        uint32_t px = 0, py = 0;
        if (finish) {
            // shader.wgsl:1367-1371: total += sample, alpha = 1. Only this slot touches the pixel.
            pixel_of_id(p, I.x, px, py);
            float4* a = p.accum + (size_t)py * p.width + px;
            float4 v = *a;
            v.x = v.x + L.x;
            v.y = v.y + L.y;
            v.z = v.z + L.z;
            v.w = 1.0f;
            *a = v;
            I.y++;
            if (I.y >= sample_end) {
                I.x += p.slots;  // this pixel is complete: on to the slot's next one
                I.y = p.sample_begin;
            }
            fresh = true;
        }
        if (fresh) {
            I.x = next_pixel_id(p, I.x, px, py);
            if (I.x == kNone) {
                state = SLOT_RETIRED;
            } else {
                const float sin_half_fov = rt_sin(p.camera.fov_y / 2.0f);
                camera_ray(p.camera, sin_half_fov, px, py, p.width, p.height, I.y, rng, o, d);
                n_paths++;
                T = v3s(1.f);
                L = v3s(0.f);
                last_pdf = 1.0f;
                bounce = 0u;
                flags = 0u;
                pending = v3s(0.f);
                emit_closest = true;
                emit_shadow = false;
                state = SLOT_TRACED;
            }
        } else if (state != SLOT_ENDING) {
            state = SLOT_TRACED;
        }
        if (emit_closest) n_rays++;
        I.z = bounce;
        I.w = state;
        p.b.pI[slot] = I;
        if (state != SLOT_RETIRED) {
            p.b.pT[slot] = make_float4(T.x, T.y, T.z, last_pdf);
            p.b.pL[slot] = make_float4(L.x, L.y, L.z, u_as_f(rng));
            p.b.pP[slot] = make_float4(pending.x, pending.y, pending.z, u_as_f(flags));
        }
    }

    // ---- pre-pass for the rays this wave emits: every lane of the warp together ----
    if (emit_closest) {
        const RayCtx r = ray_setup(o, d);
        float t0_root;
        const bool root_ok = root_test(p.scene, r, t0_root);
        FHit best;
        Fallback fb;
        prepass(p.scene, spre, r, root_ok, best, fb);
        p.b.ray_o[slot] = make_float4(o.x, o.y, o.z, best.t);
        p.b.ray_d[slot] = make_float4(d.x, d.y, d.z, u_as_f(best.rec));
        p.b.fb[slot] = make_float2(fb.t, u_as_f(fb.index));
    } else if (emit_shadow) {
        p.b.ray_o[slot] = make_float4(o.x, o.y, o.z, 0.0f);  // the shadow ray starts at the hit point
    }
    if (emit_shadow) {
        const RayCtx r = ray_setup(o, sdir);
        float t0_root;
        const bool root_ok = root_test(p.scene, r, t0_root);
        FHit best;
        Fallback fb;
        prepass(p.scene, spre, r, root_ok, best, fb);
        shadow_blocked = best.rec != kNone;
        p.b.occl[slot] = shadow_blocked ? 1u : 0u;
        if (!shadow_blocked) p.b.sray_d[slot] = make_float4(sdir.x, sdir.y, sdir.z, 0.0f);
    }
    queue_push(p, ctrl, emit_closest, slot);
    queue_push(p, ctrl, emit_shadow && !shadow_blocked, slot | kShadowBit);

    // ---- bookkeeping ----
    const unsigned live = __ballot_sync(0xFFFFFFFFu, state != SLOT_RETIRED);
    if ((threadIdx.x & 31u) == 0u && live) atomicAdd(&ctrl->live, (uint32_t)__popc(live));
    fast_flush_counts(n_paths, n_rays, n_shadow, p.counters);
}

#ifndef RT_WAVE_TRACE_MIN_BLOCKS
#define RT_WAVE_TRACE_MIN_BLOCKS 4
#endif

// The triangle walk of every queued ray. Persistent warps; a lane whose ray is finished
// takes the next queue entry at the next round boundary.
__global__ void __launch_bounds__(kTraceThreads, RT_WAVE_TRACE_MIN_BLOCKS) wave_trace_kernel(const WaveParams p) {
    WaveCtrl* ctrl = p.b.ctrl + (p.wave & 1u);
    const uint32_t total = ctrl->queued;
    const uint32_t lane = threadIdx.x & 31u;
    uint2 stack[kStackCap];

    bool busy = false, any_hit = false;
    uint32_t slot = 0;
    RayCtx r = ray_setup(v3s(0.f), v3s(1.f));
    FHit best;
    best.t = RT_INFINITY;
    best.rec = kNone;
    best.u = best.v = 0.0f;
    Walk w;
    w.cur = kNone;
    w.sp = 0u;
    w.limit = RT_INFINITY;
    bool drained = false;  // the queue has nothing more to hand out

    for (;;) {
        // ---- refill: idle lanes take the next entries ----
        const unsigned idle = __ballot_sync(0xFFFFFFFFu, !busy);
        if (idle && !drained) {
            const int leader = __ffs(idle) - 1;
            uint32_t base = 0;
            if ((int)lane == leader) base = atomicAdd(&ctrl->head, (uint32_t)__popc(idle));
            base = __shfl_sync(0xFFFFFFFFu, base, leader);
            if (base + (uint32_t)__popc(idle) >= total) drained = true;
            if (!busy) {
                const uint32_t idx = base + __popc(idle & ((1u << lane) - 1u));
                if (idx < total) {
                    const uint32_t e = p.b.queue[idx];
                    slot = e & ~kShadowBit;
                    any_hit = (e & kShadowBit) != 0u;
                    const float4 o4 = p.b.ray_o[slot];
                    const float4 d4 = any_hit ? p.b.sray_d[slot] : p.b.ray_d[slot];
                    r = ray_setup(xyz(o4), xyz(d4));
                    best.t = any_hit ? RT_INFINITY : o4.w;
                    best.rec = any_hit ? kNone : f_as_u(d4.w);
                    best.u = best.v = 0.0f;
                    float t0_root;
                    const bool root_ok = root_test(p.scene, r, t0_root);
                    walk_begin(p.scene, root_ok, t0_root, best, w);
                    busy = true;
                }
            }
        }
        if (!__any_sync(0xFFFFFFFFu, busy)) break;

        // ---- one round of the walk: interior records, then the leaf ----
        while (busy && w.cur < kLeafBit) walk_node(p.scene, r, w, stack);
        bool settled = false;
        if (busy && w.cur != kNone) settled = walk_leaf(p.scene, r, any_hit, best, w, stack);
        if (busy && (settled || w.cur == kNone)) {
            // ---- this lane's ray is finished ----
            if (any_hit) {
                if (best.rec != kNone) p.b.occl[slot] = 1u;
            } else {
                if (best.rec == kNone) {
                    const float2 f = p.b.fb[slot];
                    Fallback fb;
                    fb.t = f.x;
                    fb.index = f_as_u(f.y);
                    resolve_fallback(p.scene, r, fb, best);
                }
                p.b.hit[slot] = make_float4(best.t, u_as_f(best.rec), best.u, best.v);
            }
            busy = false;
        }
    }
}

}  // namespace rtw
