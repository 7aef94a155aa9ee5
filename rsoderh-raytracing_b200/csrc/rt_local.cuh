// rt_local.cuh — the persistent alternative to the launch-per-stage pipeline of rt_wave.cuh
// (RT_KERNEL=local): ONE kernel per rt_render in which every CTA runs its own small wavefront
// pipeline over a private pool of path slots. It is bit-exact with every other kernel, fully
// asynchronous and needs 80 MB instead of gigabytes, but it is NOT the default: it renders
// house.toml at 6.5-6.8 Grays/s against 7.85 for rt_wave.cuh. ncu says why (profiles/README.md,
// round 2): the kernel is 112 KB of SASS, more than an SM's instruction cache holds while its
// two CTAs sit in different phases, so it re-fetches phase code from the GPC-level instruction
// cache all the time (sm__icc hit rate 94.8 %, gcc instruction requests at 71 % of their peak),
// and 28 % of its stall samples wait at the phase barriers. Letting every WARP run the phases on
// its own (experiments/rt_local_warp_engines.cuh: no barriers, full 32-lane batches) makes it
// worse, not better: 32 warps in four code regions drop the hit rate to 54 %, saturate the
// GPC instruction cache (98 % of peak, 66 % of stall samples "no instruction") and halve the
// throughput. One launch per stage keeps each kernel's code (18-43 KB) resident: 99.8 % hits.
//
// The path loop of shader.wgsl:1213-1303 is cut at its ray casts into four phases, as in
// rt_wave.cuh, but the phases are loops inside one kernel, separated by __syncthreads()
// instead of kernel boundaries, and every CTA owns everything it touches:
//
//   shade    slots whose ray hit -> settle the last shadow ray, shade (emission, next-event
//            set-up, BSDF sample), emit the bounce ray + shadow ray into the CTA's ray queue
//   finish   slots whose ray escaped / path ended / that hold no path -> sky light, park the
//            finished sample, take the next path of the CTA's open windows (camera ray) and
//            run that ray's pre-pass on the spot
//   prepass  ray queue -> analytic pre-pass (+ the leaf-box mask of the flat walk); rays with
//            no triangle to test are finished here, the others go to the walk queue
//   walk     walk queue -> triangle tests; persistent lanes that take the next ray the
//            moment their own is finished
//
// What this buys over one kernel launch per phase (rt_wave.cuh):
//   * no launches and no grid-wide barriers between phases: nothing to amortise, so a CTA's
//     pool is ~1 K slots instead of 8 Mi for the whole GPU — 296 pools x 137 B x 1 K slots =
//     41 MB of path state that stays in L2 (and partly in the SM's own L1, which a persistent
//     kernel never flushes) instead of streaming 1.3 GB through DRAM four times per wave;
//   * queues, lists and counters live in shared memory (atomics cost a few cycles);
//   * rt_render is a single asynchronous launch; the pipeline is full from the first
//     microsecond and drains one tile at a time, which is what a 32-spp share of a
//     sample-split frame needs.
//
// Work distribution: the image is cut into 8x8-pixel tiles handed out from a global counter.
// A CTA renders ALL samples of a tile, in windows of W consecutive samples; a window's
// 64 x W paths are the work items its slots draw from. A finished sample is parked in the
// window's staging block, and when the window's last path is in, 64 threads add its planes
// to the accumulation buffer in sample order — the reference's frame-by-frame accumulation
// (shader.wgsl:1367-1371). Windows are committed in the order they were opened, so every bit
// of the image is independent of scheduling. Several windows (of one or more tiles) are open
// at a time, so the slots never wait for a straggler path.
//
// The walk phase has two forms: for a scene whose effective tree has at most 32 leaves
// (FastScene::flat — house.toml, default.toml) the pre-pass tests every leaf box in
// straight-line code from shared memory and hands the walk a bit mask of the leaves to test,
// in the ray's own visiting order; the walk then tests ONE triangle per lane and iteration,
// with no stack and no node phase. Larger trees use the N2 stack walk of rt_fast.cuh.
#pragma once

#include "rt_fast.cuh"

namespace rtl {
using namespace rtd;
using namespace rtf;

#ifndef RT_LOCAL_THREADS
#define RT_LOCAL_THREADS 512
#endif
#ifndef RT_LOCAL_MIN_BLOCKS
#define RT_LOCAL_MIN_BLOCKS 2
#endif
constexpr int kThreads = RT_LOCAL_THREADS;
constexpr int kWarps = kThreads / 32;
constexpr int kMinBlocks = RT_LOCAL_MIN_BLOCKS;
constexpr uint32_t kRing = 32;         // windows in flight per CTA, at most
constexpr uint32_t kMaxCommit = 4;     // windows committed per maintenance step
constexpr uint32_t kFresh = 0x8000u;   // finish-list entry: the slot holds no path
constexpr uint32_t kShadow = 0x8000u;  // ray / walk queue entry: the slot's shadow ray
constexpr uint32_t kSlotMask = 0x7FFFu;
constexpr uint32_t kMaxSlots = 0x7FFFu;

// low byte of the path record's last word; the other bytes hold the ring entry and the
// pixel's position inside its tile
enum PathState : uint32_t {
    PATH_TRACED = 2,  // a closest-hit ray is in flight / has come back
    PATH_DONE = 4     // the path is over
};
constexpr uint32_t kFlagShadow = 1u;  // a shadow ray was cast for the last shaded hit

struct RingEntry {
    uint32_t tx, ty;     // tile (8x8 pixels)
    uint32_t w0, count;  // first sample of the window, samples in it
    uint32_t left;       // paths of the window that have not parked their sample yet
    uint32_t pad[3];
};

struct LocalCtrl {
    uint32_t n_shade[2], n_fin[2];   // list lengths, by wave parity
    uint32_t n_ray, n_walk, walk_head;
    uint32_t handed, limit;          // path handout: virtual indices < limit belong to open windows
    uint32_t ring_head, ring_tail;   // windows [head, tail) are open or waiting for their commit
    uint32_t tile_next, tile_end;    // tiles taken from the global counter, not yet opened
    uint32_t tile_valid, cur_tx, cur_ty, next_window;
    uint32_t exhausted;              // the global counter has run out
    uint32_t commit_n;
    uint32_t quit;
    // profiling (LocalParams::profile): cycles per phase, items per phase, waves
    unsigned long long t_last, cycles[5], items[4], waves;
    RingEntry commit[kMaxCommit];
    RingEntry ring[kRing];
};

struct LocalParams {
    FastScene scene;
    EnvView env;
    CameraParams camera;
    float4* accum;
    unsigned long long* counters;
    uint32_t width, height;
    uint32_t tiles_x, tiles_total;
    uint32_t sample_begin, sample_count;
    uint32_t max_bounces;
    uint32_t slots;          // path slots per CTA (<= kMaxSlots)
    uint32_t window_shift;   // W = 1 << window_shift samples per window
    uint32_t ring_entries;   // power of two <= kRing
    uint32_t tile_grab;      // tiles taken per atomic on the global counter
    // per-CTA pools, indexed by blockIdx.x * slots + slot
    float4* rOD;      // [2] ray origin (.w = closest analytic distance), direction (.w = bits rec of it)
    float4* pTLPI;    // [4] throughput + last_pdf, radiance + bits rng, pending light + bits flags,
                      //     bits (pixel, sample, bounce, state | ring entry << 8 | pixel in tile << 16)
    float4* sray_d;   // direction of the shadow ray
    float4* hit;      // (t, bits rec, u, v)
    float2* fb;       // brute-force fallback winner of the closest-hit ray
    uint8_t* occl;    // 1 when the shadow ray is blocked
    float4* staging;  // [CTA][ring_entries][W][64] parked samples
    uint32_t* next_tile;
    // optional: [0..4] cycles in shade, window bookkeeping + commits, finish, pre-pass, walk summed
    // over CTAs, [5..8] items those phases worked on (shade, finish, pre-pass, walk), [9] waves,
    // [10] CTAs
    unsigned long long* profile;
};

// Shared-memory layout of a CTA: pre-pass program | flat program + triangle records | control
// block | walk masks | shade lists, finish lists, ray queue, walk queue. Sized by the scene.
struct SmemLayout {
    uint32_t pre_f4, flat_f4, tri_f4;  // float4 counts
    size_t bytes;
};
inline __host__ __device__ SmemLayout smem_layout(const FastScene& sc, uint32_t slots, bool flat) {
    SmemLayout l;
    l.pre_f4 = sc.hoisted ? sc.pre_leaves + 2u * sc.pre_boxes + 4u * sc.pre_prims : 0u;
    l.flat_f4 = flat ? kFlatFloat4 : 0u;
    l.tri_f4 = flat ? kTriStride * sc.flat_slots : 0u;
    l.bytes = (size_t)(l.pre_f4 + l.flat_f4 + l.tri_f4) * 16u + ((sizeof(LocalCtrl) + 15u) & ~(size_t)15u) +
              (flat ? (size_t)2u * slots * 4u : 0u) + (size_t)8u * slots * 2u;
    return l;
}

struct PathRegs {
    V3 T, L, pending;
    float last_pdf;
    uint32_t rng, flags;
    uint32_t occluded;
};

struct Cta {
    LocalCtrl* c;
    const float4* spre;
    const float4* sflat;
    const float4* stri;
    uint16_t* shadeL;  // [2][S]
    uint16_t* finL;    // [2][S]
    uint16_t* rayq;    // [2S]
    uint16_t* walkq;   // [2S]
    uint32_t* walkm;   // [2S] leaf masks of the walk queue's rays (flat walk)
    size_t pool;       // first global slot of this CTA
    float4* stage;     // this CTA's staging blocks
    uint32_t lane, warp;
};

RT_DEVONLY void load_path(const LocalParams& p, size_t g, PathRegs& s, uint4& I) {
    float4 t4, l4, p4, i4;
    ld8w<0>(&p.pTLPI[4u * g + 0u], t4, l4);
    ld8w<0>(&p.pTLPI[4u * g + 2u], p4, i4);
    s.T = xyz(t4);
    s.last_pdf = t4.w;
    s.L = xyz(l4);
    s.rng = f_as_u(l4.w);
    s.pending = xyz(p4);
    s.flags = f_as_u(p4.w);
    s.occluded = p.occl[g];
    I = make_uint4(f_as_u(i4.x), f_as_u(i4.y), f_as_u(i4.z), f_as_u(i4.w));
}

RT_DEVONLY void store_path(const LocalParams& p, size_t g, const PathRegs& s, const uint4 I) {
    st8<0>(&p.pTLPI[4u * g + 0u], make_float4(s.T.x, s.T.y, s.T.z, s.last_pdf),
           make_float4(s.L.x, s.L.y, s.L.z, u_as_f(s.rng)));
    st8<0>(&p.pTLPI[4u * g + 2u], make_float4(s.pending.x, s.pending.y, s.pending.z, u_as_f(s.flags)),
           make_float4(u_as_f(I.x), u_as_f(I.y), u_as_f(I.z), u_as_f(I.w)));
}

// shader.wgsl:1249-1264: the last next-event sample counts iff nothing blocked its ray.
RT_DEVONLY void settle_shadow(PathRegs& s) {
    if ((s.flags & kFlagShadow) && s.occluded == 0u) s.L = s.L + s.pending;
    s.flags = 0u;
}

// Appends one entry per requesting lane to a shared-memory list with one atomic per warp.
// Every lane of the warp must call it.
RT_DEVONLY uint32_t reserve(uint32_t* count, bool want, uint32_t lane) {
    const unsigned m = __ballot_sync(0xFFFFFFFFu, want);
    if (!m) return 0u;
    const int leader = __ffs(m) - 1;
    uint32_t base = 0;
    if ((int)lane == leader) base = atomicAdd(count, (uint32_t)__popc(m));
    base = __shfl_sync(0xFFFFFFFFu, base, leader);
    return base + __popc(m & ((1u << lane) - 1u));
}
RT_DEVONLY void push(uint16_t* list, uint32_t* count, bool want, uint32_t entry, uint32_t lane) {
    const uint32_t at = reserve(count, want, lane);
    if (want) list[at] = (uint16_t)entry;
}

// What the pre-pass leaves for the walk.
template <bool FLAT>
struct PreOut {
    FHit best;
    Fallback fb;
    uint32_t mask;  // FLAT: leaves to test, in visiting order
    bool walk;      // the ray has triangles to test
};

// The analytic pre-pass of one ray and the decision whether it needs the walk.
template <bool FLAT>
RT_DEVONLY void pre_ray(const LocalParams& p, const Cta& k, const RayCtx& r, bool shadow, PreOut<FLAT>& out) {
    prepass(p.scene, k.spre, r, out.best, out.fb);
    if (FLAT) {
        const float limit = fminf(RT_INFINITY, out.best.t * 1.001f);
        out.mask = flat_mask(p.scene, k.sflat, r, limit);
        out.walk = out.mask != 0u;
    } else {
        out.mask = 0u;
        float t0_root;
        const bool root_ok = root_test(p.scene, r, t0_root);
        Walk w;
        walk_begin(p.scene, root_ok, t0_root, out.best, w);
        out.walk = w.cur != kNone;
    }
    if (shadow && out.best.rec != kNone) out.walk = false;  // blocked by an analytic primitive
}

// ---- phase: shade ----
template <bool FLAT>
RT_DEVONLY void phase_shade(const LocalParams& p, const Cta& k, uint32_t P, uint32_t& n_rays, uint32_t& n_shadow) {
    const uint32_t Q = P ^ 1u, S = p.slots;
    LocalCtrl* c = k.c;
    const uint32_t n = c->n_shade[P];
    for (uint32_t base = k.warp * 32u; base < n; base += (uint32_t)kWarps * 32u) {
        const uint32_t idx = base + k.lane;
        const bool on = idx < n;
        const uint32_t slot = on ? k.shadeL[P * S + idx] : 0u;
        const size_t g = k.pool + slot;
        bool emit_closest = false, emit_shadow = false, fin_now = false, fin_next = false;
        if (on) {
            PathRegs s;
            uint4 I;
            load_path(p, g, s, I);
            settle_shadow(s);
            float4 o4, d4;
            ld8w<0>(&p.rOD[2u * g], o4, d4);
            const float4 h4 = p.hit[g];
            FHit hit;
            hit.t = h4.x;
            hit.rec = f_as_u(h4.y);
            hit.u = h4.z;
            hit.v = h4.w;
            ShadeOut so;
            shade_hit(p.scene, p.env, xyz(o4), xyz(d4), hit, s.rng, s.T, s.L, so);
            I.z++;
            const bool done = so.done || I.z >= p.max_bounces;
            s.last_pdf = so.next_pdf;
            s.pending = so.pending;
            if (so.want_shadow) {
                emit_shadow = true;
                s.flags = kFlagShadow;
                p.sray_d[g] = make_float4(so.shadow_dir.x, so.shadow_dir.y, so.shadow_dir.z, 0.0f);
                p.occl[g] = 0;
                n_shadow++;
            }
            // Both rays start at the hit point (no offset: shader.wgsl:1249, 1295).
            const float4 point4 = make_float4(so.point.x, so.point.y, so.point.z, RT_INFINITY);
            if (done) {
                // With a shadow ray in flight the sample is finished one wave later.
                I.w = (I.w & ~0xFFu) | PATH_DONE;
                fin_next = emit_shadow;
                fin_now = !emit_shadow;
                if (emit_shadow) p.rOD[2u * g + 0u] = point4;
            } else {
                emit_closest = true;
                st8<0>(&p.rOD[2u * g], point4, make_float4(so.next_dir.x, so.next_dir.y, so.next_dir.z, u_as_f(kNone)));
                n_rays++;
            }
            store_path(p, g, s, I);
        }
        push(k.rayq, &c->n_ray, emit_closest, slot, k.lane);
        push(k.rayq, &c->n_ray, emit_shadow, slot | kShadow, k.lane);
        push(k.finL + P * S, &c->n_fin[P], fin_now, slot, k.lane);
        push(k.finL + Q * S, &c->n_fin[Q], fin_next, slot, k.lane);
    }
}

// Where a ray that needs no walk goes next, and its result.
RT_DEVONLY void settle_closest(const LocalParams& p, const RayCtx& r, size_t g, const Fallback& fb, FHit best,
                               bool& to_shade, bool& to_fin) {
    resolve_fallback(p.scene, r, fb, best);
    p.hit[g] = make_float4(best.t, u_as_f(best.rec), best.u, best.v);
    to_shade = best.rec != kNone;
    to_fin = !to_shade;
}

// ---- phase: finish ----
template <bool FLAT>
RT_DEVONLY void phase_finish(const LocalParams& p, const Cta& k, uint32_t P, float sin_half_fov, uint32_t& n_paths,
                             uint32_t& n_rays) {
    const uint32_t Q = P ^ 1u, S = p.slots;
    LocalCtrl* c = k.c;
    const uint32_t n = c->n_fin[P];
    const uint32_t limit = c->limit;
    const bool exhausted = c->exhausted != 0u;
    const uint32_t eshift = 6u + p.window_shift;
    const uint32_t emask = (1u << eshift) - 1u;
    for (uint32_t base = k.warp * 32u; base < n; base += (uint32_t)kWarps * 32u) {
        const uint32_t idx = base + k.lane;
        const bool on = idx < n;
        const uint32_t entry = on ? k.finL[P * S + idx] : kFresh;
        const uint32_t slot = entry & kSlotMask;
        const bool had_path = on && !(entry & kFresh);
        const size_t g = k.pool + slot;
        PathRegs s;
        s.rng = 0u;
        uint4 I = make_uint4(0u, 0u, 0u, 0u);
        float4 d4 = make_float4(0.f, 0.f, 0.f, 0.f);
        if (had_path) {
            load_path(p, g, s, I);
            d4 = p.rOD[2u * g + 1u];
        }
        // ---- the next path indices: one shared-memory atomic per warp ----
        uint32_t h = reserve(&c->handed, on, k.lane);
        if (had_path) {
            settle_shadow(s);
            if ((I.w & 0xFFu) == PATH_TRACED) {
                // shader.wgsl:1222-1231 — the ray escaped into the environment.
                const V3 d = xyz(d4);
                float u, v;
                direction_to_uv(d, u, v);
                V3 sky = env_bilinear(p.env, u, v);
                float pdf = env_direction_pdf(p.env, u, v);
                float w = power_heuristic(s.last_pdf, pdf);
                s.L = s.L + s.T * sky * w;
            }
            // Park the sample in its window's staging block.
            const uint32_t rs = (I.w >> 8) & 0xFFu, pix = I.w >> 16;
            const uint32_t kk = I.y - c->ring[rs].w0;
            __stcs(&k.stage[(((size_t)rs << p.window_shift) + kk) * 64u + pix], make_float4(s.L.x, s.L.y, s.L.z, 0.0f));
            atomicSub(&c->ring[rs].left, 1u);
        }
        bool to_walk = false, to_shade = false, to_fin = false, got = false;
        uint32_t mask = 0u;
        if (on) {
            // Draw until the index names a pixel-sample (indices of partial tiles / of a short
            // last window name none).
            uint32_t px = 0, py = 0, sample = 0, rs = 0, pix = 0;
            while ((int32_t)(h - limit) < 0) {
                const uint32_t i = h & emask;
                rs = (h >> eshift) & (p.ring_entries - 1u);
                const RingEntry& en = c->ring[rs];
                const uint32_t kk = i >> 6;
                pix = i & 63u;
                px = en.tx * 8u + (pix & 7u);
                py = en.ty * 8u + (pix >> 3);
                if (kk < en.count && px < p.width && py < p.height) {
                    sample = en.w0 + kk;
                    got = true;
                    break;
                }
                h = atomicAdd(&c->handed, 1u);
            }
            if (!got) {
                to_fin = !exhausted;  // try again next wave (as a slot without a path); else retire
            } else {
                I.x = py * p.width + px;
                I.y = sample;
                I.z = 0u;
                I.w = PATH_TRACED | (rs << 8) | (pix << 16);
                V3 o, d;
                camera_ray(p.camera, sin_half_fov, px, py, p.width, p.height, sample, s.rng, o, d);
                s.T = v3s(1.f);
                s.L = v3s(0.f);
                s.pending = v3s(0.f);
                s.last_pdf = 1.0f;
                s.flags = 0u;
                s.occluded = 0u;
                const RayCtx r = ray_setup(o, d);
                PreOut<FLAT> pre;
                pre_ray<FLAT>(p, k, r, false, pre);
                to_walk = pre.walk;
                mask = pre.mask;
                st8<0>(&p.rOD[2u * g], make_float4(o.x, o.y, o.z, to_walk ? pre.best.t : RT_INFINITY),
                       make_float4(d.x, d.y, d.z, u_as_f(to_walk ? pre.best.rec : kNone)));
                if (to_walk) p.fb[g] = make_float2(pre.fb.t, u_as_f(pre.fb.index));
                else settle_closest(p, r, g, pre.fb, pre.best, to_shade, to_fin);
                store_path(p, g, s, I);
                n_paths++;
                n_rays++;
            }
        }
        {
            const uint32_t at = reserve(&c->n_walk, to_walk, k.lane);
            if (to_walk) {
                k.walkq[at] = (uint16_t)slot;
                if (FLAT) k.walkm[at] = mask;
            }
        }
        push(k.shadeL + Q * S, &c->n_shade[Q], to_shade, slot, k.lane);
        // a slot that drew nothing re-enters the finish list without a path
        push(k.finL + Q * S, &c->n_fin[Q], to_fin, got ? slot : (slot | kFresh), k.lane);
    }
}

// ---- phase: pre-pass of the rays the shade phase emitted ----
template <bool FLAT>
RT_DEVONLY void phase_prepass(const LocalParams& p, const Cta& k, uint32_t P) {
    const uint32_t Q = P ^ 1u, S = p.slots;
    LocalCtrl* c = k.c;
    const uint32_t n = c->n_ray;
    for (uint32_t base = k.warp * 32u; base < n; base += (uint32_t)kWarps * 32u) {
        const uint32_t idx = base + k.lane;
        const bool on = idx < n;
        const uint32_t e = on ? k.rayq[idx] : 0u;
        const uint32_t slot = e & kSlotMask;
        const bool shadow = (e & kShadow) != 0u;
        const size_t g = k.pool + slot;
        bool to_walk = false, to_shade = false, to_fin = false;
        uint32_t mask = 0u;
        if (on) {
            const V3 o = xyz(p.rOD[2u * g + 0u]);
            const V3 d = xyz(shadow ? p.sray_d[g] : p.rOD[2u * g + 1u]);
            const RayCtx r = ray_setup(o, d);
            PreOut<FLAT> pre;
            pre_ray<FLAT>(p, k, r, shadow, pre);
            to_walk = pre.walk;
            mask = pre.mask;
            if (shadow) {
                if (pre.best.rec != kNone) p.occl[g] = 1;
            } else if (to_walk) {
                p.rOD[2u * g + 0u].w = pre.best.t;
                p.rOD[2u * g + 1u].w = u_as_f(pre.best.rec);
                p.fb[g] = make_float2(pre.fb.t, u_as_f(pre.fb.index));
            } else {
                settle_closest(p, r, g, pre.fb, pre.best, to_shade, to_fin);
            }
        }
        {
            const uint32_t at = reserve(&c->n_walk, to_walk, k.lane);
            if (to_walk) {
                k.walkq[at] = (uint16_t)e;
                if (FLAT) k.walkm[at] = mask;
            }
        }
        push(k.shadeL + Q * S, &c->n_shade[Q], to_shade, slot, k.lane);
        push(k.finL + Q * S, &c->n_fin[Q], to_fin, slot, k.lane);
    }
}

// A walked ray's result: closest-hit rays move to the shade or finish list of the next wave.
RT_DEVONLY void walk_result(const LocalParams& p, const RayCtx& r, size_t g, bool any_hit, FHit best, bool& to_shade,
                            bool& to_fin) {
    if (any_hit) {
        if (best.rec != kNone) p.occl[g] = 1;
        return;
    }
    if (best.rec == kNone) {
        const float2 f = p.fb[g];
        Fallback fb;
        fb.t = f.x;
        fb.index = f_as_u(f.y);
        resolve_fallback(p.scene, r, fb, best);
    }
    p.hit[g] = make_float4(best.t, u_as_f(best.rec), best.u, best.v);
    to_shade = best.rec != kNone;
    to_fin = !to_shade;
}

// ---- phase: walk. Persistent lanes; a lane whose ray is finished takes the next entry. ----
template <bool FLAT>
RT_DEVONLY void phase_walk(const LocalParams& p, const Cta& k, uint32_t P) {
    const uint32_t Q = P ^ 1u, S = p.slots;
    LocalCtrl* c = k.c;
    const uint32_t total = c->n_walk;
    const uint32_t lane = k.lane;

    bool busy = false, any_hit = false, drained = false;
    uint32_t slot = 0;
    RayCtx r = ray_setup(v3s(0.f), v3s(1.f));
    FHit best;
    best.t = RT_INFINITY;
    best.rec = kNone;
    best.u = best.v = 0.0f;
    // FLAT: leaves still to visit, triangle slots [cur, end) of the current leaf
    uint32_t mask = 0u, cur = 0u, end = 0u;
    float limit = RT_INFINITY;
    // !FLAT: the N2 stack walk
    Walk w;
    w.cur = kNone;
    w.sp = 0u;
    w.limit = RT_INFINITY;
    uint2 stack[FLAT ? 1 : kStackCap];

    for (;;) {
        // ---- refill: idle lanes take the next entries ----
        const unsigned idle = __ballot_sync(0xFFFFFFFFu, !busy);
        if (idle && !drained) {
            const int leader = __ffs(idle) - 1;
            uint32_t base = 0;
            if ((int)lane == leader) base = atomicAdd(&c->walk_head, (uint32_t)__popc(idle));
            base = __shfl_sync(0xFFFFFFFFu, base, leader);
            if (base + (uint32_t)__popc(idle) >= total) drained = true;
            if (!busy) {
                const uint32_t idx = base + __popc(idle & ((1u << lane) - 1u));
                if (idx < total) {
                    const uint32_t e = k.walkq[idx];
                    slot = e & kSlotMask;
                    any_hit = (e & kShadow) != 0u;
                    const size_t g = k.pool + slot;
                    const float4 o4 = p.rOD[2u * g + 0u];
                    const float4 d4 = any_hit ? p.sray_d[g] : p.rOD[2u * g + 1u];
                    r = ray_setup(xyz(o4), xyz(d4));
                    best.t = any_hit ? RT_INFINITY : o4.w;
                    best.rec = any_hit ? kNone : f_as_u(d4.w);
                    best.u = best.v = 0.0f;
                    if (FLAT) {
                        mask = k.walkm[idx];
                        cur = end = 0u;
                        limit = fminf(RT_INFINITY, best.t * 1.001f);
                    } else {
                        walk_begin(p.scene, true, 0.0f, best, w);  // the pre-pass checked the root box
                    }
                    busy = true;
                }
            }
        }
        if (!__any_sync(0xFFFFFFFFu, busy)) break;  // nothing was handed out: the queue is empty

        bool settled = false, finished = false;
        if (FLAT) {
            // the next leaf whose box is still within the (shrinking) limit
            const uint8_t* order = flat_order(k.sflat, r.neg);
            while (busy && cur == end && mask) {
                const uint32_t leaf = order[lowest_bit(mask)];
                mask &= mask - 1u;
                const float4 bmin = k.sflat[2u * leaf], bmax = k.sflat[2u * leaf + 1u];
                float t0;
                if (box_test(r, bmin, bmax, limit, t0)) {
                    cur = f_as_u(bmin.w);
                    end = cur + f_as_u(bmax.w);
                }
            }
            // one triangle
            if (busy && cur < end) {
                const float4* q = k.stri + kTriStride * cur;
                const float4 r0 = q[0], r1 = q[1], r2 = q[2];
                if (f_as_i(r0.w) == 2) {
                    float t = 0.0f, u = 0.0f, v = 0.0f;
                    if (test_triangle(r.o, r.d, r0, r1, r2, t, u, v))
                        offer(p.scene, r, best, t, cur, cur, u, v, best.rec < kRecAnalytic);
                    settled = any_hit && best.rec != kNone;
                }
                cur++;
                if (cur == end) limit = fminf(RT_INFINITY, best.t * 1.001f);
            }
            finished = busy && (settled || (cur == end && mask == 0u));
        } else {
            if (__all_sync(0xFFFFFFFFu, !busy || r.plain)) {
                while (busy && w.cur < kLeafBit) walk_node<true>(p.scene, r, w, stack);
            } else {
                while (busy && w.cur < kLeafBit) walk_node<false>(p.scene, r, w, stack);
            }
            if (busy && w.cur != kNone) settled = walk_leaf(p.scene, r, any_hit, best, w, stack);
            finished = busy && (settled || w.cur == kNone);
        }

        // ---- lanes whose ray is finished hand over their result ----
        bool to_shade = false, to_fin = false;
        if (finished) {
            walk_result(p, r, k.pool + slot, any_hit, best, to_shade, to_fin);
            busy = false;
        }
        push(k.shadeL + Q * S, &c->n_shade[Q], to_shade, slot, lane);
        push(k.finL + Q * S, &c->n_fin[Q], to_fin, slot, lane);
    }
}

// ---- window bookkeeping, by thread 0 between the shade and the finish phase ----
RT_DEVONLY void maintain(const LocalParams& p, LocalCtrl* c) {
    const uint32_t E = 64u << p.window_shift;
    // the handout may have run past the open windows; those indices went to nobody
    if ((int32_t)(c->handed - c->limit) > 0) c->handed = c->limit;
    // windows whose paths are all in, oldest first
    c->commit_n = 0u;
    while (c->ring_head != c->ring_tail && c->commit_n < kMaxCommit) {
        const RingEntry& e = c->ring[c->ring_head & (p.ring_entries - 1u)];
        if (e.left != 0u) break;
        c->commit[c->commit_n] = e;
        c->commit[c->commit_n].pad[0] = c->ring_head & (p.ring_entries - 1u);
        c->commit_n++;
        c->ring_head++;
    }
    // open windows until every slot could draw a path in the coming finish phase
    const uint32_t W = 1u << p.window_shift;
    while (!c->exhausted && c->ring_tail - c->ring_head < p.ring_entries && c->limit - c->handed < p.slots) {
        if (!c->tile_valid) {
            if (c->tile_next == c->tile_end) {
                const uint32_t t = atomicAdd(p.next_tile, p.tile_grab);
                if (t >= p.tiles_total) {
                    c->exhausted = 1u;
                    break;
                }
                c->tile_next = t;
                c->tile_end = t + p.tile_grab < p.tiles_total ? t + p.tile_grab : p.tiles_total;
            }
            const uint32_t t = c->tile_next++;
            c->cur_tx = t % p.tiles_x;
            c->cur_ty = t / p.tiles_x;
            c->next_window = 0u;
            c->tile_valid = 1u;
        }
        RingEntry& e = c->ring[c->ring_tail & (p.ring_entries - 1u)];
        e.tx = c->cur_tx;
        e.ty = c->cur_ty;
        const uint32_t first = c->next_window * W;
        e.w0 = p.sample_begin + first;
        e.count = p.sample_count - first < W ? p.sample_count - first : W;
        const uint32_t vw = p.width - e.tx * 8u < 8u ? p.width - e.tx * 8u : 8u;
        const uint32_t vh = p.height - e.ty * 8u < 8u ? p.height - e.ty * 8u : 8u;
        e.left = e.count * vw * vh;
        c->next_window++;
        if (c->next_window * W >= p.sample_count) c->tile_valid = 0u;
        c->ring_tail++;
        c->limit += E;
    }
}

// Adds the staged planes of the windows in c->commit to the accumulation buffer, in sample
// order (shader.wgsl:1367-1371: total += sample, alpha = 1). Thread t < 64 owns pixel t of every
// tile, so two windows of one tile are added in the order they were opened.
RT_DEVONLY void commit_windows(const LocalParams& p, const Cta& k) {
    const uint32_t pix = threadIdx.x;
    if (pix >= 64u) return;
    const uint32_t n = k.c->commit_n;
    for (uint32_t j = 0; j < n; j++) {
        const RingEntry& e = k.c->commit[j];
        const uint32_t px = e.tx * 8u + (pix & 7u), py = e.ty * 8u + (pix >> 3);
        if (px >= p.width || py >= p.height) continue;
        float4* at = &p.accum[(size_t)py * p.width + px];
        float4 a = *at;
        const float4* plane = k.stage + (((size_t)e.pad[0] << p.window_shift) * 64u + pix);
        for (uint32_t s = 0; s < e.count; s++) {
            const float4 v = __ldcs(plane + (size_t)s * 64u);
            a.x = a.x + v.x;
            a.y = a.y + v.y;
            a.z = a.z + v.z;
        }
        a.w = 1.0f;
        *at = a;
    }
}

template <bool FLAT>
__global__ void __launch_bounds__(kThreads, kMinBlocks) local_wave_kernel(const LocalParams p) {
    extern __shared__ float4 smem[];
    Cta k;
    const SmemLayout lay = smem_layout(p.scene, p.slots, FLAT);
    {
        float4* q = smem;
        k.spre = q;
        q += lay.pre_f4;
        k.sflat = q;
        q += lay.flat_f4;
        k.stri = q;
        q += lay.tri_f4;
        k.c = (LocalCtrl*)q;
        q += (sizeof(LocalCtrl) + 15u) / 16u;
        uint32_t* u = (uint32_t*)q;
        k.walkm = u;
        if (FLAT) u += 2u * p.slots;
        uint16_t* h = (uint16_t*)u;
        k.shadeL = h;
        h += 2u * p.slots;
        k.finL = h;
        h += 2u * p.slots;
        k.rayq = h;
        h += 2u * p.slots;
        k.walkq = h;
    }
    k.pool = (size_t)blockIdx.x * p.slots;
    k.stage = p.staging + (size_t)blockIdx.x * (((size_t)p.ring_entries << p.window_shift) * 64u);
    k.lane = threadIdx.x & 31u;
    k.warp = threadIdx.x >> 5;
    const uint32_t S = p.slots;

    // ---- set-up: scene programs into shared memory, every slot starts without a path ----
    {
        float4* spre = smem;
        for (uint32_t i = threadIdx.x; i < lay.pre_f4; i += kThreads) spre[i] = p.scene.pre[i];
        if (FLAT) {
            float4* sflat = spre + lay.pre_f4;
            for (uint32_t i = threadIdx.x; i < lay.flat_f4; i += kThreads) sflat[i] = p.scene.flat[i];
            float4* stri = sflat + lay.flat_f4;
            for (uint32_t i = threadIdx.x; i < lay.tri_f4; i += kThreads) stri[i] = p.scene.tri[i];
        }
        for (uint32_t i = threadIdx.x; i < S; i += kThreads) k.finL[i] = (uint16_t)(i | kFresh);
        if (threadIdx.x == 0) {
            LocalCtrl* c = k.c;
            c->n_shade[0] = c->n_shade[1] = 0u;
            c->n_fin[0] = S;
            c->n_fin[1] = 0u;
            c->n_ray = c->n_walk = c->walk_head = 0u;
            c->handed = c->limit = 0u;
            c->ring_head = c->ring_tail = 0u;
            c->tile_next = c->tile_end = 0u;
            c->tile_valid = 0u;
            c->cur_tx = c->cur_ty = c->next_window = 0u;
            c->exhausted = 0u;
            c->commit_n = 0u;
            c->quit = 0u;
        }
    }
    __syncthreads();

    const float sin_half_fov = rt_sin(p.camera.fov_y / 2.0f);
    uint32_t n_paths = 0, n_rays = 0, n_shadow = 0;
    LocalCtrl* c = k.c;

    const bool prof = p.profile != nullptr;
    auto lap = [&](int phase) {  // thread 0, right after a barrier
        const unsigned long long now = clock64();
        c->cycles[phase] += now - c->t_last;
        c->t_last = now;
    };
    if (prof && threadIdx.x == 0) {
        for (int i = 0; i < 5; i++) c->cycles[i] = 0ull;
        for (int i = 0; i < 4; i++) c->items[i] = 0ull;
        c->waves = 0ull;
        c->t_last = clock64();
    }
    for (uint32_t P = 0;; P ^= 1u) {
        phase_shade<FLAT>(p, k, P, n_rays, n_shadow);
        __syncthreads();
        if (threadIdx.x == 0) {
            if (prof) lap(0);
            maintain(p, c);
        }
        __syncthreads();
        commit_windows(p, k);  // reads staging blocks the coming finish phase may reuse
        __syncthreads();
        if (prof && threadIdx.x == 0) lap(1);
        phase_finish<FLAT>(p, k, P, sin_half_fov, n_paths, n_rays);
        __syncthreads();
        if (prof && threadIdx.x == 0) lap(2);
        phase_prepass<FLAT>(p, k, P);
        __syncthreads();
        if (prof && threadIdx.x == 0) lap(3);
        phase_walk<FLAT>(p, k, P);
        __syncthreads();
        if (threadIdx.x == 0) {
            if (prof) {
                lap(4);
                c->items[0] += c->n_shade[P];
                c->items[1] += c->n_fin[P];
                c->items[2] += c->n_ray;
                c->items[3] += c->n_walk;
                c->waves++;
            }
            c->n_shade[P] = 0u;
            c->n_fin[P] = 0u;
            c->n_ray = c->n_walk = c->walk_head = 0u;
            // every slot with a path (or waiting for one) is on a list of the next wave
            c->quit = (c->n_shade[P ^ 1u] + c->n_fin[P ^ 1u]) == 0u ? 1u : 0u;
        }
        __syncthreads();
        if (c->quit) break;
    }
    // the windows still waiting for their commit (all their paths are in)
    for (;;) {
        if (threadIdx.x == 0) maintain(p, c);
        __syncthreads();
        if (c->commit_n == 0u) break;
        commit_windows(p, k);
        __syncthreads();
    }
    fast_flush_counts(n_paths, n_rays, n_shadow, p.counters);
    if (prof && threadIdx.x == 0) {
        for (int i = 0; i < 5; i++) atomicAdd(&p.profile[i], c->cycles[i]);
        for (int i = 0; i < 4; i++) atomicAdd(&p.profile[5 + i], c->items[i]);
        atomicAdd(&p.profile[9], c->waves);
        atomicAdd(&p.profile[10], 1ull);
    }
}

}  // namespace rtl
