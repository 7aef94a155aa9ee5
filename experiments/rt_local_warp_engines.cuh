// rt_local.cuh — the persistent render kernel: ONE launch per rt_render in which every WARP is
// an independent little wavefront engine over its own pool of path slots.
//
// The path loop of shader.wgsl:1213-1303 is cut at its ray casts into four stages, as in
// rt_wave.cuh, but a stage is a batch of (up to) 32 queued items that one warp works on with
// all its lanes, the queues between the stages live in the warp's own shared memory, and
// nothing ever waits for another warp: no kernel boundaries, no __syncthreads, no atomics on
// shared counters (a queue's head and tail are warp-uniform values).
//
//   shade    slots whose rays are all back and whose closest ray hit -> settle the last shadow
//            ray, shade (emission, next-event set-up, BSDF sample), emit the bounce ray and the
//            shadow ray into the ray queue
//   finish   slots whose ray escaped / whose path ended, and idle slots -> sky light, park the
//            finished sample in its window's staging block, take the next path of the warp's
//            open windows (camera ray) and run that ray's pre-pass on the spot
//   prepass  ray queue -> analytic pre-pass; rays with no triangle to test are settled here,
//            the others go to the walk queue
//   walk     walk queue -> triangle tests with persistent lanes: a lane takes the next ray
//            the moment its own is finished. When the queue has run dry and only a few lanes
//            are still walking, they are SUSPENDED (8 words per lane in shared memory; the
//            traversal stack stays where it is) and resumed by the next walk batch, so the
//            tail of a batch is not walked with a nearly empty warp.
//
// The scheduler of a warp runs whichever stage has a full batch waiting (walk, pre-pass,
// shade, finish in that order), else the fullest one. A slot's rays are counted in a per-slot
// word (`pend`); the lane that settles the last of them files the slot under shade or finish.
//
// What this buys over one kernel launch per stage (rt_wave.cuh):
//   * no launches, no grid-wide or CTA-wide barriers: nothing to amortise, so the GPU holds
//     ~600 K path slots instead of 8 Mi, and their state (137 B per slot) stays in L2 instead
//     of streaming through DRAM four times per wave;
//   * stage batches are full (32 of 32 lanes) whenever 32 items are queued, whatever the mix of
//     path lengths;
//   * rt_render is a single asynchronous launch; the pipeline is full from the first
//     microsecond and drains one tile at a time, which is what a 32-spp share of a
//     sample-split frame needs.
//
// Work distribution: the image is cut into small tiles (8x4 or 4x4 pixels) handed out from a
// global counter. A warp renders ALL samples of a tile, in windows of W consecutive samples; a
// window's tile_pixels x W paths are the work items its slots draw from. A finished sample is
// parked in the window's staging block, and when the window's last path is in, the warp adds
// its planes to the accumulation buffer in sample order — the reference's frame-by-frame
// accumulation (shader.wgsl:1367-1371). Windows are committed in the order they were opened,
// so every bit of the image is independent of scheduling.
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
#ifndef RT_LOCAL_MIN_BUSY
#define RT_LOCAL_MIN_BUSY 16   // a walk batch whose queue is empty suspends below this many walking lanes
#endif
constexpr int kThreads = RT_LOCAL_THREADS;
constexpr int kWarps = kThreads / 32;
constexpr int kMinBlocks = RT_LOCAL_MIN_BLOCKS;
constexpr uint32_t kMinBusy = RT_LOCAL_MIN_BUSY;
constexpr uint32_t kRing = 8;          // windows in flight per warp, at most
constexpr uint32_t kMaxCommit = 4;     // windows committed per maintenance step
constexpr uint32_t kMaxSlots = 256;    // per warp
constexpr uint32_t kShadow = 0x8000u;  // ray / walk queue entry: the slot's shadow ray
constexpr uint32_t kSlotMask = 0x7FFFu;
constexpr uint32_t kFull = 0xFFFFFFFFu;
// per-slot `pend` word: rays not settled yet | the closest-hit ray found something | the path is over
constexpr uint32_t kPendCount = 0xFFu, kPendHit = 0x100u, kPendDone = 0x200u;
constexpr uint32_t kStallLimit = 100000u;  // scheduler rounds without work before the warp gives up

enum Queue : uint32_t { Q_SHADE = 0, Q_FIN, Q_FREE, Q_RAY, Q_WALK, Q_COUNT };

// low byte of the path record's last word; the other bytes hold the ring entry and the
// pixel's position inside its tile
enum PathState : uint32_t {
    PATH_TRACED = 2,  // a closest-hit ray is in flight / has come back
    PATH_DONE = 4     // the path is over
};
constexpr uint32_t kFlagShadow = 1u;  // a shadow ray was cast for the last shaded hit

struct RingEntry {
    uint32_t tx, ty;     // tile
    uint32_t w0, count;  // first sample of the window, samples in it
    uint32_t left;       // paths of the window that have not parked their sample yet
    uint32_t pad[3];
};

struct WarpCtrl {
    uint32_t head[Q_COUNT], tail[Q_COUNT];  // running indices; entry i sits at i & (capacity - 1)
    uint32_t handed, limit;          // path handout: virtual indices < limit belong to open windows
    uint32_t ring_head, ring_tail;   // windows [head, tail) are open or waiting for their commit
    uint32_t tile_next, tile_end;    // tiles taken from the global counter, not yet opened
    uint32_t tile_valid, cur_tx, cur_ty, next_window;
    uint32_t exhausted;              // the global counter has run out
    uint32_t n_susp;                 // lanes holding a suspended walk
    uint32_t commit_n;
    uint32_t pad0;
    RingEntry commit[kMaxCommit];
    RingEntry ring[kRing];
    // profiling (LocalParams::profile): cycles per stage (+ bookkeeping), items and batches per stage
    unsigned long long cycles[5], items[4], batches[4];
};

struct LocalParams {
    FastScene scene;
    EnvView env;
    CameraParams camera;
    float4* accum;
    unsigned long long* counters;
    uint32_t width, height;
    uint32_t tiles_x, tiles_total;
    uint32_t tile_w_shift, tile_shift;  // tile = (1 << tile_w_shift) x (tile pixels >> tile_w_shift); pixels = 1 << tile_shift
    uint32_t sample_begin, sample_count;
    uint32_t max_bounces;
    uint32_t slots;          // path slots per warp (<= kMaxSlots)
    uint32_t qcap;           // queue capacity: power of two >= slots
    uint32_t window_shift;   // W = 1 << window_shift samples per window
    uint32_t ring_entries;   // power of two <= kRing
    uint32_t tile_grab;      // tiles taken per atomic on the global counter
    // per-warp pools, indexed by global warp * slots + slot
    float4* rOD;      // [2] ray origin (.w = closest analytic distance), direction (.w = bits rec of it)
    float4* pTLPI;    // [4] throughput + last_pdf, radiance + bits rng, pending light + bits flags,
                      //     bits (pixel, sample, bounce, state | ring entry << 8 | pixel in tile << 16)
    float4* sray_d;   // direction of the shadow ray
    float4* hit;      // (t, bits rec, u, v)
    float2* fb;       // brute-force fallback winner of the closest-hit ray
    uint8_t* occl;    // 1 when the shadow ray is blocked
    float4* staging;  // [warp][ring_entries][W][tile pixels] parked samples
    uint32_t* next_tile;
    uint32_t* error;  // set when a warp's scheduler found no work although paths were outstanding
    // optional: [0..4] cycles in shade, finish, pre-pass, walk, bookkeeping summed over warps, [5..8]
    // items those stages worked on, [9..12] their batches, [13] warps
    unsigned long long* profile;
};

// Shared-memory layout of a CTA: pre-pass program | flat program + triangle records | per warp:
// control block, pend words, suspended walks, walk masks, the five queues.
struct SmemLayout {
    uint32_t pre_f4, flat_f4, tri_f4;  // float4 counts
    uint32_t warp_bytes;
    size_t bytes;
};
inline __host__ __device__ SmemLayout smem_layout(const FastScene& sc, uint32_t slots, uint32_t qcap, bool flat) {
    (void)slots;
    SmemLayout l;
    l.pre_f4 = sc.hoisted ? sc.pre_leaves + 2u * sc.pre_boxes + 4u * sc.pre_prims : 0u;
    l.flat_f4 = flat ? kFlatFloat4 : 0u;
    l.tri_f4 = flat ? kTriStride * sc.flat_slots : 0u;
    uint32_t w = (uint32_t)((sizeof(WarpCtrl) + 15u) & ~(size_t)15u);
    w += qcap * 4u;                    // pend
    w += 8u * 32u * 4u;                // suspended walks
    if (flat) w += 2u * qcap * 4u;     // walk masks
    w += (3u * qcap + 2u * 2u * qcap) * 2u;  // shade, finish, free | ray, walk queues
    l.warp_bytes = (w + 15u) & ~15u;
    l.bytes = (size_t)(l.pre_f4 + l.flat_f4 + l.tri_f4) * 16u + (size_t)l.warp_bytes * kWarps;
    return l;
}

struct PathRegs {
    V3 T, L, pending;
    float last_pdf;
    uint32_t rng, flags;
    uint32_t occluded;
};

struct Warp {
    WarpCtrl* c;
    const float4* spre;
    const float4* sflat;
    const float4* stri;
    uint32_t* pend;    // [qcap]
    uint32_t* susp;    // [8][32]
    uint32_t* walkm;   // [2 qcap] leaf masks of the walk queue's rays (flat walk)
    uint16_t* q[Q_COUNT];
    uint32_t qmask[Q_COUNT];
    size_t pool;       // first global slot of this warp
    float4* stage;     // this warp's staging blocks
    uint32_t lane, lt;  // lane index, mask of the lanes below
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

// Appends one entry per requesting lane; `tail` is the warp-uniform running index of the queue,
// kept in a register for the length of a batch. Every lane of the warp must call it.
RT_DEVONLY uint32_t push(const Warp& k, Queue q, uint32_t& tail, bool want, uint32_t entry) {
    const unsigned m = __ballot_sync(kFull, want);
    const uint32_t at = tail + (uint32_t)__popc(m & k.lt);
    if (want) k.q[q][at & k.qmask[q]] = (uint16_t)entry;
    tail += (uint32_t)__popc(m);
    return at;
}

// One ray of `slot` is settled. Returns where the slot goes when it was the last one:
// 0 = rays still out, 1 = shade, 2 = finish.
RT_DEVONLY uint32_t settle(const Warp& k, uint32_t slot, bool found) {
    const uint32_t delta = (found ? kPendHit : 0u) + 0xFFFFFFFFu;
    const uint32_t now = atomicAdd(&k.pend[slot], delta) + delta;
    if (now & kPendCount) return 0u;
    if (now & kPendDone) return 2u;
    return (now & kPendHit) ? 1u : 2u;
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
RT_DEVONLY void pre_ray(const LocalParams& p, const Warp& k, const RayCtx& r, bool shadow, PreOut<FLAT>& out) {
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

// ---- stage: shade ----
template <bool FLAT>
RT_DEVONLY void batch_shade(const LocalParams& p, const Warp& k, uint32_t& n_rays, uint32_t& n_shadow) {
    WarpCtrl* c = k.c;
    const uint32_t head = c->head[Q_SHADE];
    uint32_t n = c->tail[Q_SHADE] - head;
    n = n < 32u ? n : 32u;
    uint32_t t_ray = c->tail[Q_RAY], t_fin = c->tail[Q_FIN];
    const bool on = k.lane < n;
    const uint32_t slot = on ? k.q[Q_SHADE][(head + k.lane) & k.qmask[Q_SHADE]] : 0u;
    const size_t g = k.pool + slot;
    bool emit_closest = false, emit_shadow = false, fin_now = false;
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
            // With a shadow ray in flight the sample is finished when that ray is settled.
            I.w = (I.w & ~0xFFu) | PATH_DONE;
            fin_now = !emit_shadow;
            if (emit_shadow) p.rOD[2u * g + 0u] = point4;
        } else {
            emit_closest = true;
            st8<0>(&p.rOD[2u * g], point4, make_float4(so.next_dir.x, so.next_dir.y, so.next_dir.z, u_as_f(kNone)));
            n_rays++;
        }
        if (emit_closest || emit_shadow)
            k.pend[slot] = (emit_closest ? 1u : 0u) + (emit_shadow ? 1u : 0u) + (done ? kPendDone : 0u);
        store_path(p, g, s, I);
    }
    push(k, Q_RAY, t_ray, emit_closest, slot);
    push(k, Q_RAY, t_ray, emit_shadow, slot | kShadow);
    push(k, Q_FIN, t_fin, fin_now, slot);
    if (k.lane == 0u) {
        c->head[Q_SHADE] = head + n;
        c->tail[Q_RAY] = t_ray;
        c->tail[Q_FIN] = t_fin;
    }
}

// The result of a closest-hit ray that needed no walk (or whose walk found nothing).
RT_DEVONLY bool store_closest(const LocalParams& p, const RayCtx& r, size_t g, const Fallback& fb, FHit best) {
    resolve_fallback(p.scene, r, fb, best);
    p.hit[g] = make_float4(best.t, u_as_f(best.rec), best.u, best.v);
    return best.rec != kNone;
}

// ---- stage: finish (and start the next path) ----
template <bool FLAT>
RT_DEVONLY void batch_finish(const LocalParams& p, const Warp& k, float sin_half_fov, uint32_t& n_paths,
                             uint32_t& n_rays) {
    WarpCtrl* c = k.c;
    const uint32_t h_fin = c->head[Q_FIN], h_free = c->head[Q_FREE];
    uint32_t n_fin = c->tail[Q_FIN] - h_fin;
    n_fin = n_fin < 32u ? n_fin : 32u;
    uint32_t handed = c->handed;
    const uint32_t limit = c->limit;
    // idle slots join only while there are paths to hand out
    uint32_t n_idle = c->tail[Q_FREE] - h_free;
    {
        const uint32_t room = 32u - n_fin, avail = limit - handed;
        n_idle = n_idle < room ? n_idle : room;
        n_idle = n_idle < avail ? n_idle : avail;
    }
    uint32_t t_walk = c->tail[Q_WALK], t_shade = c->tail[Q_SHADE], t_fin = c->tail[Q_FIN], t_free = c->tail[Q_FREE];
    const bool had_path = k.lane < n_fin;
    const bool on = k.lane < n_fin + n_idle;
    uint32_t slot = 0u;
    if (had_path) slot = k.q[Q_FIN][(h_fin + k.lane) & k.qmask[Q_FIN]];
    else if (on) slot = k.q[Q_FREE][(h_free + k.lane - n_fin) & k.qmask[Q_FREE]];
    const size_t g = k.pool + slot;
    const uint32_t tile_px_mask = (1u << p.tile_shift) - 1u;

    PathRegs s;
    s.rng = 0u;
    uint4 I = make_uint4(0u, 0u, 0u, 0u);
    if (had_path) {
        float4 d4;
        load_path(p, g, s, I);
        d4 = p.rOD[2u * g + 1u];
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
        __stcs(&k.stage[((((size_t)rs << p.window_shift) + kk) << p.tile_shift) + pix],
               make_float4(s.L.x, s.L.y, s.L.z, 0.0f));
        atomicSub(&c->ring[rs].left, 1u);
    }
    // ---- the next path indices: handed out in lane order; an index that names no pixel-sample
    // (partial tile, short last window) is skipped and the lane draws again ----
    bool need = on, got = false;
    uint32_t px = 0, py = 0, sample = 0, rs = 0, pix = 0;
    const uint32_t eshift = p.tile_shift + p.window_shift;
    for (;;) {
        const unsigned m = __ballot_sync(kFull, need);
        if (!m || handed == limit) break;
        const uint32_t rank = (uint32_t)__popc(m & k.lt);
        uint32_t give = (uint32_t)__popc(m);
        give = give < limit - handed ? give : limit - handed;
        if (need && rank < give) {
            const uint32_t h = handed + rank;
            const uint32_t i = h & ((1u << eshift) - 1u);
            rs = (h >> eshift) & (p.ring_entries - 1u);
            const RingEntry& en = c->ring[rs];
            const uint32_t kk = i >> p.tile_shift;
            pix = i & tile_px_mask;
            px = (en.tx << p.tile_w_shift) + (pix & ((1u << p.tile_w_shift) - 1u));
            py = en.ty * ((1u << p.tile_shift) >> p.tile_w_shift) + (pix >> p.tile_w_shift);
            if (kk < en.count && px < p.width && py < p.height) {
                sample = en.w0 + kk;
                got = true;
                need = false;
            }
        }
        handed += give;
    }
    bool to_walk = false, to_shade = false, to_fin = false;
    uint32_t mask = 0u;
    if (got) {
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
        if (to_walk) {
            p.fb[g] = make_float2(pre.fb.t, u_as_f(pre.fb.index));
            k.pend[slot] = 1u;
        } else {
            to_shade = store_closest(p, r, g, pre.fb, pre.best);
            to_fin = !to_shade;
        }
        store_path(p, g, s, I);
        n_paths++;
        n_rays++;
    }
    {
        const uint32_t at = push(k, Q_WALK, t_walk, to_walk, slot);
        if (FLAT && to_walk) k.walkm[at & k.qmask[Q_WALK]] = mask;
    }
    push(k, Q_SHADE, t_shade, to_shade, slot);
    // The finish queue is read from h_fin + n_fin on, so entries appended here are not the ones
    // this batch is working on.
    push(k, Q_FIN, t_fin, to_fin, slot);
    push(k, Q_FREE, t_free, on && !got, slot);
    if (k.lane == 0u) {
        c->head[Q_FIN] = h_fin + n_fin;
        c->head[Q_FREE] = h_free + n_idle;
        c->tail[Q_WALK] = t_walk;
        c->tail[Q_SHADE] = t_shade;
        c->tail[Q_FIN] = t_fin;
        c->tail[Q_FREE] = t_free;
        c->handed = handed;
    }
}

// ---- stage: pre-pass of the rays the shade stage emitted ----
template <bool FLAT>
RT_DEVONLY void batch_prepass(const LocalParams& p, const Warp& k) {
    WarpCtrl* c = k.c;
    const uint32_t head = c->head[Q_RAY];
    uint32_t n = c->tail[Q_RAY] - head;
    n = n < 32u ? n : 32u;
    uint32_t t_walk = c->tail[Q_WALK], t_shade = c->tail[Q_SHADE], t_fin = c->tail[Q_FIN];
    const bool on = k.lane < n;
    const uint32_t e = on ? k.q[Q_RAY][(head + k.lane) & k.qmask[Q_RAY]] : 0u;
    const uint32_t slot = e & kSlotMask;
    const bool shadow = (e & kShadow) != 0u;
    const size_t g = k.pool + slot;
    bool to_walk = false;
    uint32_t route = 0u, mask = 0u;
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
            if (!to_walk) route = settle(k, slot, false);
        } else if (to_walk) {
            p.rOD[2u * g + 0u].w = pre.best.t;
            p.rOD[2u * g + 1u].w = u_as_f(pre.best.rec);
            p.fb[g] = make_float2(pre.fb.t, u_as_f(pre.fb.index));
        } else {
            route = settle(k, slot, store_closest(p, r, g, pre.fb, pre.best));
        }
    }
    {
        const uint32_t at = push(k, Q_WALK, t_walk, to_walk, e);
        if (FLAT && to_walk) k.walkm[at & k.qmask[Q_WALK]] = mask;
    }
    push(k, Q_SHADE, t_shade, route == 1u, slot);
    push(k, Q_FIN, t_fin, route == 2u, slot);
    if (k.lane == 0u) {
        c->head[Q_RAY] = head + n;
        c->tail[Q_WALK] = t_walk;
        c->tail[Q_SHADE] = t_shade;
        c->tail[Q_FIN] = t_fin;
    }
}

// A walked ray's result. Returns the slot's route (see settle()).
RT_DEVONLY uint32_t walk_result(const LocalParams& p, const Warp& k, const RayCtx& r, uint32_t slot, bool any_hit,
                                FHit best) {
    const size_t g = k.pool + slot;
    if (any_hit) {
        if (best.rec != kNone) p.occl[g] = 1;
        return settle(k, slot, false);
    }
    if (best.rec == kNone) {
        const float2 f = p.fb[g];
        Fallback fb;
        fb.t = f.x;
        fb.index = f_as_u(f.y);
        resolve_fallback(p.scene, r, fb, best);
    }
    p.hit[g] = make_float4(best.t, u_as_f(best.rec), best.u, best.v);
    return settle(k, slot, best.rec != kNone);
}

// ---- stage: walk. Persistent lanes; a lane whose ray is finished takes the next entry. With
// `flush` the batch runs until queue and lanes are empty; otherwise it leaves once the queue is
// empty and fewer than kMinBusy lanes still walk, parking those lanes' state for the next batch.
template <bool FLAT>
RT_DEVONLY void batch_walk(const LocalParams& p, const Warp& k, uint2* stack, bool flush) {
    WarpCtrl* c = k.c;
    uint32_t head = c->head[Q_WALK];
    const uint32_t total = c->tail[Q_WALK];
    uint32_t t_shade = c->tail[Q_SHADE], t_fin = c->tail[Q_FIN];
    const uint32_t lane = k.lane;

    bool busy = false, any_hit = false;
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

    // ---- resume the walks the last batch parked (each lane its own: the stack never moved) ----
    if (c->n_susp) {
        const uint32_t tag = k.susp[lane];
        if (tag != kNone) {
            slot = tag & kSlotMask;
            any_hit = (tag & kShadow) != 0u;
            const size_t g = k.pool + slot;
            const float4 o4 = p.rOD[2u * g + 0u];
            const float4 d4 = any_hit ? p.sray_d[g] : p.rOD[2u * g + 1u];
            r = ray_setup(xyz(o4), xyz(d4));
            best.t = u_as_f(k.susp[32u + lane]);
            best.rec = k.susp[64u + lane];
            best.u = u_as_f(k.susp[96u + lane]);
            best.v = u_as_f(k.susp[128u + lane]);
            const uint32_t a = k.susp[160u + lane], b = k.susp[192u + lane];
            const float l = u_as_f(k.susp[224u + lane]);
            if (FLAT) {
                mask = a;
                cur = b & 0xFFFFu;
                end = b >> 16;
                limit = l;
            } else {
                w.cur = a;
                w.sp = b;
                w.limit = l;
            }
            busy = true;
        }
    }

    for (;;) {
        // ---- refill: idle lanes take the next entries ----
        const unsigned idle = __ballot_sync(kFull, !busy);
        if (idle && head != total) {
            uint32_t take = (uint32_t)__popc(idle);
            take = take < total - head ? take : total - head;
            const uint32_t rank = (uint32_t)__popc(idle & k.lt);
            if (!busy && rank < take) {
                const uint32_t idx = (head + rank) & k.qmask[Q_WALK];
                const uint32_t e = k.q[Q_WALK][idx];
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
            head += take;
        }
        const uint32_t walking = (uint32_t)__popc(__ballot_sync(kFull, busy));
        if (!walking) break;
        if (!flush && head == total && walking < kMinBusy) break;

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
            if (__all_sync(kFull, !busy || r.plain)) {
                while (busy && w.cur < kLeafBit) walk_node<true>(p.scene, r, w, stack);
            } else {
                while (busy && w.cur < kLeafBit) walk_node<false>(p.scene, r, w, stack);
            }
            if (busy && w.cur != kNone) settled = walk_leaf(p.scene, r, any_hit, best, w, stack);
            finished = busy && (settled || w.cur == kNone);
        }

        // ---- lanes whose ray is finished hand over their result ----
        uint32_t route = 0u;
        if (finished) {
            route = walk_result(p, k, r, slot, any_hit, best);
            busy = false;
        }
        push(k, Q_SHADE, t_shade, route == 1u, slot);
        push(k, Q_FIN, t_fin, route == 2u, slot);
    }

    // ---- park what is still walking ----
    const unsigned left = __ballot_sync(kFull, busy);
    k.susp[lane] = busy ? (slot | (any_hit ? kShadow : 0u)) : kNone;
    if (busy) {
        k.susp[32u + lane] = f_as_u(best.t);
        k.susp[64u + lane] = best.rec;
        k.susp[96u + lane] = f_as_u(best.u);
        k.susp[128u + lane] = f_as_u(best.v);
        k.susp[160u + lane] = FLAT ? mask : w.cur;
        k.susp[192u + lane] = FLAT ? (cur | (end << 16)) : w.sp;
        k.susp[224u + lane] = f_as_u(FLAT ? limit : w.limit);
    }
    if (lane == 0u) {
        c->n_susp = (uint32_t)__popc(left);
        c->head[Q_WALK] = head;
        c->tail[Q_SHADE] = t_shade;
        c->tail[Q_FIN] = t_fin;
    }
}

// ---- window bookkeeping, by lane 0 ----
RT_DEVONLY void maintain(const LocalParams& p, WarpCtrl* c) {
    const uint32_t E = 1u << (p.tile_shift + p.window_shift);
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
    // open windows until every slot could draw a path
    const uint32_t W = 1u << p.window_shift;
    const uint32_t tile_w = 1u << p.tile_w_shift, tile_h = (1u << p.tile_shift) >> p.tile_w_shift;
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
        // a window is a block of E consecutive virtual path indices: the handout must stand at
        // the block's first index when the window before it was cut short by nothing — it never
        // is: indices past a window's real paths are drawn and skipped (batch_finish)
        RingEntry& e = c->ring[c->ring_tail & (p.ring_entries - 1u)];
        e.tx = c->cur_tx;
        e.ty = c->cur_ty;
        const uint32_t first = c->next_window * W;
        e.w0 = p.sample_begin + first;
        e.count = p.sample_count - first < W ? p.sample_count - first : W;
        const uint32_t vw = p.width - e.tx * tile_w < tile_w ? p.width - e.tx * tile_w : tile_w;
        const uint32_t vh = p.height - e.ty * tile_h < tile_h ? p.height - e.ty * tile_h : tile_h;
        e.left = e.count * vw * vh;
        c->next_window++;
        if (c->next_window * W >= p.sample_count) c->tile_valid = 0u;
        c->ring_tail++;
        c->limit += E;
    }
}

// Adds the staged planes of the windows in c->commit to the accumulation buffer, in sample
// order (shader.wgsl:1367-1371: total += sample, alpha = 1). The same lane owns a pixel in every
// window of a tile, and windows are committed in the order they were opened.
RT_DEVONLY void commit_windows(const LocalParams& p, const Warp& k) {
    const uint32_t n = k.c->commit_n;
    const uint32_t tile_px = 1u << p.tile_shift;
    const uint32_t tile_h = tile_px >> p.tile_w_shift;
    for (uint32_t j = 0; j < n; j++) {
        const RingEntry& e = k.c->commit[j];
        for (uint32_t pix = k.lane; pix < tile_px; pix += 32u) {
            const uint32_t px = (e.tx << p.tile_w_shift) + (pix & ((1u << p.tile_w_shift) - 1u));
            const uint32_t py = e.ty * tile_h + (pix >> p.tile_w_shift);
            if (px >= p.width || py >= p.height) continue;
            float4* at = &p.accum[(size_t)py * p.width + px];
            float4 a = *at;
            const float4* plane = k.stage + ((((size_t)e.pad[0] << p.window_shift)) << p.tile_shift) + pix;
            for (uint32_t s = 0; s < e.count; s++) {
                const float4 v = __ldcs(plane + ((size_t)s << p.tile_shift));
                a.x = a.x + v.x;
                a.y = a.y + v.y;
                a.z = a.z + v.z;
            }
            a.w = 1.0f;
            *at = a;
        }
    }
}

template <bool FLAT>
__global__ void __launch_bounds__(kThreads, kMinBlocks) local_wave_kernel(const LocalParams p) {
    extern __shared__ float4 smem[];
    const SmemLayout lay = smem_layout(p.scene, p.slots, p.qcap, FLAT);
    const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    Warp k;
    {
        const float4* q = smem;
        k.spre = q;
        q += lay.pre_f4;
        k.sflat = q;
        q += lay.flat_f4;
        k.stri = q;
        q += lay.tri_f4;
        char* b = (char*)q + (size_t)warp * lay.warp_bytes;
        k.c = (WarpCtrl*)b;
        b += (sizeof(WarpCtrl) + 15u) & ~(size_t)15u;
        k.pend = (uint32_t*)b;
        b += p.qcap * 4u;
        k.susp = (uint32_t*)b;
        b += 8u * 32u * 4u;
        k.walkm = (uint32_t*)b;
        if (FLAT) b += 2u * p.qcap * 4u;
        uint16_t* h = (uint16_t*)b;
        k.q[Q_SHADE] = h;
        h += p.qcap;
        k.q[Q_FIN] = h;
        h += p.qcap;
        k.q[Q_FREE] = h;
        h += p.qcap;
        k.q[Q_RAY] = h;
        h += 2u * p.qcap;
        k.q[Q_WALK] = h;
        k.qmask[Q_SHADE] = k.qmask[Q_FIN] = k.qmask[Q_FREE] = p.qcap - 1u;
        k.qmask[Q_RAY] = k.qmask[Q_WALK] = 2u * p.qcap - 1u;
    }
    const size_t gwarp = (size_t)blockIdx.x * kWarps + warp;
    k.pool = gwarp * p.slots;
    k.stage = p.staging + gwarp * ((((size_t)p.ring_entries << p.window_shift)) << p.tile_shift);
    k.lane = lane;
    k.lt = (1u << lane) - 1u;
    WarpCtrl* c = k.c;

    // ---- set-up: scene programs into shared memory (the one CTA-wide step), every slot idle ----
    {
        float4* spre = smem;
        for (uint32_t i = threadIdx.x; i < lay.pre_f4; i += kThreads) spre[i] = p.scene.pre[i];
        if (FLAT) {
            float4* sflat = spre + lay.pre_f4;
            for (uint32_t i = threadIdx.x; i < lay.flat_f4; i += kThreads) sflat[i] = p.scene.flat[i];
            float4* stri = sflat + lay.flat_f4;
            for (uint32_t i = threadIdx.x; i < lay.tri_f4; i += kThreads) stri[i] = p.scene.tri[i];
        }
        for (uint32_t i = lane; i < p.slots; i += 32u) k.q[Q_FREE][i] = (uint16_t)i;
        k.susp[lane] = kNone;
        if (lane == 0u) {
            for (uint32_t i = 0; i < Q_COUNT; i++) c->head[i] = c->tail[i] = 0u;
            c->tail[Q_FREE] = p.slots;
            c->handed = c->limit = 0u;
            c->ring_head = c->ring_tail = 0u;
            c->tile_next = c->tile_end = 0u;
            c->tile_valid = 0u;
            c->cur_tx = c->cur_ty = c->next_window = 0u;
            c->exhausted = 0u;
            c->n_susp = 0u;
            c->commit_n = 0u;
            for (int i = 0; i < 5; i++) c->cycles[i] = 0ull;
            for (int i = 0; i < 4; i++) c->items[i] = c->batches[i] = 0ull;
        }
    }
    __syncthreads();

    const float sin_half_fov = rt_sin(p.camera.fov_y / 2.0f);
    uint32_t n_paths = 0, n_rays = 0, n_shadow = 0;
    uint2 stack[FLAT ? 1 : kStackCap];
    const bool prof = p.profile != nullptr;
    unsigned long long t_last = prof ? clock64() : 0ull;
    auto lap = [&](int what, uint32_t items) {  // lane 0 keeps the books
        if (lane == 0u) {
            const unsigned long long now = clock64();
            c->cycles[what] += now - t_last;
            t_last = now;
            if (what < 4) {
                c->items[what] += items;
                c->batches[what]++;
            }
        }
    };

    uint32_t stall = 0u;
    for (;;) {
        __syncwarp();
        // ---- windows: commit the finished ones, open new ones while slots could use more paths ----
        {
            const bool can_commit = c->ring_head != c->ring_tail && c->ring[c->ring_head & (p.ring_entries - 1u)].left == 0u;
            const bool can_open = !c->exhausted && c->ring_tail - c->ring_head < p.ring_entries &&
                                  c->limit - c->handed < p.slots;
            if (can_commit || can_open) {
                if (lane == 0u) maintain(p, c);
                __syncwarp();
                commit_windows(p, k);  // reads staging blocks the coming finish batches may reuse
                __syncwarp();
            }
        }
        const uint32_t n_shade = c->tail[Q_SHADE] - c->head[Q_SHADE], n_fin = c->tail[Q_FIN] - c->head[Q_FIN];
        const uint32_t n_free = c->tail[Q_FREE] - c->head[Q_FREE], n_ray = c->tail[Q_RAY] - c->head[Q_RAY];
        const uint32_t n_walk = c->tail[Q_WALK] - c->head[Q_WALK] + c->n_susp;
        const uint32_t avail = c->limit - c->handed;
        const uint32_t n_spawn = n_fin + (n_free < avail ? n_free : avail);
        if (prof) lap(4, 0u);
        int stage;
        if (n_walk >= 32u) stage = 3;
        else if (n_ray >= 32u) stage = 2;
        else if (n_shade >= 32u) stage = 0;
        else if (n_spawn >= 32u) stage = 1;
        else {
            // nothing fills a batch: the fullest stage, the walk run to its end
            uint32_t most = n_walk;
            stage = 3;
            if (n_ray > most) most = n_ray, stage = 2;
            if (n_shade > most) most = n_shade, stage = 0;
            if (n_spawn > most) most = n_spawn, stage = 1;
            if (most == 0u) {
                if (c->exhausted && c->ring_head == c->ring_tail) break;  // every window is committed
                if (++stall > kStallLimit) {
                    if (lane == 0u) atomicExch(p.error, 1u);
                    break;
                }
                continue;
            }
            if (stage == 3) stage = 4;  // flush
        }
        stall = 0u;
        switch (stage) {
            case 0:
                batch_shade<FLAT>(p, k, n_rays, n_shadow);
                if (prof) lap(0, n_shade < 32u ? n_shade : 32u);
                break;
            case 1:
                batch_finish<FLAT>(p, k, sin_half_fov, n_paths, n_rays);
                if (prof) lap(1, n_spawn < 32u ? n_spawn : 32u);
                break;
            case 2:
                batch_prepass<FLAT>(p, k);
                if (prof) lap(2, n_ray < 32u ? n_ray : 32u);
                break;
            default:
                batch_walk<FLAT>(p, k, stack, stage == 4);
                if (prof) lap(3, n_walk - (c->tail[Q_WALK] - c->head[Q_WALK] + c->n_susp));
                break;
        }
    }
    fast_flush_counts(n_paths, n_rays, n_shadow, p.counters);
    if (prof && lane == 0u) {
        for (int i = 0; i < 5; i++) atomicAdd(&p.profile[i], c->cycles[i]);
        for (int i = 0; i < 4; i++) atomicAdd(&p.profile[5 + i], c->items[i]);
        for (int i = 0; i < 4; i++) atomicAdd(&p.profile[9 + i], c->batches[i]);
        atomicAdd(&p.profile[13], 1ull);
    }
}

}  // namespace rtl
