// rt_wave.cuh — the wavefront render pipeline: the path loop of shader.wgsl:1213-1303 cut at
// its ray casts and run as four small, dense kernels per "wave".
//
// Why: as one kernel (rt_fast.cuh's render_fast_kernel) the loop body is ~60 KB of SASS — far
// beyond the 32 KB instruction cache level — and a warp's lanes sit in different parts of
// it: ncu showed 40 % issue utilisation with "no instruction" as the top stall and 8 to 10
// of 32 lanes active. Here every kernel does ONE thing for a compacted list of path slots,
// so all 32 lanes of a warp run the same code:
//   wave_shade_kernel    slots whose ray hit -> settle last shadow ray, shade (emission,
//                        next-event set-up, BSDF sample), emit the bounce ray + shadow ray
//   wave_finish_kernel   slots whose path is over / ray escaped / are new -> settle, sky light,
//                        park the finished sample, take the next path (camera ray) and run that
//                        ray's analytic pre-pass on the spot (the kernel mostly waits on memory)
//   wave_prepass_kernel  ray queue -> the analytic pre-pass (rt_fast.cuh) of the rays the shade
//                        kernel emitted; rays with no triangle to test are finished here, the
//                        others go, compacted, into the walk queue
//   wave_walk_kernel     walk queue -> the triangle walk alone (a few KB of code): persistent
//                        warps whose lanes pull the next ray the moment their own is finished
//                        (the queue has no holes: a lane that drew a settled entry would sit out
//                        a whole round)
// shade and finish find their slots by sweeping a one-byte stage array in slot order and
// compacting inside the warp (ballot + a small shared-memory list), so a warp's 32 lanes work
// on 32 NEARBY slots: full lanes and coalesced state traffic, no global lists.
// The bounce ray and the shadow ray of one hit do not depend on each other and travel in
// the same wave. The slots are split into two groups whose waves run half a wave apart on
// two streams, so one group's memory-bound shade / finish overlaps the other's issue-bound
// pre-pass / walk. Path state lives in global memory (~162 B per slot).
//
// Paths are independent work items: a slot that finishes a path takes the next path index
// from a global counter (sample-major, pixels in 8x8-tile order inside a sample), so the
// slot population stays full until the work runs out. To keep every bit of the result, the
// radiance of sample k of a pixel is parked in a staging plane and `wave_accumulate_kernel`
// adds the planes to the accumulation buffer in sample order at the end of a batch — the
// reference's frame-by-frame order (shader.wgsl:1367-1371) — whatever order the paths
// finished in.
#pragma once

#include "rt_fast.cuh"

namespace rtw {
using namespace rtd;
using namespace rtf;

constexpr uint32_t kShadowBit = 0x80000000u;  // ray queue entry: the slot's shadow ray
constexpr uint32_t kDeadEntry = 0xFFFFFFFFu;  // ray queue entry settled by the pre-pass
#ifndef RT_SHADE_THREADS
#define RT_SHADE_THREADS 128
#endif
#ifndef RT_FINISH_THREADS
#define RT_FINISH_THREADS 128
#endif
#ifndef RT_PREPASS_THREADS
#define RT_PREPASS_THREADS 256
#endif
#ifndef RT_WALK_THREADS
#define RT_WALK_THREADS 256
#endif
constexpr int kShadeThreads = RT_SHADE_THREADS;
constexpr int kFinishThreads = RT_FINISH_THREADS;
constexpr int kPrepassThreads = RT_PREPASS_THREADS;
constexpr int kWalkThreads = RT_WALK_THREADS;

// state word of a slot (pTLPI[4 * slot + 3].w)
enum SlotState : uint32_t {
    SLOT_RETIRED = 0,
    SLOT_NEW = 1,     // needs its first camera ray
    SLOT_TRACED = 2,  // a closest-hit ray is in flight / has come back
    SLOT_ENDING = 3,  // the path is over but its last shadow ray is still in flight
    SLOT_DONE = 4     // the path is over, nothing in flight
};
constexpr uint32_t kFlagShadow = 1u;  // a shadow ray was cast for the last shaded hit

// stage[slot]
enum SlotStage : uint8_t {
    STAGE_IDLE = 0,         // retired
    STAGE_SHADE = 1,        // its closest-hit ray hit something
    STAGE_FINISH = 2,       // its ray escaped, its path ended, or it is new
    STAGE_FLIGHT = 3,       // a closest-hit ray is in flight
    STAGE_FINISH_NEXT = 4   // path over, last shadow ray in flight: finish in the next wave
};

// The ray queue is double-buffered by wave parity P = wave & 1: half P (ray_n[P] entries) is
// filled by shade(w) + finish(w) and read by prepass(w), which writes the rays that still need
// the triangle walk, compacted, into the idle half Q (walk_n entries) for walk(w) to pull with
// ray_head[P]. shade(w) clears the counters of half Q and walk_n.
struct WaveCtrl {
    uint32_t ray_n[2];
    uint32_t ray_head[2];
    uint32_t live;       // slots not yet retired
    uint32_t walk_n;     // rays the pre-pass kernel handed on to the walk kernel (this wave)
    uint32_t pad[2];
};

struct WaveBuffers {
    // Vectors that are always used together are neighbours in memory (a 32-byte ray pair, a
    // 64-byte path record), so the scattered slots a warp works on waste no half sectors.
    // (One record per slot holding everything, 128 or 64 + 64 bytes, measured slower: the
    // pre-pass and the walk touch only the ray vectors.)
    float4* rOD;      // [S][2] ray origin (.w = closest analytic distance from the pre-pass),
                      //        direction of the closest-hit ray (.w = bits rec of that analytic hit)
    float4* sray_d;   // [S] direction of the shadow ray
    float2* fb;       // [S] brute-force fallback winner of the closest-hit ray (t, bits index)
    float4* hit;      // [S] (t, bits rec, u, v) written by the walk kernel
    uint8_t* occl;    // [S] 1 when the shadow ray is blocked
    float4* pTLPI;    // [S][4] 64 contiguous bytes only shade / finish touch: throughput + last_pdf,
                      //        radiance of the current sample + bits rng, pending next-event light +
                      //        bits flags, bits (pixel, sample, bounce, state)
    uint8_t* stage;   // [S] what the slot needs next (SlotStage)
    uint32_t* rayq;   // [2][2S]
    WaveCtrl* ctrl;
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
    float4* staging;       // [batch_count][width * height] radiance of each finished sample
    uint32_t* next_path;   // next path index of the batch to hand out (shared by all slot groups)
    uint32_t slots;
    uint32_t batch_begin, batch_count;  // the samples of this batch
    uint32_t path_total;   // batch_count * pixel_ids path indices (ids of partial tiles are skipped)
    uint32_t max_bounces;
    uint32_t wave;
};

RT_DEVONLY bool pixel_of_id(const WaveParams& p, uint32_t id, uint32_t& px, uint32_t& py) {
    uint32_t tile = id >> 6, in = id & 63u;
    px = (tile % p.tiles_x) * 8u + (in & 7u);
    py = (tile / p.tiles_x) * 8u + (in >> 3);
    return px < p.width && py < p.height;
}


// Appends one entry per requesting lane with a single atomic per warp. Every lane of the
// warp must call it.
RT_DEVONLY void list_push(uint32_t* list, uint32_t* count, bool want, uint32_t entry) {
    const unsigned m = __ballot_sync(0xFFFFFFFFu, want);
    if (!m) return;
    const uint32_t lane = threadIdx.x & 31u;
    const int leader = __ffs(m) - 1;
    uint32_t base = 0;
    if ((int)lane == leader) base = atomicAdd(count, (uint32_t)__popc(m));
    base = __shfl_sync(0xFFFFFFFFu, base, leader);
    if (want) list[base + __popc(m & ((1u << lane) - 1u))] = entry;
}

// Path state streams through memory once per wave: load / store it "cache streaming" (evict
// first) so that it does not push the environment map, alias table and scene out of L2.
#ifndef RT_WAVE_STREAM
#define RT_WAVE_STREAM 0
#endif
// L2 eviction priority of the 256-bit state accesses of shade / finish: 1 = evict first.
#ifndef RT_L2_HINT_PATH
#define RT_L2_HINT_PATH 1
#endif
#ifndef RT_L2_HINT_RAY
#define RT_L2_HINT_RAY 0
#endif
template <typename T>
RT_DEVONLY T lds(const T* ptr) {
#if RT_WAVE_STREAM
    return __ldcs(ptr);
#else
    return *ptr;
#endif
}
template <typename T>
RT_DEVONLY void sts(T* ptr, T v) {
#if RT_WAVE_STREAM
    __stcs(ptr, v);
#else
    *ptr = v;
#endif
}

// Two optional entries per lane (a and b) with ONE atomic per warp: the a entries first, then
// the b entries. Every lane of the warp must call it.
RT_DEVONLY void list_push2(uint32_t* list, uint32_t* count, bool want_a, uint32_t entry_a, bool want_b,
                           uint32_t entry_b) {
    const unsigned ma = __ballot_sync(0xFFFFFFFFu, want_a), mb = __ballot_sync(0xFFFFFFFFu, want_b);
    if (!(ma | mb)) return;
    const uint32_t lane = threadIdx.x & 31u;
    const uint32_t na = (uint32_t)__popc(ma);
    uint32_t base = 0;
    if (lane == 0u) base = atomicAdd(count, na + (uint32_t)__popc(mb));
    base = __shfl_sync(0xFFFFFFFFu, base, 0);
    const unsigned below = (1u << lane) - 1u;
    if (want_a) list[base + __popc(ma & below)] = entry_a;
    if (want_b) list[base + na + __popc(mb & below)] = entry_b;
}

__global__ void wave_init_kernel(const WaveParams p) {
    uint32_t slot = blockIdx.x * blockDim.x + threadIdx.x;
    if (slot < p.slots) {
        sts((uint4*)&p.b.pTLPI[4u * (size_t)slot + 3u], make_uint4(0u, 0u, 0u, SLOT_NEW));
        p.b.stage[slot] = STAGE_FINISH;  // every slot starts its first path in wave 0's finish pass
    }
    if (slot == 0) {
        WaveCtrl c;
        memset(&c, 0, sizeof c);
        c.live = p.slots;
        *p.b.ctrl = c;
    }
}

struct PathRegs {
    V3 T, L, pending;
    float last_pdf;
    uint32_t rng, flags;
    uint32_t occluded;  // the slot's occlusion byte, fetched with the rest (no dependent load)
};

// The 64-byte path record with two 256-bit loads / stores.
RT_DEVONLY void load_path(const WaveParams& p, uint32_t slot, PathRegs& s, uint4& I) {
    float4 t4, l4, p4, i4;
    ld8w<RT_L2_HINT_PATH>(&p.b.pTLPI[4u * (size_t)slot + 0u], t4, l4);
    ld8w<RT_L2_HINT_PATH>(&p.b.pTLPI[4u * (size_t)slot + 2u], p4, i4);
    s.T = xyz(t4);
    s.last_pdf = t4.w;
    s.L = xyz(l4);
    s.rng = f_as_u(l4.w);
    s.pending = xyz(p4);
    s.flags = f_as_u(p4.w);
    s.occluded = p.b.occl[slot];
    I = make_uint4(f_as_u(i4.x), f_as_u(i4.y), f_as_u(i4.z), f_as_u(i4.w));
}

RT_DEVONLY void store_path(const WaveParams& p, uint32_t slot, const PathRegs& s, const uint4 I) {
    st8<RT_L2_HINT_PATH>(&p.b.pTLPI[4u * (size_t)slot + 0u], make_float4(s.T.x, s.T.y, s.T.z, s.last_pdf),
        make_float4(s.L.x, s.L.y, s.L.z, u_as_f(s.rng)));
    st8<RT_L2_HINT_PATH>(&p.b.pTLPI[4u * (size_t)slot + 2u], make_float4(s.pending.x, s.pending.y, s.pending.z, u_as_f(s.flags)),
        make_float4(u_as_f(I.x), u_as_f(I.y), u_as_f(I.z), u_as_f(I.w)));
}

// shader.wgsl:1249-1264: the last next-event sample counts iff nothing blocked its ray.
RT_DEVONLY void settle_shadow(const WaveParams& p, uint32_t slot, PathRegs& s) {
    if ((s.flags & kFlagShadow) && s.occluded == 0u) s.L = s.L + s.pending;
    s.flags = 0u;
}

#ifndef RT_WAVE_SHADE_MIN_BLOCKS
#define RT_WAVE_SHADE_MIN_BLOCKS 8
#endif

// Sweeps `stage` in slot order; `body(slot, on)` runs with up to 32 compacted slots per
// warp call (on == false for lanes without one). Leftovers are carried to the next chunk so
// the lanes stay full; every lane of the warp reaches every body call. The body has ONE call
// site (the last, partial batch goes through the same loop): inlined twice it doubled the
// kernel's footprint in the instruction cache.
template <int THREADS, typename Body>
RT_DEVONLY void sweep_stage(const WaveParams& p, uint8_t want, uint8_t convert_from, uint8_t convert_to,
                            Body&& body) {
    __shared__ uint32_t lists[THREADS / 32][64];
    const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    uint32_t* list = lists[warp];
    const uint32_t warps = gridDim.x * (THREADS / 32);
    const uint32_t chunks = (p.slots + 31u) / 32u;
    uint32_t n = 0;
    // The stage bytes of the next chunk are requested before the current chunk is worked on, so
    // the sweep does not wait a memory latency per chunk.
    auto load_stage = [&](uint32_t chunk) -> uint8_t {
        const uint32_t slot = chunk * 32u + lane;
        return (chunk < chunks && slot < p.slots) ? p.b.stage[slot] : (uint8_t)STAGE_IDLE;
    };
    uint32_t chunk = blockIdx.x * (THREADS / 32) + warp;
    uint8_t st_next = load_stage(chunk);
    for (;;) {
        const bool more = chunk < chunks;
        if (more) {
            const uint32_t slot = chunk * 32u + lane;
            const uint8_t st = st_next;
            chunk += warps;
            st_next = load_stage(chunk);
            if (convert_from && st == convert_from) p.b.stage[slot] = convert_to;
            const bool need = st == want;
            const unsigned m = __ballot_sync(0xFFFFFFFFu, need);
            if (need) list[n + __popc(m & ((1u << lane) - 1u))] = slot;
            n += __popc(m);
            __syncwarp();
        }
        if (n >= 32u || (!more && n > 0u)) {
            const uint32_t take = n < 32u ? n : 32u;
            body(lane < take ? list[lane] : 0u, lane < take);
            n -= take;
            const uint32_t carry = lane < n ? list[32u + lane] : 0u;
            __syncwarp();
            if (lane < n) list[lane] = carry;
            __syncwarp();
        } else if (!more) {
            break;
        }
    }
}

// Slots whose ray hit something: shade, emit the next rays.
__global__ void __launch_bounds__(kShadeThreads, RT_WAVE_SHADE_MIN_BLOCKS) wave_shade_kernel(const WaveParams p) {
    const uint32_t P = p.wave & 1u, Q = P ^ 1u;
    WaveCtrl* c = p.b.ctrl;
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        c->ray_n[Q] = 0u;  // next wave's ray queue
        c->ray_head[Q] = 0u;
        c->walk_n = 0u;
    }
    uint32_t* rayq = p.b.rayq + (size_t)P * 2u * p.slots;
    uint32_t n_rays = 0, n_shadow = 0;
    sweep_stage<kShadeThreads>(p, STAGE_SHADE, 0, 0, [&](uint32_t slot, bool on) {
        bool emit_closest = false, emit_shadow = false;
        if (on) {
            PathRegs s;
            uint4 I;
            load_path(p, slot, s, I);
            settle_shadow(p, slot, s);
            float4 o4, d4;
            ld8w<RT_L2_HINT_RAY>(&p.b.rOD[2u * (size_t)slot], o4, d4);
            const float4 h4 = lds(&p.b.hit[slot]);
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
                sts(&p.b.sray_d[slot], make_float4(so.shadow_dir.x, so.shadow_dir.y, so.shadow_dir.z, 0.0f));
                p.b.occl[slot] = 0;
                n_shadow++;
            }
            // Both rays start at the hit point (no offset: shader.wgsl:1249, 1295).
            const float4 point4 = make_float4(so.point.x, so.point.y, so.point.z, RT_INFINITY);
            if (done) {
                // With a shadow ray in flight the sample is finished one wave later.
                I.w = emit_shadow ? SLOT_ENDING : SLOT_DONE;
                p.b.stage[slot] = emit_shadow ? STAGE_FINISH_NEXT : STAGE_FINISH;
                sts(&p.b.rOD[2u * (size_t)slot + 0u], point4);
            } else {
                emit_closest = true;
                st8<RT_L2_HINT_RAY>(&p.b.rOD[2u * (size_t)slot], point4, make_float4(so.next_dir.x, so.next_dir.y, so.next_dir.z, u_as_f(kNone)));
                p.b.stage[slot] = STAGE_FLIGHT;
                n_rays++;
            }
            store_path(p, slot, s, I);
        }
        list_push2(rayq, &c->ray_n[P], emit_closest, slot, emit_shadow, slot | kShadowBit);
    });
    fast_flush_counts(0u, n_rays, n_shadow, p.counters);
}

// 1: the finish kernel runs the analytic pre-pass of the camera rays it creates.
#ifndef RT_FINISH_PREPASS
#define RT_FINISH_PREPASS 1
#endif
#ifndef RT_WAVE_FINISH_MIN_BLOCKS
#define RT_WAVE_FINISH_MIN_BLOCKS 8
#endif

// Slots whose ray escaped (sky light), whose path ended, or that are new: park the finished
// sample and start the next path of the batch.
__global__ void __launch_bounds__(kFinishThreads, RT_WAVE_FINISH_MIN_BLOCKS) wave_finish_kernel(const WaveParams p) {
    const uint32_t P = p.wave & 1u;
    WaveCtrl* c = p.b.ctrl;
#if !RT_FINISH_PREPASS
    uint32_t* rayq = p.b.rayq + (size_t)P * 2u * p.slots;
#endif
    const float sin_half_fov = rt_sin(p.camera.fov_y / 2.0f);
    const size_t pixels = (size_t)p.width * p.height;
    uint32_t n_paths = 0, n_rays = 0;
#if RT_FINISH_PREPASS
    // The camera ray's analytic pre-pass happens right here, while the ray is in registers (this
    // kernel mostly waits on memory, so the arithmetic is nearly free, and the pre-pass kernel
    // has a quarter fewer rays to fetch). Rays that go on to the triangle walk are appended to
    // the walk queue directly, 96-128 at a time per warp.
    __shared__ float4 spre[kPreMaxFloat4];
    stage_pre(p.scene, spre);
    __shared__ uint32_t keep[kFinishThreads / 32][128];
    uint32_t* mine = keep[threadIdx.x >> 5];
    uint32_t* walkq = p.b.rayq + (size_t)(P ^ 1u) * 2u * p.slots;
    uint32_t kept = 0u;
    auto flush = [&]() {
        const uint32_t lane = threadIdx.x & 31u;
        uint32_t base = 0u;
        if (lane == 0u) base = atomicAdd(&c->walk_n, kept);
        base = __shfl_sync(0xFFFFFFFFu, base, 0);
        for (uint32_t k = lane; k < kept; k += 32u) walkq[base + k] = mine[k];
        __syncwarp();
        kept = 0u;
    };
#endif
    sweep_stage<kFinishThreads>(p, STAGE_FINISH, STAGE_FINISH_NEXT, STAGE_FINISH, [&](uint32_t slot, bool on) {
        bool emit = false, retire = false;
        uint4 I = make_uint4(0u, 0u, 0u, SLOT_RETIRED);
        PathRegs s;
        s.rng = 0u;
        float4 d4 = make_float4(0.f, 0.f, 0.f, 0.f);
        if (on) {
            // Everything the slot may need is requested at once (no load waits on another);
            // a new slot's record is uninitialised and ignored.
            load_path(p, slot, s, I);
            d4 = lds(&p.b.rOD[2u * (size_t)slot + 1u]);
        }
        // ---- hand out the next path indices: one atomic per warp, in flight during the sky lookup ----
        uint32_t q = kNone;
        {
            const unsigned m = __ballot_sync(0xFFFFFFFFu, on);
            if (m) {
                const uint32_t lane = threadIdx.x & 31u;
                const int leader = __ffs(m) - 1;
                uint32_t base = 0;
                if ((int)lane == leader) base = atomicAdd(p.next_path, (uint32_t)__popc(m));
                base = __shfl_sync(0xFFFFFFFFu, base, leader);
                if (on) q = base + __popc(m & ((1u << lane) - 1u));
            }
        }
        if (on && I.w != SLOT_NEW) {
            settle_shadow(p, slot, s);
            if (I.w == SLOT_TRACED) {
                // shader.wgsl:1222-1231 — the ray escaped into the environment.
                const V3 d = xyz(d4);
                float u, v;
                direction_to_uv(d, u, v);
                V3 sky = env_bilinear(p.env, u, v);
                float pdf = env_direction_pdf(p.env, u, v);
                float w = power_heuristic(s.last_pdf, pdf);
                s.L = s.L + s.T * sky * w;
            }
            // Park the sample; wave_accumulate_kernel adds the planes in sample order.
            p.staging[(size_t)(I.y - p.batch_begin) * pixels + I.x] = make_float4(s.L.x, s.L.y, s.L.z, 0.0f);
        }
        if (on) {
            uint32_t px = 0, py = 0;
            // ids of partial tiles name no pixel: draw again (only for sizes not divisible by 8)
            while (q < p.path_total && !pixel_of_id(p, q % p.pixel_ids, px, py)) q = atomicAdd(p.next_path, 1u);
            if (q >= p.path_total) {
                I.w = SLOT_RETIRED;
                p.b.stage[slot] = STAGE_IDLE;
                retire = true;
                sts((uint4*)&p.b.pTLPI[4u * (size_t)slot + 3u], I);
            } else {
                I.x = py * p.width + px;
                I.y = p.batch_begin + q / p.pixel_ids;
                V3 o, d;
                camera_ray(p.camera, sin_half_fov, px, py, p.width, p.height, I.y, s.rng, o, d);
                s.T = v3s(1.f);
                s.L = v3s(0.f);
                s.pending = v3s(0.f);
                s.last_pdf = 1.0f;
                s.flags = 0u;
                I.z = 0u;
                I.w = SLOT_TRACED;
#if RT_FINISH_PREPASS
                const RayCtx r = ray_setup(o, d);
                FHit best;
                Fallback fb;
                prepass(p.scene, spre, r, best, fb);
                float t0_root;
                const bool root_ok = root_test(p.scene, r, t0_root);
                Walk w;
                walk_begin(p.scene, root_ok, t0_root, best, w);
                emit = w.cur != kNone;  // triangles to test: on to the walk kernel
                st8<RT_L2_HINT_RAY>(&p.b.rOD[2u * (size_t)slot], make_float4(o.x, o.y, o.z, emit ? best.t : RT_INFINITY),
                                    make_float4(d.x, d.y, d.z, u_as_f(emit ? best.rec : kNone)));
                if (emit) {
                    sts(&p.b.fb[slot], make_float2(fb.t, u_as_f(fb.index)));
                    p.b.stage[slot] = STAGE_FLIGHT;
                } else {
                    resolve_fallback(p.scene, r, fb, best);
                    sts(&p.b.hit[slot], make_float4(best.t, u_as_f(best.rec), best.u, best.v));
                    p.b.stage[slot] = best.rec != kNone ? STAGE_SHADE : STAGE_FINISH;
                }
#else
                st8<RT_L2_HINT_RAY>(&p.b.rOD[2u * (size_t)slot], make_float4(o.x, o.y, o.z, RT_INFINITY), make_float4(d.x, d.y, d.z, u_as_f(kNone)));
                p.b.stage[slot] = STAGE_FLIGHT;
                emit = true;
#endif
                store_path(p, slot, s, I);
                n_paths++;
                n_rays++;
            }
        }
#if RT_FINISH_PREPASS
        {
            const uint32_t lane = threadIdx.x & 31u;
            const unsigned m = __ballot_sync(0xFFFFFFFFu, emit);
            if (emit) mine[kept + __popc(m & ((1u << lane) - 1u))] = slot;
            kept += (uint32_t)__popc(m);
            __syncwarp();
            if (kept > 96u) flush();
        }
#else
        list_push(rayq, &c->ray_n[P], emit, slot);
#endif
        const unsigned gone = __ballot_sync(0xFFFFFFFFu, retire);
        if (gone && (threadIdx.x & 31u) == 0u) atomicSub(&c->live, (uint32_t)__popc(gone));
    });
#if RT_FINISH_PREPASS
    if (kept) flush();
#endif
    fast_flush_counts(n_paths, n_rays, 0u, p.counters);
}

// End of a batch: add the staged sample planes to the accumulation buffer in sample order
// (shader.wgsl:1367-1371: total += sample, alpha = 1).
// Last node of the wave loop's graph body (rt_backend.cu, wave_render): the loop goes on while any
// slot of any group still holds or can get a path.
struct WaveLoopCtrl {
    const WaveCtrl* group[4];
    uint32_t groups;
    uint32_t body_kernels;              // kernels one body runs, this one included
    unsigned long long* launches;       // [0] running total of kernels, for rt_launch_count;
                                        // [1] bodies of the current launch (the host zeroes it before a launch);
                                        // [2] set to 1 when a launch was cut off at `max_bodies`
    uint32_t max_bodies;                // far above what the batch can need: a loop that does not end is
                                        // stopped and reported (rt_sync) instead of hanging the device
};
__global__ void wave_continue_kernel(cudaGraphConditionalHandle handle, const WaveLoopCtrl l) {
    uint32_t live = 0u;
    for (uint32_t g = 0; g < l.groups; g++) live += l.group[g]->live;
    l.launches[0] += l.body_kernels;
    if (++l.launches[1] >= l.max_bodies && live != 0u) {
        l.launches[2] = 1ull;
        live = 0u;
    }
    cudaGraphSetConditional(handle, live != 0u ? 1u : 0u);
}

__global__ void wave_accumulate_kernel(const WaveParams p) {
    const size_t pixels = (size_t)p.width * p.height;
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= pixels) return;
    float4 a = p.accum[i];
    for (uint32_t k = 0; k < p.batch_count; k++) {
        const float4 v = __ldcs(&p.staging[(size_t)k * pixels + i]);
        a.x = a.x + v.x;
        a.y = a.y + v.y;
        a.z = a.z + v.z;
    }
    a.w = 1.0f;
    p.accum[i] = a;
}

#ifndef RT_WAVE_PREPASS_MIN_BLOCKS
#define RT_WAVE_PREPASS_MIN_BLOCKS 4
#endif

// Ray queue of this wave: the analytic pre-pass, one ray per thread, every lane busy. A ray
// that the pre-pass settles — a blocked shadow ray, or any ray that misses the box of the
// effective tree (or enters it beyond the closest analytic hit) — is finished here and
// leaves the queue; only rays with triangles to test go on to the walk.
__global__ void __launch_bounds__(kPrepassThreads, RT_WAVE_PREPASS_MIN_BLOCKS) wave_prepass_kernel(const WaveParams p) {
    __shared__ float4 spre[kPreMaxFloat4];
    stage_pre(p.scene, spre);
    const uint32_t P = p.wave & 1u;
    WaveCtrl* c = p.b.ctrl;
    const uint32_t count = c->ray_n[P];
    const uint32_t* rayq = p.b.rayq + (size_t)P * 2u * p.slots;
    uint32_t* walkq = p.b.rayq + (size_t)(P ^ 1u) * 2u * p.slots;  // the idle half
    const uint32_t stride = gridDim.x * kPrepassThreads;
    const uint32_t lane = threadIdx.x & 31u;
    // Rays that go on collect in a per-warp buffer and reach the walk queue 96-128 at a time
    // (one atomic per warp and iteration on a single counter would be the kernel's bottleneck).
    // The walk kernel thus pulls from a queue without holes: a lane that drew a settled ray
    // used to sit out a whole round.
    __shared__ uint32_t keep[kPrepassThreads / 32][128];
    uint32_t* mine = keep[threadIdx.x >> 5];
    uint32_t kept = 0u;
    auto flush = [&]() {
        uint32_t base = 0u;
        if (lane == 0u) base = atomicAdd(&c->walk_n, kept);
        base = __shfl_sync(0xFFFFFFFFu, base, 0);
        for (uint32_t k = lane; k < kept; k += 32u) walkq[base + k] = mine[k];
        __syncwarp();
        kept = 0u;
    };
    // every lane of a warp runs the same number of iterations (warp-wide votes inside)
    for (uint32_t i0 = blockIdx.x * kPrepassThreads + (threadIdx.x & ~31u); i0 < count; i0 += stride) {
        const uint32_t i = i0 + lane;
        const uint32_t e = i < count ? rayq[i] : kDeadEntry;
        bool walk = false;
        if (e != kDeadEntry) {
            const uint32_t slot = e & ~kShadowBit;
            const bool shadow = (e & kShadowBit) != 0u;
            const V3 o = xyz(lds(&p.b.rOD[2u * (size_t)slot + 0u]));
            const V3 d = xyz(shadow ? lds(&p.b.sray_d[slot]) : lds(&p.b.rOD[2u * (size_t)slot + 1u]));
            const RayCtx r = ray_setup(o, d);
            FHit best;
            Fallback fb;
            prepass(p.scene, spre, r, best, fb);
            float t0_root;
            const bool root_ok = root_test(p.scene, r, t0_root);
            Walk w;
            walk_begin(p.scene, root_ok, t0_root, best, w);
            walk = w.cur != kNone && !(shadow && best.rec != kNone);
            if (shadow) {
                if (best.rec != kNone) p.b.occl[slot] = 1;
            } else if (walk) {
                p.b.rOD[2u * (size_t)slot + 0u].w = best.t;
                p.b.rOD[2u * (size_t)slot + 1u].w = u_as_f(best.rec);
                sts(&p.b.fb[slot], make_float2(fb.t, u_as_f(fb.index)));
            } else {
                resolve_fallback(p.scene, r, fb, best);
                sts(&p.b.hit[slot], make_float4(best.t, u_as_f(best.rec), best.u, best.v));
                p.b.stage[slot] = best.rec != kNone ? STAGE_SHADE : STAGE_FINISH;
            }
        }
        const unsigned m = __ballot_sync(0xFFFFFFFFu, walk);
        if (walk) mine[kept + __popc(m & ((1u << lane) - 1u))] = e;
        kept += (uint32_t)__popc(m);
        __syncwarp();
        if (kept > 96u) flush();
    }
    if (kept) flush();
}

#ifndef RT_WAVE_WALK_MIN_BLOCKS
#define RT_WAVE_WALK_MIN_BLOCKS 5
#endif

// The triangle walk of every queued ray. Persistent warps; a lane whose ray is finished
// takes the next queue entry at the next round boundary.
// SHALLOW: the tree is at most kShallowDepth deep, so the traversal stack fits in shared memory
// ([entry][thread]: two data-stage passes per push / pop whatever the lanes' depths, where the
// thread-private stack costs one pass per distinct depth and word — a third of the kernel's
// load/store work in house.toml).
constexpr uint32_t kShallowDepth = 8u;
template <bool SHALLOW>
__global__ void __launch_bounds__(kWalkThreads, RT_WAVE_WALK_MIN_BLOCKS) wave_walk_kernel(const WaveParams p) {
    const uint32_t P = p.wave & 1u;
    WaveCtrl* c = p.b.ctrl;
    const uint32_t total = c->walk_n;
    const uint32_t* rayq = p.b.rayq + (size_t)(P ^ 1u) * 2u * p.slots;  // compacted by the pre-pass kernel
    const uint32_t lane = threadIdx.x & 31u;
    __shared__ uint2 shallow_stack[SHALLOW ? kShallowDepth * kWalkThreads : 1];
    uint2 deep_stack[SHALLOW ? 1 : kStackCap];
    uint2* stack = SHALLOW ? shallow_stack + threadIdx.x : deep_stack;
    constexpr uint32_t kStride = SHALLOW ? (uint32_t)kWalkThreads : 1u;

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
            if ((int)lane == leader) base = atomicAdd(&c->ray_head[P], (uint32_t)__popc(idle));
            base = __shfl_sync(0xFFFFFFFFu, base, leader);
            if (base + (uint32_t)__popc(idle) >= total) drained = true;
            if (!busy) {
                const uint32_t idx = base + __popc(idle & ((1u << lane) - 1u));
                const uint32_t e = idx < total ? rayq[idx] : kDeadEntry;
                if (e != kDeadEntry) {
                    slot = e & ~kShadowBit;
                    any_hit = (e & kShadowBit) != 0u;
                    const float4 o4 = lds(&p.b.rOD[2u * (size_t)slot + 0u]);
                    const float4 d4 = any_hit ? lds(&p.b.sray_d[slot]) : lds(&p.b.rOD[2u * (size_t)slot + 1u]);
                    r = ray_setup(xyz(o4), xyz(d4));
                    best.t = any_hit ? RT_INFINITY : o4.w;
                    best.rec = any_hit ? kNone : f_as_u(d4.w);
                    best.u = best.v = 0.0f;
                    walk_begin(p.scene, true, 0.0f, best, w);  // the pre-pass kernel checked the root box
                    busy = true;
                }
            }
        }
        if (!__any_sync(0xFFFFFFFFu, busy)) {
            if (drained) break;
            continue;  // only dead entries came back: fetch again
        }

        // ---- one round of the walk: interior records, then the leaf ----
        // Rays with an infinite 1/d component are rare (axis-aligned directions): when the warp
        // holds none, the box tests skip the NaN-safe form without a branch.
        if (__all_sync(0xFFFFFFFFu, !busy || r.plain)) {
            while (busy && w.cur < kLeafBit) walk_node<true>(p.scene, r, w, stack, kStride);
        } else {
            while (busy && w.cur < kLeafBit) walk_node<false>(p.scene, r, w, stack, kStride);
        }
        bool settled = false;
        if (busy && w.cur != kNone) settled = walk_leaf(p.scene, r, any_hit, best, w, stack, kStride);

        // ---- lanes whose ray is finished hand over their result ----
        const bool finished = busy && (settled || w.cur == kNone);
        if (finished) {
            if (any_hit) {
                if (best.rec != kNone) p.b.occl[slot] = 1;
            } else {
                if (best.rec == kNone) {
                    const float2 f = lds(&p.b.fb[slot]);
                    Fallback fb;
                    fb.t = f.x;
                    fb.index = f_as_u(f.y);
                    resolve_fallback(p.scene, r, fb, best);
                }
                sts(&p.b.hit[slot], make_float4(best.t, u_as_f(best.rec), best.u, best.v));
                p.b.stage[slot] = best.rec != kNone ? STAGE_SHADE : STAGE_FINISH;
            }
            busy = false;
        }
    }
}

}  // namespace rtw
