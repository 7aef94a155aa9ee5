// rt_fast.cuh — the traversal and shading building blocks every render kernel shares (two-children-
// per-fetch BVH records, the hoisted analytic pre-pass, the tie rule, the flat leaf-mask walk, hit
// reconstruction, shade_hit), plus two kernels that are NOT the render path: `render_fast_kernel`
// (RT_KERNEL=fast), a thread-per-pixel megakernel over the same code kept as a cross-check of the
// wavefront kernels, and `first_hit_fast_kernel`, the primary-ray parity hook behind rt_first_hit.
// The render path is rt_wave.cuh (four dense kernels per wave inside one CUDA-graph loop);
// rt_local.cuh (RT_KERNEL=local) is the persistent single-launch alternative, kept tested.
//
// Semantics follow /root/reference/src/shaders/shader.wgsl (cast_ray_bvh 469-564,
// cast_ray 567-601, trace_ray 1213-1303, main 1305-1373); the arithmetic is
// rt_device.cuh's (one IEEE binary32 operation per WGSL operator, no contraction). What
// is this backend's own is how the work is organised:
//
//  * Node layout "N2": one 64-byte record per INTERIOR node holding the boxes of both
//    children, so one fetch decides both subtrees, children whose box is missed are never
//    pushed, and leaves are entered without a box test of their own. The far child is
//    pushed together with its entry distance and re-checked against the (shrunken)
//    pruning limit when popped — the same decision the reference order takes when it
//    tests that box later, so the set of primitives tested is unchanged.
//  * Analytic primitives (spheres, planes) are hoisted out of the divergent BVH walk when
//    a scene has few of them: every ray first runs the same straight-line "pre-pass" —
//    each analytic primitive's test plus the box tests of the BVH path that leads to its
//    leaf (the reference finds a primitive iff every box on that path is hit, because its
//    traversal does not prune). The walk that follows sees triangles only and starts with
//    the pruning limit already tightened. The brute-force fallback of cast_ray
//    (shader.wgsl:583-598) falls out of the same tests at no extra cost.
//  * Ties (equal distance, bit for bit) are resolved as the reference's strict `<` in
//    visiting order does: lower slot first inside a leaf; across leaves by the sign of
//    1/d on the split axis of the lowest common ancestor (`precedes`).
#pragma once

#include "rt_device.cuh"

namespace rtf {
using namespace rtd;

constexpr uint32_t kNone = 0xFFFFFFFFu;
constexpr uint32_t kLeafBit = 0x80000000u;
constexpr uint32_t kLeafLenShift = 26u;
constexpr uint32_t kLeafSlotMask = 0x03FFFFFFu;
constexpr uint32_t kLeafLenEscape = 31u;   // real length in big_len[first_slot]
constexpr uint32_t kRecAnalytic = 0x40000000u;  // Hit.rec: analytic primitive reached through the BVH
constexpr uint32_t kRecFallback = 0x60000000u;  // Hit.rec: analytic primitive found by the brute-force fallback
constexpr uint32_t kRecIndexMask = 0x1FFFFFFFu;
constexpr int kStackCap = 64;  // the reference's own stack size (shader.wgsl:482); upload rejects deeper trees

// Pre-pass program limits (hoisting is disabled for scenes beyond them).
constexpr uint32_t kPreMaxLeaves = 16, kPreMaxBoxes = 64, kPreMaxPrims = 32;
constexpr uint32_t kTriStride = 3u;  // float4 per primitive slot record
constexpr uint32_t kPreMaxFloat4 = kPreMaxLeaves + 2 * kPreMaxBoxes + 4 * kPreMaxPrims;
// Flat walk program (small effective trees): leaf boxes, then per-octant visiting orders.
constexpr uint32_t kFlatMaxLeaves = 32;
constexpr uint32_t kFlatMaxSlots = 192;  // leaf slots of a flat scene (its triangle records fit shared memory)
constexpr uint32_t kFlatOrderAt = 2 * kFlatMaxLeaves;            // float4 index of order[8][32] (bytes)
constexpr uint32_t kFlatRankAt = kFlatOrderAt + 8 * 32 / 16;      // float4 index of rank[8][32] (bytes)
constexpr uint32_t kFlatFloat4 = kFlatRankAt + 8 * 32 / 16;

// Device view of a scene in the fast layout.
//   n2:   4 x float4 per interior node of the EFFECTIVE tree: (c1.min, ref1) (c1.max, ref2)
//         (c2.min, -) (c2.max, axis); ref = interior record index or kLeafBit | len << 26 |
//         first_slot. The effective tree is the reference's tree without the leaves whose
//         primitives are all hoisted: a node left with one child is spliced out and every
//         interior box is the union of the LEAF boxes below it (the reference's own leaf
//         boxes). A leaf box lies inside each of its ancestors' boxes and the slab test is
//         monotone in the box, so "the leaf's box is hit" already implies that every box
//         above it is hit: tighter interior boxes only skip work, never a primitive. Leaf
//         order under a surviving node is the reference's (same split axis), so ties among
//         triangles resolve by visiting order as before.
//   tri:  3 x float4 per leaf slot: (a, type) (e0, material) (e1, -); for an analytic slot
//         the first vector is (bits analytic index, -, -, type).
//   ana:  4 x float4 per analytic primitive, spheres then planes in reference order:
//         sphere (pos, 0) (radius, radius^2, bits material, bits slot) (-, -, -, bits index) (-)
//         plane  (pos, 1) (normal, bits material) (M row x, bits slot) (M row z, bits index)
//   pre:  the pre-pass program: leaf headers (uint4: box index of the leaf, prim_begin,
//         prim_count, -), then the distinct boxes on the reference's paths to those leaves,
//         parents first (2 x float4, min.w = bits parent box index or kNone), then primitives
//         (copies of `ana` records).
//   flat: the "flat walk" program of a scene whose effective tree has at most kFlatMaxLeaves
//         leaves (flat_leaves > 0; needs hoisting): per leaf its reference box, 2 x float4
//         (min, bits first slot) (max, bits length); then order[neg][i] = the i-th leaf the
//         reference's walk visits for a ray with sign mask neg (shader.wgsl:536 at every
//         interior node), and rank = its inverse, one byte each. A leaf is visited by the
//         reference iff its own box is hit (every box above contains it and the slab test is
//         monotone in the box), so testing the leaf boxes alone and going through the hit ones
//         in that order tests the same triangles in the same order — without a stack.
//   ana_paths: per analytic primitive its root-to-leaf path in the REFERENCE tree, for
//         precedes(): [analytic_count] offsets, then words m | axis << 28 | side << 30, ended
//         by kNone (m = first slot of the node's second child, side = which child holds it).
struct FastScene {
    const float4* __restrict__ n2;
    const float4* __restrict__ tri;
    const float4* __restrict__ tri_normals;
    const float4* __restrict__ ana;
    const float4* __restrict__ materials;
    const uint32_t* __restrict__ big_len;
    const float4* __restrict__ pre;
    const uint32_t* __restrict__ ana_paths;
    const float4* __restrict__ flat;
    float root_min[3], root_max[3];  // box of the effective tree's root
    uint32_t root_ref;
    uint32_t analytic_count;
    uint32_t pre_leaves, pre_boxes, pre_prims;
    uint32_t hoisted;
    uint32_t flat_leaves;  // 0: no flat walk program
    uint32_t flat_slots;   // leaf slots of a flat scene (its `tri` records are staged in shared memory)
    uint32_t no_prune;  // a plane may stick out of its leaf box (scene.rs:203-207 bounds only two of
                        // its corners) and is not hoisted: the walk must not prune, like the reference
};

struct FHit {
    float t;
    uint32_t rec;  // slot | kRecAnalytic+index | kRecFallback+index | kNone
    float u, v;
};

struct RayCtx {
    V3 o, d, inv;
    uint32_t neg;  // bit a set iff inv[a] < 0: the second child is visited first on axis a
    bool plain;    // no component of 1/d is infinite or NaN
};

// ray_intersects_bounds (shader.wgsl:262-293) with the far bound starting at `limit`.
// Returns the entry distance t0 through `t0_out`.
// ASSUME_PLAIN: the caller knows r.plain holds for every lane (no branch on it).
template <bool ASSUME_PLAIN = false>
RT_DEV bool box_test(const RayCtx& r, float4 bmin, float4 bmax, float limit, float& t0_out) {
    float tnx = (bmin.x - r.o.x) * r.inv.x, tfx = (bmax.x - r.o.x) * r.inv.x;
    float tny = (bmin.y - r.o.y) * r.inv.y, tfy = (bmax.y - r.o.y) * r.inv.y;
    float tnz = (bmin.z - r.o.z) * r.inv.z, tfz = (bmax.z - r.o.z) * r.inv.z;
    float t0, t1;
    if (ASSUME_PLAIN || r.plain) {
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
    t0_out = t0;
    return !(t0 > t1);
}

RT_DEV uint32_t ana_slot(float4 r0, float4 r1, float4 r2) {
    return f_as_i(r0.w) == 0 ? f_as_u(r1.w) : f_as_u(r2.w);
}
RT_DEV uint32_t ana_index(float4 r0, float4 r2, float4 r3) {
    return f_as_i(r0.w) == 0 ? f_as_u(r2.w) : f_as_u(r3.w);
}

// Does the reference's traversal visit the analytic primitive `a_index` (in leaf slot
// a_slot) before slot b for a ray with sign mask `neg`? Only consulted when two candidates
// tie bit for bit. Slots are in DFS order, so at every node on a's path the first child owns
// the slots below `m`; the first node that separates the two decides by the sign of 1/d on
// its split axis (shader.wgsl:536).
RT_COLD bool precedes(const uint32_t* __restrict__ paths, uint32_t a_index, uint32_t a_slot, uint32_t b_slot,
                      uint32_t neg) {
    for (uint32_t i = paths[a_index];; i++) {
        const uint32_t e = paths[i];
        if (e == kNone) break;
        const uint32_t m = e & kLeafSlotMask, axis = (e >> 28) & 3u, side = e >> 30;
        const uint32_t b_side = b_slot >= m ? 1u : 0u;
        if (b_side != side) return side == ((neg >> axis) & 1u);
    }
    return a_slot < b_slot;  // same leaf: slot order
}

RT_DEV uint32_t hit_slot(const FastScene& sc, uint32_t rec) {
    if (rec < kRecAnalytic) return rec;
    const float4* p = sc.ana + 4u * (size_t)(rec & kRecIndexMask);
    return ana_slot(ld4(p), ld4(p + 1), ld4(p + 2));
}

// Candidate from the BVH (triangle or analytic primitive in visiting order or from the
// pre-pass): keep it iff the reference's traversal would.
RT_DEV void offer(const FastScene& sc, const RayCtx& r, FHit& best, float t, uint32_t rec, uint32_t slot,
                  float u, float v, bool in_order) {
    bool take = t < best.t;
    if (!take && t == best.t && best.rec != kNone && !in_order) {
        // one of the two (or both) is a hoisted analytic primitive
        if (rec & kRecAnalytic)
            take = precedes(sc.ana_paths, rec & kRecIndexMask, slot, hit_slot(sc, best.rec), r.neg);
        else
            take = !precedes(sc.ana_paths, best.rec & kRecIndexMask, hit_slot(sc, best.rec), slot, r.neg);
    }
    if (take) {
        best.t = t;
        best.rec = rec;
        best.u = u;
        best.v = v;
    }
}

// An analytic primitive met inside a leaf or in the fallback list of a scene that is not
// hoisted (many spheres / planes). Cold for the scenes the reference ships.
// (Takes and returns the current best by value: reference parameters of an out-of-line
// function would pin the caller's best hit to local memory for the whole walk.)
struct BestRec {
    float t;
    uint32_t rec;
};
RT_COLD BestRec test_analytic_cold(const float4* __restrict__ ana, uint32_t index, V3 o, V3 d, float t_best,
                                   uint32_t rec_best, uint32_t rec_tag) {
    const float4* p = ana + 4u * (size_t)index;
    float4 r0 = ld4(p), r1 = ld4(p + 1);
    float t = 0.0f;
    bool hit;
    if (f_as_i(r0.w) == 0) hit = test_sphere(o, d, r0, r1, t);
    else hit = test_plane(o, d, r0, r1, ld4(p + 2), ld4(p + 3), t);
    BestRec out;
    out.t = t_best;
    out.rec = rec_best;
    if (hit && t < t_best) {
        out.t = t;
        out.rec = rec_tag | index;
    }
    return out;
}

RT_DEV RayCtx ray_setup(V3 o, V3 d) {
    RayCtx r;
    r.o = o;
    r.d = d;
    r.inv = v3(1.0f / d.x, 1.0f / d.y, 1.0f / d.z);
    r.plain = fabsf(r.inv.x) < 3.0e38f && fabsf(r.inv.y) < 3.0e38f && fabsf(r.inv.z) < 3.0e38f;
    r.neg = (r.inv.x < 0.0f ? 1u : 0u) | (r.inv.y < 0.0f ? 2u : 0u) | (r.inv.z < 0.0f ? 4u : 0u);
    return r;
}

// Box of the effective tree's root, un-pruned.
RT_DEV bool root_test(const FastScene& sc, const RayCtx& r, float& t0_root) {
    return box_test(r, make_float4(sc.root_min[0], sc.root_min[1], sc.root_min[2], 0.f),
                    make_float4(sc.root_max[0], sc.root_max[1], sc.root_max[2], 0.f), RT_INFINITY, t0_root);
}

// Brute-force fallback winner of cast_ray (shader.wgsl:583-598) as the pre-pass finds it.
struct Fallback {
    float t;
    uint32_t index;  // analytic index, kNone = nothing hit
};

// The pre-pass over the hoisted analytic primitives: straight-line code, the same for every
// ray. `best` receives the closest analytic hit the reference's walk would reach; `fb` the
// winner of the brute-force loop (spheres, then planes, strict <). `spre` is the block's
// shared copy of the pre-pass program.
RT_DEV void prepass(const FastScene& sc, const float4* __restrict__ spre, const RayCtx& r, FHit& best,
                    Fallback& fb) {
    best.t = RT_INFINITY;
    best.rec = kNone;
    best.u = 0.0f;
    best.v = 0.0f;
    fb.t = RT_INFINITY;
    fb.index = kNone;
    if (!sc.hoisted) return;
    const uint32_t box_base = sc.pre_leaves, prim_base = sc.pre_leaves + 2u * sc.pre_boxes;
    // Which boxes would the reference's (un-pruned) walk get through? Parents come first.
    // A leaf's box lies inside every box above it and, without NaN slabs, the slab test is
    // monotone in the box: for a "plain" ray the leaf's own box decides alone.
    unsigned long long reached = 0ull;
    for (uint32_t b = 0; b < sc.pre_boxes; b++) {
        const float4 bmin = spre[box_base + 2u * b], bmax = spre[box_base + 2u * b + 1u];
        const uint32_t parent = f_as_u(bmin.w);
        const bool is_leaf_box = f_as_u(bmax.w) != 0u;
        bool ok = true;
        if (is_leaf_box || !r.plain) {
            float t0;
            ok = box_test(r, bmin, bmax, RT_INFINITY, t0) &&
                 (r.plain || parent == kNone || ((reached >> parent) & 1ull));
        }
        reached |= (unsigned long long)ok << b;
    }
    uint32_t best_leaf = kNone;  // pre-pass leaf the current best came from
    for (uint32_t l = 0; l < sc.pre_leaves; l++) {
        const float4 hdr = spre[l];
        const uint32_t p0 = f_as_u(hdr.y), pn = f_as_u(hdr.z);
        const bool reach = ((reached >> f_as_u(hdr.x)) & 1ull) != 0ull;
        for (uint32_t k = 0; k < pn; k++) {
            const float4* p = spre + prim_base + 4u * (p0 + k);
            const float4 r0 = p[0], r1 = p[1], r2 = p[2], r3 = p[3];
            float t = 0.0f;
            bool hit;
            if (f_as_i(r0.w) == 0) hit = test_sphere(r.o, r.d, r0, r1, t);
            else hit = test_plane(r.o, r.d, r0, r1, r2, r3, t);
            if (hit) {
                const uint32_t index = ana_index(r0, r2, r3);
                if (t < fb.t || (t == fb.t && index < fb.index)) {
                    fb.t = t;
                    fb.index = index;
                }
                // Slot order inside a leaf makes strict < right; across leaves a tie goes
                // through precedes().
                if (reach) {
                    const uint32_t before = best.rec;
                    offer(sc, r, best, t, kRecAnalytic | index, ana_slot(r0, r1, r2), 0.0f, 0.0f, best_leaf == l);
                    if (best.rec != before) best_leaf = l;
                }
            }
        }
    }
}

// State of a walk in progress (so that it can live in a lane across refills).
struct Walk {
    uint32_t cur;  // interior record, leaf ref or kNone
    uint32_t sp;
    float limit;
};

RT_DEV void walk_begin(const FastScene& sc, bool root_ok, float t0_root, const FHit& best, Walk& w) {
    w.limit = sc.no_prune ? RT_INFINITY : fminf(RT_INFINITY, best.t * 1.001f);
    w.cur = (root_ok && !(t0_root > w.limit)) ? sc.root_ref : kNone;
    w.sp = 0u;
}

// `stride`: distance between a lane's consecutive stack entries — 1 for a private array, the
// block size for a shared-memory stack laid out [entry][thread] (conflict-free for any mix of
// depths in the warp).
RT_DEV void walk_pop(Walk& w, const uint2* stack, uint32_t stride = 1u) {
    w.cur = kNone;
    while (w.sp > 0u) {
        const uint2 e = stack[(--w.sp) * stride];
        if (!(u_as_f(e.y) > w.limit)) {
            w.cur = e.x;
            break;
        }
    }
}

// One interior record: both children's boxes, near child next, far child pushed.
template <bool ASSUME_PLAIN = false>
RT_DEV void walk_node(const FastScene& sc, const RayCtx& r, Walk& w, uint2* stack, uint32_t stride = 1u) {
    const float4* q = sc.n2 + 4u * (size_t)w.cur;
    const float4 q0 = ld4(q), q1 = ld4(q + 1), q2 = ld4(q + 2), q3 = ld4(q + 3);
    float t0a, t0b;
    // A missed child is marked by an entry distance no hit can have (t0 <= limit < inf).
    const float kMiss = __builtin_huge_valf();
    if (!box_test<ASSUME_PLAIN>(r, q0, q1, w.limit, t0a)) t0a = kMiss;
    if (!box_test<ASSUME_PLAIN>(r, q2, q3, w.limit, t0b)) t0b = kMiss;
    const bool second_first = ((r.neg >> f_as_u(q3.w)) & 1u) != 0u;  // shader.wgsl:536
    const uint32_t rn = second_first ? f_as_u(q1.w) : f_as_u(q0.w);
    const uint32_t rf = second_first ? f_as_u(q0.w) : f_as_u(q1.w);
    const float tn0 = second_first ? t0b : t0a, tf0 = second_first ? t0a : t0b;
    if (tn0 != kMiss) {
        w.cur = rn;
        if (tf0 != kMiss) stack[(w.sp++) * stride] = make_uint2(rf, f_as_u(tf0));
    } else if (tf0 != kMiss) {
        w.cur = rf;
    } else {
        walk_pop(w, stack, stride);
    }
}

// The primitives of the leaf in w.cur, in slot order (shader.wgsl:493-524); then the next
// node off the stack. Returns true when an any-hit query is settled.
RT_DEV bool walk_leaf(const FastScene& sc, const RayCtx& r, bool any_hit, FHit& best, Walk& w,
                      const uint2* stack, uint32_t stride = 1u) {
    const uint32_t first = w.cur & kLeafSlotMask;
    uint32_t len = (w.cur >> kLeafLenShift) & 31u;
    if (len == kLeafLenEscape) len = sc.big_len[first];
    for (uint32_t slot = first; slot < first + len; slot++) {
        const float4* p = sc.tri + kTriStride * (size_t)slot;
        const float4 r0 = ld4(p), r1 = ld4(p + 1), r2 = ld4(p + 2);
        if (f_as_i(r0.w) == 2) {
            float t = 0.0f, u = 0.0f, v = 0.0f;
            if (test_triangle(r.o, r.d, r0, r1, r2, t, u, v))
                offer(sc, r, best, t, slot, slot, u, v, !sc.hoisted || best.rec < kRecAnalytic);
        } else if (!sc.hoisted) {
            const BestRec b = test_analytic_cold(sc.ana, f_as_u(r0.x), r.o, r.d, best.t, best.rec, kRecAnalytic);
            best.t = b.t;
            best.rec = b.rec;
        }
        if (any_hit && best.rec != kNone) return true;
    }
    if (!sc.no_prune) w.limit = fminf(RT_INFINITY, best.t * 1.001f);
    walk_pop(w, stack, stride);
    return false;
}

// cast_ray's brute-force fallback: only when the BVH found nothing (shader.wgsl:576-598).
RT_DEV void resolve_fallback(const FastScene& sc, const RayCtx& r, const Fallback& fb, FHit& best) {
    if (best.rec != kNone) return;
    if (sc.hoisted) {
        if (fb.index != kNone) {
            best.t = fb.t;
            best.rec = kRecFallback | fb.index;
        }
    } else {
        for (uint32_t i = 0; i < sc.analytic_count; i++)
        {
            const BestRec b = test_analytic_cold(sc.ana, i, r.o, r.d, best.t, best.rec, kRecFallback);
            best.t = b.t;
            best.rec = b.rec;
        }
    }
}

// cast_ray (any_hit == false) or the occlusion query of a shadow ray (any_hit == true;
// the shader calls cast_ray_bvh and reads did_hit only). `spre` is the block's shared
// copy of the pre-pass program; `stack` is kStackCap entries of (ref, entry distance).
RT_DEV void trace_fast(const FastScene& sc, const float4* __restrict__ spre, V3 o, V3 d, bool any_hit,
                       FHit& best, uint2* stack) {
    const RayCtx r = ray_setup(o, d);
    Fallback fb;
    prepass(sc, spre, r, best, fb);
    if (any_hit && best.rec != kNone) return;
    float t0_root;
    const bool root_ok = root_test(sc, r, t0_root);

    // the walk ("while-while"): interior records until this lane holds a leaf, then the leaf
    Walk w;
    walk_begin(sc, root_ok, t0_root, best, w);
    for (;;) {
        while (w.cur < kLeafBit) walk_node(sc, r, w, stack);
        if (w.cur == kNone) break;
        if (walk_leaf(sc, r, any_hit, best, w, stack)) return;
    }
    if (!any_hit) resolve_fallback(sc, r, fb, best);
}

// ---- flat walk (FastScene::flat) ----
RT_DEV const uint8_t* flat_order(const float4* __restrict__ sflat, uint32_t neg) {
    return (const uint8_t*)(sflat + kFlatOrderAt) + 32u * neg;
}
RT_DEV const uint8_t* flat_rank(const float4* __restrict__ sflat, uint32_t neg) {
    return (const uint8_t*)(sflat + kFlatRankAt) + 32u * neg;
}

// Bit i is set iff the i-th leaf in the ray's visiting order has its box hit within `limit`.
// The loop runs over the leaves in storage order (every lane reads the same box) and files
// each answer under the leaf's rank for this ray's octant.
RT_DEV uint32_t flat_mask(const FastScene& sc, const float4* __restrict__ sflat, const RayCtx& r, float limit) {
    const uint8_t* rank = flat_rank(sflat, r.neg);
    uint32_t mask = 0u;
    for (uint32_t leaf = 0; leaf < sc.flat_leaves; leaf++) {
        float t0;
        const bool ok = box_test(r, sflat[2u * leaf], sflat[2u * leaf + 1u], limit, t0);
        mask |= (ok ? 1u : 0u) << rank[leaf];
    }
    return mask;
}

RT_DEV uint32_t lowest_bit(uint32_t m) {
#ifdef __CUDA_ARCH__
    return (uint32_t)__ffs((int)m) - 1u;
#else
    return (uint32_t)__builtin_ctz(m);
#endif
}

// cast_ray / the occlusion query like trace_fast, over the flat walk program. `sflat` is
// the (shared-memory) copy of FastScene::flat.
RT_DEV void trace_flat(const FastScene& sc, const float4* __restrict__ spre, const float4* __restrict__ sflat,
                       V3 o, V3 d, bool any_hit, FHit& best) {
    const RayCtx r = ray_setup(o, d);
    Fallback fb;
    prepass(sc, spre, r, best, fb);
    if (any_hit && best.rec != kNone) return;
    float limit = fminf(RT_INFINITY, best.t * 1.001f);
    uint32_t mask = flat_mask(sc, sflat, r, limit);
    const uint8_t* order = flat_order(sflat, r.neg);
    while (mask) {
        const uint32_t leaf = order[lowest_bit(mask)];
        mask &= mask - 1u;
        const float4 bmin = sflat[2u * leaf], bmax = sflat[2u * leaf + 1u];
        float t0;
        if (!box_test(r, bmin, bmax, limit, t0)) continue;  // the limit has shrunk since the mask was taken
        const uint32_t first = f_as_u(bmin.w), len = f_as_u(bmax.w);
        for (uint32_t slot = first; slot < first + len; slot++) {
            const float4* p = sc.tri + kTriStride * (size_t)slot;
            const float4 r0 = ld4(p), r1 = ld4(p + 1), r2 = ld4(p + 2);
            if (f_as_i(r0.w) != 2) continue;  // a hoisted sphere / plane sharing the leaf
            float t = 0.0f, u = 0.0f, v = 0.0f;
            if (test_triangle(r.o, r.d, r0, r1, r2, t, u, v))
                offer(sc, r, best, t, slot, slot, u, v, best.rec < kRecAnalytic);
            if (any_hit && best.rec != kNone) return;
        }
        limit = fminf(RT_INFINITY, best.t * 1.001f);
    }
    if (!any_hit) resolve_fallback(sc, r, fb, best);
}

// Deferred hit reconstruction: hit_point, shading normal, material row.
RT_DEV void reconstruct_fast(const FastScene& sc, V3 o, V3 d, const FHit& h, V3& point, V3& normal,
                             uint32_t& material) {
    point = o + d * h.t;
    if (h.rec < kRecAnalytic) {
        // shader.wgsl:446-457
        const float4* np = sc.tri_normals + 3u * (size_t)h.rec;
        V3 n0 = xyz(ld4(np)), n1 = xyz(ld4(np + 1)), n2 = xyz(ld4(np + 2));
        V3 n = normalize(((1.0f - h.u - h.v) * n0 + h.u * n1) + h.v * n2);
        if (dot(n, d) > 0.0f) n = n * -1.0f;
        normal = n;
        material = f_as_u(ld4(sc.tri + kTriStride * (size_t)h.rec + 1).w);
        return;
    }
    const float4* p = sc.ana + 4u * (size_t)(h.rec & kRecIndexMask);
    float4 r0 = ld4(p), r1 = ld4(p + 1);
    if (f_as_i(r0.w) == 0) {
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

struct FastParams {
    FastScene scene;
    EnvView env;
    CameraParams camera;
    float4* accum;
    unsigned long long* counters;
    uint32_t* slot_out;  // first-hit kernel only
    float* t_out;        // first-hit kernel only
    uint32_t width, height;
    uint32_t sample_begin, sample_count;
    uint32_t max_bounces;
    uint32_t sphere_count;
    uint32_t row_begin;  // first-hit kernel only: first image row of this launch's band
};

constexpr int kFastTileW = 16, kFastTileH = 8, kFastThreads = kFastTileW * kFastTileH;

RT_DEVONLY void fast_thread_pixel(uint32_t& px, uint32_t& py) {
    uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    px = blockIdx.x * kFastTileW + (warp & 1u) * 8u + (lane & 7u);
    py = blockIdx.y * kFastTileH + (warp >> 1) * 4u + (lane >> 3);
}

RT_DEVONLY void stage_pre(const FastScene& sc, float4* spre) {
    const uint32_t total = sc.hoisted ? sc.pre_leaves + 2u * sc.pre_boxes + 4u * sc.pre_prims : 0u;
    for (uint32_t i = threadIdx.x; i < total; i += blockDim.x) spre[i] = sc.pre[i];
    __syncthreads();
}

RT_DEVONLY void fast_flush_counts(uint32_t paths, uint32_t rays, uint32_t shadow, unsigned long long* out) {
    uint32_t vals[3] = {paths, rays, shadow};
    for (int i = 0; i < 3; i++) {
        uint32_t v = vals[i];
        for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(0xFFFFFFFFu, v, off);
        if ((threadIdx.x & 31u) == 0u && v) atomicAdd(out + i, (unsigned long long)v);
    }
}

#ifndef RT_FAST_MIN_BLOCKS
#define RT_FAST_MIN_BLOCKS 5
#endif


// What a shaded hit hands back to the path loop.
struct ShadeOut {
    V3 point;        // hit point: origin of the shadow ray and of the next bounce (no offset: shader.wgsl:1249, 1295)
    V3 shadow_dir;   // next-event direction (valid when want_shadow)
    V3 pending;      // light the next-event sample adds if its shadow ray is free
    V3 next_dir;     // BSDF-sampled direction (valid when !done)
    float next_pdf;
    bool want_shadow, done;
};

// shader.wgsl:1233-1298 for one closest hit: emission, next-event estimation set-up,
// bsdf_sample and the throughput update. L and T are updated in place.
RT_DEVONLY void shade_hit(const FastScene& sc, const EnvView& env, V3 o, V3 d, const FHit& hit, uint32_t& rng,
                             V3& T, V3& L, ShadeOut& out) {
    V3 n;
    uint32_t mat_id;
    reconstruct_fast(sc, o, d, hit, out.point, n, mat_id);
    SceneView mv;
    mv.materials = sc.materials;
    const Material m = load_material(mv, mat_id);
    L = L + T * m.emission;  // shader.wgsl:1236

    const V3 wo_world = -d;
    const Frame frame = make_frame(n);
    const V3 wo = to_local(frame, wo_world);

    // ---- next-event estimation: sample_environment (shader.wgsl:782-820) ----
    out.want_shadow = false;
    out.shadow_dir = v3s(0.f);
    out.pending = v3s(0.f);
    {
        uint32_t length = env.width * env.height;
        uint32_t index = min((uint32_t)(rng_uniform(rng) * (float)length), length - 1u);
        float4 entry = ld4(env.alias + index);
        uint32_t pick = rng_uniform(rng) < entry.x ? index : f_as_u(entry.y);
        uint32_t ex = pick % env.width, ey = pick / env.width;
        float jx = rng_uniform(rng), jy = rng_uniform(rng);
        float eu = ((float)ex + jx) / (float)env.width;
        float ev = ((float)ey + jy) / (float)env.height;
        float phi = (2.0f * eu - 1.0f) * RT_PI;
        float theta = RT_PI * ev;
        float st, ct, sph, cph;
        rt_sincos(theta, &st, &ct);
        rt_sincos(phi, &sph, &cph);
        V3 wi_world = v3(st * cph, ct, st * sph);
        float pmf = pick == index ? entry.z : ld4(env.alias + pick).z;
        float d_phi = RT_TWO_PI / (float)env.width;
        float d_theta = RT_PI / (float)env.height;
        float pdf_env = pmf / (d_phi * d_theta * fmaxf(1.0e-6f, st));
        float cos_theta = fmaxf(0.0f, dot(n, wi_world));
        if (cos_theta > 0.0f && pdf_env > 0.0f) {
            // The contribution is computed now and added only if the shadow ray is free.
            out.want_shadow = true;
            out.shadow_dir = wi_world;
            V3 radiance = env_bilinear(env, eu, ev);
            V3 wi = to_local(frame, wi_world);
            V3 f;
            float pdf_bsdf;
            bsdf_eval_pdf(wo, wi, m, f, pdf_bsdf);
            float w = power_heuristic(pdf_env, pdf_bsdf);
            out.pending = T * w * radiance * f * cos_theta / pdf_env;
        }
    }

    // ---- bsdf_sample (shader.wgsl:1116-1202) + path update (1270-1298) ----
    out.done = false;
    out.next_dir = v3s(0.f);
    out.next_pdf = 0.0f;
    if (dot(n, wo_world) <= 0.0f) {
        // The debug colour replaces everything gathered so far (shader.wgsl:1271-1276);
        // the shadow ray is still traced (the shader traces it first) but adds nothing.
        L = v3(0.f, 0.f, 1.f);
        out.pending = v3s(0.f);
        out.done = true;
    } else if (wo.z <= 0.0f) {
        L = v3(0.f, 1.f, 0.f);
        out.pending = v3s(0.f);
        out.done = true;
    } else {
        float sample_u = rng_uniform(rng);
        V3 wi;
        bool blocked = false;
        if (sample_u < m.p_diff) {
            float sx = sample_u / fmaxf(m.p_diff, 1.e-6f);
            float sy = rng_uniform(rng);
            float rr = sqrtf(sx);
            float s, c;
            rt_sincos(RT_TWO_PI * sy, &s, &c);
            float x = rr * c, y = rr * s;
            wi = v3(x, y, sqrtf(fmaxf(0.0f, 1.0f - x * x - y * y)));
        } else {
            float sx = (sample_u - m.p_diff) / fmaxf(m.p_spec, 1.e-6f);
            float sy = rng_uniform(rng);
            V3 h = sample_ggx_vndf(sx, sy, wo, m.alpha);
            wi = reflect(-wo, h);
            blocked = wi.z <= 0.0f;
        }
        if (blocked) {
            out.done = true;  // direction (1,0,0) with pdf 0: the path ends, L kept
        } else {
            V3 f;
            float pdf;
            bsdf_eval_pdf(wo, wi, m, f, pdf);
            V3 wi_world = to_world(frame, wi);
            if (dot(n, wi_world) < 0.0f) {
                L = v3(0.f, 1.f, 0.f);
                out.pending = v3s(0.f);
                out.done = true;
            } else if (pdf <= 0.0f) {
                out.done = true;
            } else {
                float cos_theta = fmaxf(0.0f, dot(n, wi_world));
                T = T * (f * (cos_theta / pdf));
                if (length(T) < 0.001f) {
                    out.done = true;
                } else {
                    out.next_pdf = pdf;
                    out.next_dir = wi_world;
                }
            }
        }
    }
}

// One thread owns one pixel for all of its samples (accumulator in registers, one 16-byte
// load and store per pixel per launch). Every bounce is two passes over ONE copy of the
// traversal code: pass 0 casts the lanes' closest-hit rays and shades the hits together,
// pass 1 casts the shadow rays those hits asked for. A lane whose path ends starts its
// pixel's next sample at once, so warps stay populated although path lengths differ.
__global__ void __launch_bounds__(kFastThreads, RT_FAST_MIN_BLOCKS) render_fast_kernel(const FastParams p) {
    __shared__ float4 spre[kPreMaxFloat4];
    stage_pre(p.scene, spre);

    uint32_t px, py;
    fast_thread_pixel(px, py);
    const bool active = px < p.width && py < p.height;
    const size_t pixel = (size_t)py * p.width + px;
    uint2 stack[kStackCap];

    const float sin_half_fov = rt_sin(p.camera.fov_y / 2.0f);
    const uint32_t sample_end = p.sample_begin + p.sample_count;
    uint32_t n_paths = 0, n_rays = 0, n_shadow = 0;

    float4 acc4 = active ? p.accum[pixel] : make_float4(0.f, 0.f, 0.f, 0.f);
    V3 acc = v3(acc4.x, acc4.y, acc4.z);
    uint32_t sample = active ? p.sample_begin : sample_end;

    uint32_t rng = 0u, bounce = 0u;
    V3 o = v3s(0.f), d = v3s(0.f), T = v3s(1.f), L = v3s(0.f);
    V3 pending = v3s(0.f), next_dir = v3s(0.f);
    float last_pdf = 1.0f;
    bool fresh = true;  // the lane needs a camera ray

    while (sample < sample_end) {
        const bool alive = sample < sample_end;
        if (alive && fresh) {
            camera_ray(p.camera, sin_half_fov, px, py, p.width, p.height, sample, rng, o, d);
            n_paths++;
            T = v3s(1.f);
            L = v3s(0.f);
            last_pdf = 1.0f;
            bounce = 0u;
            fresh = false;
        }
        bool want_shadow = false, finish = false;
#pragma unroll 1
        for (int pass = 0; pass < 2; pass++) {
            const bool go = pass == 0 ? alive : want_shadow;
            FHit hit;
            hit.rec = kNone;
            if (go) trace_fast(p.scene, spre, o, d, pass == 1, hit, stack);
            if (pass == 0) {
                if (alive) {
                    n_rays++;
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
                        finish = so.done || bounce >= p.max_bounces;
                        o = so.point;
                        pending = so.pending;
                        next_dir = so.next_dir;
                        last_pdf = so.next_pdf;
                        want_shadow = so.want_shadow;
                        if (want_shadow) {
                            d = so.shadow_dir;
                            n_shadow++;
                        }
                    }
                }
            } else if (go) {
                // shader.wgsl:1249-1264: the sample counts iff nothing blocks it.
                if (hit.rec == kNone) L = L + pending;
            }
        }
        if (alive) {
            if (finish) {
                acc = acc + L;  // shader.wgsl:1367-1368
                sample++;
                fresh = true;
            } else {
                d = next_dir;
            }
        }
    }

    if (active) p.accum[pixel] = make_float4(acc.x, acc.y, acc.z, 1.0f);  // shader.wgsl:1371
    fast_flush_counts(n_paths, n_rays, n_shadow, p.counters);
}

__global__ void __launch_bounds__(kFastThreads) first_hit_fast_kernel(const FastParams p) {
    __shared__ float4 spre[kPreMaxFloat4];
    stage_pre(p.scene, spre);
    uint32_t px, py;
    fast_thread_pixel(px, py);
    py += p.row_begin;
    uint32_t n = 0;
    if (px < p.width && py < p.height) {
        uint2 stack[kStackCap];
        uint32_t rng;
        V3 o, d;
        camera_ray(p.camera, rt_sin(p.camera.fov_y / 2.0f), px, py, p.width, p.height, p.sample_begin, rng,
                   o, d);
        FHit hit;
        trace_fast(p.scene, spre, o, d, false, hit, stack);
        n = 1;
        uint32_t code = kNone;
        if (hit.rec != kNone) {
            if ((hit.rec & kRecFallback) == kRecFallback) {
                // 0x80000000 | type << 28 | index in the type's own array (include/rt_backend.h)
                uint32_t index = hit.rec & kRecIndexMask;
                code = index < p.sphere_count ? (0x80000000u | index)
                                              : (0x90000000u | (index - p.sphere_count));
            } else {
                code = hit_slot(p.scene, hit.rec);
            }
        }
        size_t i = (size_t)py * p.width + px;
        p.slot_out[i] = code;
        p.t_out[i] = hit.rec != kNone ? hit.t : 0.0f;
    }
    fast_flush_counts(n, n, 0u, p.counters);
}

}  // namespace rtf
