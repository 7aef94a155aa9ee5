// rt_fast_host.hpp — host-side construction of the fast device layout (rt_fast.cuh) from
// the reference-layout scene arrays (include/rt_types.h). Inputs must already be validated
// (rt_upload_scene does that); shared with the CPU emulation used by the tests.
#pragma once

#include <cmath>
#include <cstring>
#include <memory>
#include <thread>
#include <vector>

#include "rt_fast.cuh"
#include "rt_types.h"

namespace rtf {

inline float bits(uint32_t u) {
    float f;
    memcpy(&f, &u, 4);
    return f;
}
inline uint32_t word(float f) {
    uint32_t u;
    memcpy(&u, &f, 4);
    return u;
}

// std::vector whose resize() leaves new elements uninitialised: the big per-slot arrays are written
// in full right after (by several threads), and zero-filling 0.5 GB first costs as much as filling it.
template <typename T>
struct DefaultInitAllocator : std::allocator<T> {
    template <typename U>
    struct rebind {
        using other = DefaultInitAllocator<U>;
    };
    template <typename U, typename... Args>
    void construct(U* p, Args&&... args) {
        if constexpr (sizeof...(Args) == 0) ::new ((void*)p) U;
        else ::new ((void*)p) U(static_cast<Args&&>(args)...);
    }
};
using RawFloat4Vec = std::vector<float4, DefaultInitAllocator<float4>>;

// fn(begin, end) over [0, n) cut into one contiguous range per host thread (large scenes: the
// per-slot and per-node loops of the upload are independent per element).
template <typename F>
inline void parallel_ranges(size_t n, F fn) {
    unsigned threads = std::thread::hardware_concurrency();
    threads = threads < 1u ? 1u : (threads > 32u ? 32u : threads);
    if (threads == 1u || n < ((size_t)1 << 16)) {
        fn((size_t)0, n);
        return;
    }
    std::vector<std::thread> pool;
    for (unsigned t = 0; t < threads; t++) {
        const size_t lo = n * t / threads, hi = n * (t + 1) / threads;
        pool.emplace_back([=] { fn(lo, hi); });
    }
    for (auto& th : pool) th.join();
}

// ---- fast layout (rt_fast.cuh): N2 interior records, 48-byte triangle records, analytic
//      records and the pre-pass program. Inputs are already validated. ----
// Plane::bounds (scene.rs:203-207) is the box of two opposite corners, pos and
// pos + forward + right. It contains the other two corners iff forward and right never
// point opposite ways on any axis. forward / right are recovered from the base-change
// matrix (columns of its inverse); anything doubtful counts as loose, which only costs speed.
inline bool plane_is_loose(const RtPlane& pl) {
    double m[3][3];
    for (int c = 0; c < 3; c++)
        for (int r = 0; r < 3; r++) m[c][r] = pl.base_change[c][r];  // m[c] = column c
    auto cof = [&](int c0, int r0, int c1, int r1) { return m[c0][r0] * m[c1][r1] - m[c0][r1] * m[c1][r0]; };
    const double det = m[0][0] * cof(1, 1, 2, 2) - m[1][0] * cof(0, 1, 2, 2) + m[2][0] * cof(0, 1, 1, 2);
    if (!(det == det) || det == 0.0) return true;
    // inverse column j, row i = cofactor(j, i) / det
    double inv[3][3];
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++) {
            const int c0 = (i + 1) % 3, c1 = (i + 2) % 3, r0 = (j + 1) % 3, r1 = (j + 2) % 3;
            inv[j][i] = cof(c0, r0, c1, r1) / det;  // column j, row i
        }
    const double* right = inv[0];
    const double* forward = inv[2];
    double lr = 0, lf = 0;
    for (int k = 0; k < 3; k++) {
        lr += right[k] * right[k];
        lf += forward[k] * forward[k];
    }
    if (!(lr == lr) || !(lf == lf)) return true;
    const double tr = 1e-6 * std::sqrt(lr), tf = 1e-6 * std::sqrt(lf);
    for (int k = 0; k < 3; k++)
        if (right[k] * forward[k] < 0.0 && std::fabs(right[k]) > tr && std::fabs(forward[k]) > tf) return true;
    return false;
}

inline bool scene_has_loose_plane(const RtSceneDesc& s) {
    for (uint32_t i = 0; i < s.plane_count; i++)
        if (plane_is_loose(s.planes[i])) return true;
    return false;
}

struct FastLayout {
    std::vector<float4> n2, ana, pre, flat;
    RawFloat4Vec tri;
    std::vector<uint32_t> big_len, ana_paths;
    FastScene view{};
};

inline void build_fast_layout(const RtSceneDesc& s, FastLayout& out) {
    const uint32_t nodes = s.bvh_node_count, slots = s.primitive_count;
    const uint32_t analytic = s.sphere_count + s.plane_count;

    std::vector<int64_t> parent(nodes, -1);
    for (uint32_t i = 0; i < nodes; i++) {
        const RtBvhNode& n = s.bvh_nodes[i];
        if (n.primitives_len == 0) {
            parent[i + 1] = i;
            parent[n.primitives_or_second_child_index] = i;
        }
    }
    auto first_slot_of = [&](uint32_t i) {
        while (s.bvh_nodes[i].primitives_len == 0) i = i + 1;
        return s.bvh_nodes[i].primitives_or_second_child_index;
    };
    auto leaf_has_analytic = [&](const RtBvhNode& n) {
        for (uint32_t k = 0; k < n.primitives_len; k++)
            if (s.primitives[n.primitives_or_second_child_index + k].primitive_type != 2u) return true;
        return false;
    };
    auto leaf_has_triangle = [&](const RtBvhNode& n) {
        for (uint32_t k = 0; k < n.primitives_len; k++)
            if (s.primitives[n.primitives_or_second_child_index + k].primitive_type == 2u) return true;
        return false;
    };

    // ---- hoisting decision: few analytic primitives, few distinct boxes on their paths ----
    std::vector<uint32_t> ana_leaves;  // leaves holding at least one analytic primitive, in node (= slot) order
    std::vector<uint32_t> box_nodes;   // distinct reference nodes on the paths to them, parents first
    {
        std::vector<char> on_path(nodes, 0);
        for (uint32_t i = 0; i < nodes; i++) {
            const RtBvhNode& n = s.bvh_nodes[i];
            if (n.primitives_len == 0 || !leaf_has_analytic(n)) continue;
            ana_leaves.push_back(i);
            for (int64_t x = i; x >= 0; x = parent[x]) on_path[(size_t)x] = 1;
        }
        for (uint32_t i = 0; i < nodes && box_nodes.size() <= kPreMaxBoxes; i++)  // node order = parents first
            if (on_path[i]) box_nodes.push_back(i);
    }
    const bool hoisted = analytic <= kPreMaxPrims && ana_leaves.size() <= kPreMaxLeaves &&
                         box_nodes.size() <= kPreMaxBoxes;

    // ---- the effective tree: interior records over the leaves the walk has to test ----
    struct Sub {
        uint32_t ref;  // kNone, leaf ref or interior record index
        float mn[3], mx[3];
    };
    bool any_big = false;
    std::vector<float4>& n2 = out.n2;
    n2.clear();
    n2.reserve((size_t)(nodes / 2 + 1) * 4);  // one record per interior node at most
    // Iterative post-order over the reference tree (depth <= 64, but node counts reach 10^7).
    std::vector<Sub> result(nodes);
    {
        std::vector<uint32_t> order;
        order.reserve(nodes);
        std::vector<uint32_t> todo{0u};
        while (!todo.empty()) {
            uint32_t i = todo.back();
            todo.pop_back();
            order.push_back(i);
            const RtBvhNode& n = s.bvh_nodes[i];
            if (n.primitives_len == 0) {
                todo.push_back(i + 1);
                todo.push_back(n.primitives_or_second_child_index);
            }
        }
        for (size_t k = order.size(); k-- > 0;) {  // children before parents
            const uint32_t i = order[k];
            const RtBvhNode& n = s.bvh_nodes[i];
            Sub& r = result[i];
            if (n.primitives_len != 0) {
                const bool walked = hoisted ? leaf_has_triangle(n) : true;
                if (!walked) {
                    r.ref = kNone;
                    continue;
                }
                uint32_t len = n.primitives_len;
                if (len >= kLeafLenEscape) {
                    any_big = true;
                    len = kLeafLenEscape;
                }
                r.ref = kLeafBit | (len << kLeafLenShift) | n.primitives_or_second_child_index;
                for (int a = 0; a < 3; a++) {
                    r.mn[a] = n.min[a];  // the reference's own leaf box
                    r.mx[a] = n.max[a];
                }
                continue;
            }
            const Sub& c1 = result[i + 1];
            const Sub& c2 = result[n.primitives_or_second_child_index];
            if (c1.ref == kNone && c2.ref == kNone) {
                r.ref = kNone;
            } else if (c1.ref == kNone || c2.ref == kNone) {
                r = c1.ref == kNone ? c2 : c1;  // spliced out
            } else {
                r.ref = (uint32_t)(n2.size() / 4);
                n2.push_back(make_float4(c1.mn[0], c1.mn[1], c1.mn[2], bits(c1.ref)));
                n2.push_back(make_float4(c1.mx[0], c1.mx[1], c1.mx[2], bits(c2.ref)));
                n2.push_back(make_float4(c2.mn[0], c2.mn[1], c2.mn[2], 0.f));
                n2.push_back(make_float4(c2.mx[0], c2.mx[1], c2.mx[2], bits(n.split_axis)));
                for (int a = 0; a < 3; a++) {
                    r.mn[a] = c1.mn[a] < c2.mn[a] ? c1.mn[a] : c2.mn[a];
                    r.mx[a] = c1.mx[a] > c2.mx[a] ? c1.mx[a] : c2.mx[a];
                }
            }
        }
    }
    const Sub root = result[0];
    if (n2.empty()) n2.assign(4, make_float4(0.f, 0.f, 0.f, 0.f));  // never read: the root is a leaf or empty

    // ---- triangle and analytic records ----
    out.ana.assign((size_t)(analytic ? analytic : 1) * 4, make_float4(0.f, 0.f, 0.f, 0.f));
    std::vector<uint32_t> slot_of_analytic(analytic, kNone);
    out.tri.resize((size_t)slots * kTriStride);
    parallel_ranges(slots, [&](size_t lo, size_t hi) {
        const float4 zero = make_float4(0.f, 0.f, 0.f, 0.f);
        for (size_t slot = lo; slot < hi; slot++) {
            const RtPrimitiveInfo& info = s.primitives[slot];
            float4* q = &out.tri[slot * kTriStride];
            if (info.primitive_type == 2u) {
                const RtTriangle& t = s.triangles[info.index];
                const RtVec3 &a = s.vertices[t.vertex_0], &b = s.vertices[t.vertex_1], &c = s.vertices[t.vertex_2];
                // edge_0 = b - a, edge_1 = c - a (shader.wgsl:415-416): single f32 subtractions,
                // so hoisting them to upload time leaves every bit unchanged.
                q[0] = make_float4(a.x, a.y, a.z, bits(2u));
                q[1] = make_float4(b.x - a.x, b.y - a.y, b.z - a.z, bits(t.material_id));
                q[2] = make_float4(c.x - a.x, c.y - a.y, c.z - a.z, 0.f);
            } else {
                const uint32_t index = info.primitive_type == 0u ? info.index : s.sphere_count + info.index;
                slot_of_analytic[index] = (uint32_t)slot;  // an analytic primitive sits in exactly one slot
                q[0] = make_float4(bits(index), 0.f, 0.f, bits(info.primitive_type));
                q[1] = zero;
                q[2] = zero;
            }
        }
    });
    for (uint32_t i = 0; i < s.sphere_count; i++) {
        const RtSphere& sp = s.spheres[i];
        float4* r = &out.ana[(size_t)i * 4];
        r[0] = make_float4(sp.pos[0], sp.pos[1], sp.pos[2], bits(0u));
        r[1] = make_float4(sp.radius, sp.radius * sp.radius, bits(sp.material_id), bits(slot_of_analytic[i]));
        r[2] = make_float4(0.f, 0.f, 0.f, bits(i));
    }
    for (uint32_t i = 0; i < s.plane_count; i++) {
        const RtPlane& pl = s.planes[i];
        const uint32_t index = s.sphere_count + i;
        float4* r = &out.ana[(size_t)index * 4];
        r[0] = make_float4(pl.pos[0], pl.pos[1], pl.pos[2], bits(1u));
        r[1] = make_float4(pl.normal[0], pl.normal[1], pl.normal[2], bits(pl.material_id));
        r[2] = make_float4(pl.base_change[0][0], pl.base_change[1][0], pl.base_change[2][0], bits(slot_of_analytic[index]));
        r[3] = make_float4(pl.base_change[0][2], pl.base_change[1][2], pl.base_change[2][2], bits(index));
    }

    if (any_big) {
        out.big_len.assign(slots, 0u);
        for (uint32_t i = 0; i < nodes; i++)
            if (s.bvh_nodes[i].primitives_len >= kLeafLenEscape)
                out.big_len[s.bvh_nodes[i].primitives_or_second_child_index] = s.bvh_nodes[i].primitives_len;
    }

    // ---- pre-pass program: leaf headers | boxes (parents first) | primitive records ----
    uint32_t pre_leaves = 0, pre_boxes = 0, pre_prims = 0;
    std::vector<float4> headers, boxes, prims;
    out.ana_paths.assign(analytic ? analytic : 1, 0u);
    if (hoisted) {
        std::vector<uint32_t> box_index(nodes, kNone);
        for (uint32_t node : box_nodes) {
            const RtBvhNode& c = s.bvh_nodes[node];
            const uint32_t par = parent[node] >= 0 ? box_index[(size_t)parent[node]] : kNone;
            box_index[node] = pre_boxes++;
            boxes.push_back(make_float4(c.min[0], c.min[1], c.min[2], bits(par)));
            boxes.push_back(make_float4(c.max[0], c.max[1], c.max[2], bits(c.primitives_len > 0 ? 1u : 0u)));
        }
        for (uint32_t leaf : ana_leaves) {
            const RtBvhNode& n = s.bvh_nodes[leaf];
            const uint32_t p0 = pre_prims;
            // the leaf's path in the reference tree, root first, for precedes()
            std::vector<uint32_t> path;
            for (int64_t x = leaf; parent[x] >= 0; x = parent[x]) {
                const uint32_t par = (uint32_t)parent[x];
                const RtBvhNode& pn = s.bvh_nodes[par];
                const uint32_t side = (uint32_t)x == par + 1u ? 0u : 1u;
                path.push_back(first_slot_of(pn.primitives_or_second_child_index) | (pn.split_axis << 28) | (side << 30));
            }
            for (uint32_t k = 0; k < n.primitives_len; k++) {
                const RtPrimitiveInfo& info = s.primitives[n.primitives_or_second_child_index + k];
                if (info.primitive_type == 2u) continue;
                const uint32_t index = info.primitive_type == 0u ? info.index : s.sphere_count + info.index;
                out.ana_paths[index] = (uint32_t)out.ana_paths.size();
                // A primitive whose geometry repeats, bit for bit, an earlier one of the same
                // leaf that also comes first in the fallback order can never win (strict <):
                // house.toml holds its ground plane twice.
                bool shadowed = false;
                for (uint32_t j = 0; j < k && !shadowed; j++) {
                    const RtPrimitiveInfo& other = s.primitives[n.primitives_or_second_child_index + j];
                    if (other.primitive_type != info.primitive_type || other.index >= info.index) continue;
                    if (info.primitive_type == 0u) {
                        const RtSphere &a = s.spheres[other.index], &b = s.spheres[info.index];
                        shadowed = memcmp(a.pos, b.pos, 12) == 0 && memcmp(&a.radius, &b.radius, 4) == 0;
                    } else {
                        const RtPlane &a = s.planes[other.index], &b = s.planes[info.index];
                        shadowed = memcmp(a.pos, b.pos, 12) == 0 && memcmp(a.normal, b.normal, 12) == 0 &&
                                   memcmp(a.base_change, b.base_change, sizeof a.base_change) == 0;
                    }
                }
                if (shadowed) {
                    for (size_t j = path.size(); j-- > 0;) out.ana_paths.push_back(path[j]);
                    out.ana_paths.push_back(kNone);
                    continue;
                }
                for (int j = 0; j < 4; j++) prims.push_back(out.ana[(size_t)index * 4 + j]);
                pre_prims++;
                for (size_t j = path.size(); j-- > 0;) out.ana_paths.push_back(path[j]);
                out.ana_paths.push_back(kNone);
            }
            headers.push_back(make_float4(bits(box_index[leaf]), bits(p0), bits(pre_prims - p0), 0.f));
            pre_leaves++;
        }
    }
    out.pre.clear();
    out.pre.insert(out.pre.end(), headers.begin(), headers.end());
    out.pre.insert(out.pre.end(), boxes.begin(), boxes.end());
    out.pre.insert(out.pre.end(), prims.begin(), prims.end());
    if (out.pre.empty()) out.pre.push_back(make_float4(0.f, 0.f, 0.f, 0.f));

    // ---- flat walk program: leaf boxes + the visiting order of the leaves per ray octant ----
    uint32_t flat_leaves = 0;
    out.flat.assign(kFlatFloat4, make_float4(0.f, 0.f, 0.f, 0.f));
    {
        const size_t interior = root.ref == kNone || (root.ref & kLeafBit) ? 0 : n2.size() / 4;
        const size_t leaves = root.ref == kNone ? 0 : interior + 1;  // a binary tree without empty children
        if (hoisted && leaves >= 1 && leaves <= kFlatMaxLeaves && slots <= kFlatMaxSlots) {
            // storage order = the visiting order of a ray with no negative direction component
            std::vector<uint32_t> refs;
            auto visit = [&](auto&& self, uint32_t ref, uint32_t neg, std::vector<uint32_t>& seq) -> void {
                if (ref & kLeafBit) {
                    seq.push_back(ref);
                    return;
                }
                const float4* q = &n2[4u * (size_t)ref];
                const uint32_t c1 = word(q[0].w), c2 = word(q[1].w), axis = word(q[3].w);
                const bool second_first = ((neg >> axis) & 1u) != 0u;  // shader.wgsl:536
                self(self, second_first ? c2 : c1, neg, seq);
                self(self, second_first ? c1 : c2, neg, seq);
            };
            visit(visit, root.ref, 0u, refs);
            flat_leaves = (uint32_t)refs.size();
            // a leaf's reference box: find it through its first slot
            std::vector<uint32_t> node_of_first(slots, kNone);
            for (uint32_t i = 0; i < nodes; i++)
                if (s.bvh_nodes[i].primitives_len != 0)
                    node_of_first[s.bvh_nodes[i].primitives_or_second_child_index] = i;
            for (uint32_t l = 0; l < flat_leaves; l++) {
                const uint32_t first = refs[l] & kLeafSlotMask;
                const RtBvhNode& n = s.bvh_nodes[node_of_first[first]];
                out.flat[2u * l] = make_float4(n.min[0], n.min[1], n.min[2], bits(first));
                out.flat[2u * l + 1u] = make_float4(n.max[0], n.max[1], n.max[2], bits(n.primitives_len));
            }
            uint8_t* order = (uint8_t*)&out.flat[kFlatOrderAt];
            uint8_t* rank = (uint8_t*)&out.flat[kFlatRankAt];
            for (uint32_t neg = 0; neg < 8u; neg++) {
                std::vector<uint32_t> seq;
                visit(visit, root.ref, neg, seq);
                for (uint32_t i = 0; i < flat_leaves; i++) {
                    uint32_t l = 0;
                    while (refs[l] != seq[i]) l++;
                    order[32u * neg + i] = (uint8_t)l;
                    rank[32u * neg + l] = (uint8_t)i;
                }
            }
        }
    }

    FastScene& v = out.view;
    for (int k = 0; k < 3; k++) {
        v.root_min[k] = root.ref == kNone ? 0.f : root.mn[k];
        v.root_max[k] = root.ref == kNone ? 0.f : root.mx[k];
    }
    v.root_ref = root.ref;
    v.analytic_count = analytic;
    v.pre_leaves = pre_leaves;
    v.pre_boxes = pre_boxes;
    v.pre_prims = pre_prims;
    v.hoisted = hoisted ? 1u : 0u;
    v.flat_leaves = flat_leaves;
    v.flat_slots = flat_leaves ? slots : 0u;
    v.no_prune = (!hoisted && scene_has_loose_plane(s)) ? 1u : 0u;
}

}  // namespace rtf
