// rt_fast_host.hpp — host-side construction of the fast device layout (rt_fast.cuh) from
// the reference-layout scene arrays (include/rt_types.h). Inputs must already be validated
// (rt_upload_scene does that); shared with the CPU emulation used by the tests.
#pragma once

#include <cmath>
#include <cstring>
#include <vector>

#include "rt_fast.cuh"
#include "rt_types.h"

namespace rtf {

inline float bits(uint32_t u) {
    float f;
    memcpy(&f, &u, 4);
    return f;
}

// ---- fast layout (rt_fast.cuh): N2 interior records, 48-byte triangle records, analytic
//      records and the pre-pass program. Inputs are already validated. ----
struct FastLayout {
    std::vector<float4> n2, tri, ana, pre;
    std::vector<uint32_t> big_len;
    FastScene view{};
};

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

inline void build_fast_layout(const RtSceneDesc& s, FastLayout& out) {
        const uint32_t nodes = s.bvh_node_count, slots = s.primitive_count;
    const uint32_t analytic = s.sphere_count + s.plane_count;

    // parents (for the box chains) and interior record numbering in node order
    std::vector<int64_t> parent(nodes, -1);
    std::vector<uint32_t> record_of(nodes, kNone);
    uint32_t interior = 0;
    for (uint32_t i = 0; i < nodes; i++) {
        const RtBvhNode& n = s.bvh_nodes[i];
        if (n.primitives_len == 0) {
            record_of[i] = interior++;
            parent[i + 1] = i;
            parent[n.primitives_or_second_child_index] = i;
        }
    }

    // hoisting decision: few analytic primitives, short chains
    std::vector<uint32_t> ana_leaves;  // leaves holding at least one analytic primitive, in node (= slot) order
    uint32_t chain_boxes = 0;
    for (uint32_t i = 0; i < nodes; i++) {
        const RtBvhNode& n = s.bvh_nodes[i];
        if (n.primitives_len == 0) continue;
        bool any = false;
        for (uint32_t k = 0; k < n.primitives_len; k++)
            any |= s.primitives[n.primitives_or_second_child_index + k].primitive_type != 2u;
        if (!any) continue;
        ana_leaves.push_back(i);
        for (int64_t x = i; x > 0; x = parent[x]) chain_boxes++;  // the root is tested once for all
    }
    const bool hoisted = analytic <= kPreMaxPrims && ana_leaves.size() <= kPreMaxLeaves &&
                         chain_boxes <= kPreMaxBoxes;

    auto leaf_all_analytic = [&](const RtBvhNode& n) {
        for (uint32_t k = 0; k < n.primitives_len; k++)
            if (s.primitives[n.primitives_or_second_child_index + k].primitive_type == 2u) return false;
        return true;
    };
    bool any_big = false;
    auto ref_of = [&](uint32_t i) -> uint32_t {
        const RtBvhNode& n = s.bvh_nodes[i];
        if (n.primitives_len == 0) return record_of[i];
        if (hoisted && leaf_all_analytic(n)) return kNone;
        uint32_t len = n.primitives_len;
        if (len >= kLeafLenEscape) {
            any_big = true;
            len = kLeafLenEscape;
        }
        return kLeafBit | (len << kLeafLenShift) | n.primitives_or_second_child_index;
    };
    auto first_slot_of = [&](uint32_t i) {
        while (s.bvh_nodes[i].primitives_len == 0) i = i + 1;
        return s.bvh_nodes[i].primitives_or_second_child_index;
    };

    out.n2.assign((size_t)interior * 4, make_float4(0.f, 0.f, 0.f, 0.f));
    for (uint32_t i = 0; i < nodes; i++) {
        const RtBvhNode& n = s.bvh_nodes[i];
        if (n.primitives_len != 0) continue;
        const uint32_t c1 = i + 1, c2 = n.primitives_or_second_child_index;
        const RtBvhNode &a = s.bvh_nodes[c1], &b = s.bvh_nodes[c2];
        float4* q = &out.n2[(size_t)record_of[i] * 4];
        q[0] = make_float4(a.min[0], a.min[1], a.min[2], bits(ref_of(c1)));
        q[1] = make_float4(a.max[0], a.max[1], a.max[2], bits(ref_of(c2)));
        q[2] = make_float4(b.min[0], b.min[1], b.min[2], bits(first_slot_of(c2)));
        q[3] = make_float4(b.max[0], b.max[1], b.max[2], bits(n.split_axis));
    }
    if (out.n2.empty()) out.n2.assign(4, make_float4(0.f, 0.f, 0.f, 0.f));  // never read: root is a leaf

    // analytic records: spheres then planes, reference order; slots filled in below
    out.ana.assign((size_t)(analytic ? analytic : 1) * 4, make_float4(0.f, 0.f, 0.f, 0.f));
    std::vector<uint32_t> slot_of_analytic(analytic, kNone);
    out.tri.assign((size_t)slots * 3, make_float4(0.f, 0.f, 0.f, 0.f));
    for (uint32_t slot = 0; slot < slots; slot++) {
        const RtPrimitiveInfo& info = s.primitives[slot];
        if (info.primitive_type == 2u) {
            const RtTriangle& t = s.triangles[info.index];
            const RtVec3 &a = s.vertices[t.vertex_0], &b = s.vertices[t.vertex_1], &c = s.vertices[t.vertex_2];
            // edge_0 = b - a, edge_1 = c - a (shader.wgsl:415-416): single f32 subtractions,
            // so hoisting them to upload time leaves every bit unchanged.
            out.tri[(size_t)slot * 3 + 0] = make_float4(a.x, a.y, a.z, bits(2u));
            out.tri[(size_t)slot * 3 + 1] = make_float4(b.x - a.x, b.y - a.y, b.z - a.z, bits(t.material_id));
            out.tri[(size_t)slot * 3 + 2] = make_float4(c.x - a.x, c.y - a.y, c.z - a.z, 0.f);
        } else {
            const uint32_t index = info.primitive_type == 0u ? info.index : s.sphere_count + info.index;
            slot_of_analytic[index] = slot;
            out.tri[(size_t)slot * 3 + 0] = make_float4(bits(index), 0.f, 0.f, bits(info.primitive_type));
        }
    }
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

    // pre-pass program: headers | boxes | primitive records
    uint32_t pre_leaves = 0, pre_boxes = 0, pre_prims = 0;
    std::vector<float4> headers, boxes, prims;
    if (hoisted) {
        for (uint32_t leaf : ana_leaves) {
            const RtBvhNode& n = s.bvh_nodes[leaf];
            const uint32_t b0 = pre_boxes, p0 = pre_prims;
            for (int64_t x = leaf; x > 0; x = parent[x]) {
                const RtBvhNode& c = s.bvh_nodes[x];
                boxes.push_back(make_float4(c.min[0], c.min[1], c.min[2], 0.f));
                boxes.push_back(make_float4(c.max[0], c.max[1], c.max[2], 0.f));
                pre_boxes++;
            }
            for (uint32_t k = 0; k < n.primitives_len; k++) {
                const RtPrimitiveInfo& info = s.primitives[n.primitives_or_second_child_index + k];
                if (info.primitive_type == 2u) continue;
                const uint32_t index = info.primitive_type == 0u ? info.index : s.sphere_count + info.index;
                for (int j = 0; j < 4; j++) prims.push_back(out.ana[(size_t)index * 4 + j]);
                pre_prims++;
            }
            headers.push_back(make_float4(bits(b0), bits(pre_boxes - b0), bits(p0), bits(pre_prims - p0)));
            pre_leaves++;
        }
        // analytic primitives that sit in no leaf cannot exist: every primitive has a slot
    }
    out.pre.clear();
    out.pre.insert(out.pre.end(), headers.begin(), headers.end());
    out.pre.insert(out.pre.end(), boxes.begin(), boxes.end());
    out.pre.insert(out.pre.end(), prims.begin(), prims.end());
    if (out.pre.empty()) out.pre.push_back(make_float4(0.f, 0.f, 0.f, 0.f));

    FastScene& v = out.view;
    const RtBvhNode& root = s.bvh_nodes[0];
    for (int k = 0; k < 3; k++) {
        v.root_min[k] = root.min[k];
        v.root_max[k] = root.max[k];
    }
    v.root_ref = ref_of(0);
    v.analytic_count = analytic;
    v.pre_leaves = pre_leaves;
    v.pre_boxes = pre_boxes;
    v.pre_prims = pre_prims;
    v.hoisted = hoisted ? 1u : 0u;
    v.no_prune = (!hoisted && scene_has_loose_plane(s)) ? 1u : 0u;
}


}  // namespace rtf
