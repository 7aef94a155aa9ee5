// bvh.cpp — binned-SAH BVH build producing the same flattened node array and
// primitive order as the reference (src/bvh.rs:13-337). Output order is part of the
// drop-in contract: leaf primitive slots and node indices are what first-hit IDs and
// traversal order are defined on.
//
// Differences in mechanism only: the reference builds a boxed tree and flattens it in
// a second pre-order pass (bvh.rs:155-178); here nodes are emitted directly in
// pre-order while recursing, which yields the identical array. Primitive bounds and
// centroids are computed once instead of per query (same f32 values). Large subtrees are
// built on separate threads into their own arrays (indices relative to the subtree) and
// spliced in pre-order with the offsets added: the two children of a node work on disjoint
// primitive ranges and a subtree's content does not depend on where it ends up, so the
// result is the serial one byte for byte.
#include <cstdlib>
#include <cstring>
#include <future>
#include <thread>

#include "scene.hpp"

namespace rt {
namespace {

constexpr size_t kMaxPrimitivesPerLeaf = 5;  // bvh.rs:219
constexpr int kBucketCount = 12;             // bvh.rs:220

struct BuildPrim {
    uint32_t type, index;
    Bounds3 bounds;
    Vec3 center;
};

constexpr size_t kParallelMin = 1u << 16;  // subtrees below this many primitives stay on their thread

struct Builder {
    std::vector<BuildPrim>& prims;
    std::vector<RtPrimitiveInfo>& ordered;
    std::vector<RtBvhNode>& nodes;
    int tasks_left = 0;  // how many more times this subtree may fork

    // Appends a subtree built with indices relative to itself: node indices move by the
    // position it lands at, leaf slots by `slot_base`.
    void splice(const std::vector<RtBvhNode>& sub_nodes, const std::vector<RtPrimitiveInfo>& sub_ordered,
                uint32_t slot_base) {
        const uint32_t node_base = (uint32_t)nodes.size();
        const uint32_t slot_shift = slot_base;  // sub slots start at 0
        nodes.reserve(nodes.size() + sub_nodes.size());
        for (RtBvhNode n : sub_nodes) {
            n.primitives_or_second_child_index += n.primitives_len > 0 ? slot_shift : node_base;
            nodes.push_back(n);
        }
        ordered.insert(ordered.end(), sub_ordered.begin(), sub_ordered.end());
    }

    static RtBvhNode make_node(const Bounds3& b, uint32_t idx, uint32_t len, uint32_t axis) {
        RtBvhNode n;
        std::memset(&n, 0, sizeof n);
        n.min[0] = b.min.x; n.min[1] = b.min.y; n.min[2] = b.min.z;
        n.max[0] = b.max.x; n.max[1] = b.max.y; n.max[2] = b.max.z;
        n.primitives_or_second_child_index = idx;
        n.primitives_len = len;
        n.split_axis = axis;
        return n;
    }

    Bounds3 emit_leaf(size_t begin, size_t end, const Bounds3& bounds) {
        uint32_t first = (uint32_t)ordered.size();
        for (size_t i = begin; i < end; i++)
            ordered.push_back(RtPrimitiveInfo{prims[i].type, prims[i].index});
        nodes.push_back(make_node(bounds, first, (uint32_t)(end - begin), 0));  // bvh.rs:102-109
        return bounds;
    }

    // build_sah (bvh.rs:215-337) over prims[begin, end). Returns the node's bounds and
    // reports the subtree height through *depth (tree_depth, bvh.rs:124-134).
    Bounds3 build(size_t begin, size_t end, uint32_t* depth) {
        size_t count = end - begin;
        Bounds3 bounds = Bounds3::identity();
        for (size_t i = begin; i < end; i++) bounds = bounds.unite(prims[i].bounds);

        if (count <= kMaxPrimitivesPerLeaf) {
            *depth = 0;
            return emit_leaf(begin, end, bounds);
        }

        Bounds3 center_bounds = Bounds3::identity();
        for (size_t i = begin; i < end; i++) center_bounds = center_bounds.extend(prims[i].center);
        int axis = center_bounds.max_axis();
        float min_c = center_bounds.min[axis], max_c = center_bounds.max[axis];
        if (min_c == max_c) {  // bvh.rs:238-244: degenerate centroid extent -> leaf
            *depth = 0;
            return emit_leaf(begin, end, bounds);
        }

        auto bucket_of = [&](const BuildPrim& p) -> int {
            float c = p.center[axis];
            // bvh.rs:258-269: (12f32 * ((c - min) / (max - min))) as usize, 12 -> 11.
            float scaled = (float)kBucketCount * ((c - min_c) / (max_c - min_c));
            int b = scaled > 0.f ? (int)scaled : 0;  // `as usize` saturates negatives / NaN to 0
            if (b >= kBucketCount) b = kBucketCount - 1;
            return b;
        };

        size_t bucket_count[kBucketCount] = {0};
        Bounds3 bucket_bounds[kBucketCount];
        for (int b = 0; b < kBucketCount; b++) bucket_bounds[b] = Bounds3::identity();
        for (size_t i = begin; i < end; i++) {
            int b = bucket_of(prims[i]);
            bucket_count[b]++;
            bucket_bounds[b] = bucket_bounds[b].unite(prims[i].bounds);
        }

        // bvh.rs:279-292. Bounds3::from_bounds folds from the identity box, so a side made
        // only of empty buckets keeps (MAX, MIN) extents exactly like the reference.
        float costs[kBucketCount - 1];
        float total_area = bounds.surface_area();
        for (int split = 0; split < kBucketCount - 1; split++) {
            Bounds3 b0 = Bounds3::identity(), b1 = Bounds3::identity();
            size_t c0 = 0, c1 = 0;
            for (int b = 0; b <= split; b++) {
                b0 = b0.unite(bucket_bounds[b]);
                c0 += bucket_count[b];
            }
            for (int b = split + 1; b < kBucketCount; b++) {
                b1 = b1.unite(bucket_bounds[b]);
                c1 += bucket_count[b];
            }
            costs[split] =
                0.125f + ((float)c0 * b0.surface_area() + (float)c1 * b1.surface_area()) / total_area;
        }
        // min_by_key over NotNan keeps the FIRST minimum (bvh.rs:294-300).
        int min_cost_index = 0;
        for (int i = 1; i < kBucketCount - 1; i++) {
            if (costs[i] != costs[i])
                throw SceneError(RT_ERR_SCENE, "bounds have non-zero surface area");  // NotNan::new
            if (costs[i] < costs[min_cost_index]) min_cost_index = i;
        }
        if (costs[0] != costs[0]) throw SceneError(RT_ERR_SCENE, "bounds have non-zero surface area");

        // bvh.rs:304-315: in-place partition, swapping rejects to the shrinking end.
        size_t split_index = begin;
        {
            size_t end_index = end;
            while (split_index < end_index) {
                if (bucket_of(prims[split_index]) <= min_cost_index) {
                    split_index++;
                } else {
                    end_index--;
                    std::swap(prims[split_index], prims[end_index]);
                }
            }
        }
        if (split_index == begin || split_index == end) {
            // bvh.rs:317-326 falls back to select_nth_unstable_by here. The branch is
            // unreachable for a non-degenerate centroid extent: the minimum-centroid
            // primitive maps to bucket 0 (<= any split) and the maximum-centroid one to
            // bucket 11 (> any split <= 10). Fail loudly instead of guessing an order.
            throw SceneError(RT_ERR_SCENE, "BVH partition produced an empty side");
        }

        size_t node_index = nodes.size();
        nodes.push_back(make_node(bounds, 0, 0, (uint32_t)axis));
        uint32_t d0 = 0, d1 = 0;
        Bounds3 left, right;
        uint32_t second_child;
        if (count >= kParallelMin && tasks_left > 0) {
            // Build both children into private arrays, the left one on another thread.
            tasks_left--;
            std::vector<RtPrimitiveInfo> lo, ro;
            std::vector<RtBvhNode> ln, rn;
            Builder lb{prims, lo, ln, tasks_left / 2}, rb{prims, ro, rn, tasks_left - tasks_left / 2};
            auto fut = std::async(std::launch::async, [&] { return lb.build(begin, split_index, &d0); });
            right = rb.build(split_index, end, &d1);
            left = fut.get();
            const uint32_t slot0 = (uint32_t)ordered.size();
            splice(ln, lo, slot0);
            second_child = (uint32_t)nodes.size();
            splice(rn, ro, slot0 + (uint32_t)lo.size());
        } else {
            left = build(begin, split_index, &d0);
            second_child = (uint32_t)nodes.size();
            right = build(split_index, end, &d1);
        }
        // BvhBuildNode::new_interior: bounds = union of the children (bvh.rs:197-203).
        Bounds3 node_bounds = Bounds3::identity().unite(left).unite(right);
        nodes[node_index] = make_node(node_bounds, second_child, 0, (uint32_t)axis);
        *depth = (d0 > d1 ? d0 : d1) + 1;
        return node_bounds;
    }
};

}  // namespace

void build_bvh(HostScene& scene) {
    // PrimitiveInfo::build_array (bvh.rs:40-72): spheres, then planes, then triangles.
    std::vector<BuildPrim> prims;
    prims.reserve(scene.spheres.size() + scene.planes.size() + scene.u_triangles.size());
    for (size_t i = 0; i < scene.spheres.size(); i++) {
        const Sphere& s = scene.spheres[i];
        Vec3 r = v3(s.radius, s.radius, s.radius);
        Bounds3 b{s.pos - r, s.pos + r};  // scene.rs:173-180
        prims.push_back(BuildPrim{0, (uint32_t)i, b, b.center()});
    }
    for (size_t i = 0; i < scene.planes.size(); i++) {
        const Plane& p = scene.planes[i];
        // scene.rs:203-207: from_points([pos, pos + forward + right])
        Bounds3 b = Bounds3::identity().extend(p.pos).extend(p.pos + p.forward + p.right);
        prims.push_back(BuildPrim{1, (uint32_t)i, b, b.center()});
    }
    for (size_t i = 0; i < scene.u_triangles.size(); i++) {
        const RtTriangle& t = scene.u_triangles[i];
        auto vert = [&](uint32_t k) {
            const RtVec3& v = scene.u_vertices[k];
            return v3(v.x, v.y, v.z);
        };
        // mesh.rs:143-147
        Bounds3 b = Bounds3::identity().extend(vert(t.vertex_0)).extend(vert(t.vertex_1)).extend(
            vert(t.vertex_2));
        prims.push_back(BuildPrim{2, (uint32_t)i, b, b.center()});
    }

    scene.u_primitives.clear();
    scene.u_nodes.clear();
    scene.bvh_depth = 0;
    if (prims.empty()) throw SceneError(RT_ERR_SCENE, "scene has no primitives");  // bvh.rs:222
    scene.u_primitives.reserve(prims.size());
    int threads = (int)std::thread::hardware_concurrency();
    if (threads < 1) threads = 1;
    if (threads > 64) threads = 64;
    if (const char* k = std::getenv("RT_BVH_THREADS")) threads = std::atoi(k) > 0 ? std::atoi(k) : 1;
    Builder builder{prims, scene.u_primitives, scene.u_nodes, threads - 1};
    builder.build(0, prims.size(), &scene.bvh_depth);
}

}  // namespace rt
