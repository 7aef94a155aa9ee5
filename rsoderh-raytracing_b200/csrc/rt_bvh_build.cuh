// rt_bvh_build.cuh — build_sah (src/bvh.rs:215-337) on the device, level by level, with the
// reference's output byte for byte: same flattened pre-order node array (bvh.rs:155-178), same
// leaf slot order, same bounds down to the sign of a zero.
//
// The reference recurses depth first and partitions each node's primitives with a sequential
// swap loop (bvh.rs:304-315: rejects are swapped to a shrinking end). Three observations make
// that a data-parallel build with an identical result:
//   * every decision of a node — centroid extent, axis, the 12 bucket boxes and counts, the 11
//     SAH costs — depends on its primitives as a SET (min / max / count are exact and
//     order-free; the sign of a zero never reaches a comparison or a cost), so the nodes of one
//     tree level are evaluated together, one pass over all primitives per level;
//   * the swap loop is a fixed permutation of its input with a closed form (partition_dest
//     below; tests/test_bvh_levels.py checks it against the literal loop): it needs one
//     prefix count of the "stays left" flags and two small scatter tables, all parallel;
//   * node indices are pre-order positions, which follow from subtree sizes: the tree is built
//     in creation order, then sizes go bottom-up and positions top-down, level by level.
// Only the FINAL bounds are order-sensitive (f32::min / max keep their first argument on ties,
// so -0.0 and +0.0 resolve by position): leaves fold their primitives in slot order and
// interior nodes fold (left, right), exactly as Bounds3::from_bounds does (bvh.rs:197-203).
//
// Memory (N primitives): two 32-byte record buffers, two node-id maps, prefix / scatter tables
// (~90 B per primitive), node arrays for up to 2N nodes (104 B each) and 624 B of bucket
// accumulators per node that can split in one level (at most N / 6 of them).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include "rt_types.h"

namespace rtb {

constexpr uint32_t kNoNode = 0xFFFFFFFFu;
constexpr int kBuckets = 12;           // bvh.rs:220
constexpr uint32_t kLeafMax = 5u;      // bvh.rs:219
constexpr uint32_t kSplitLeaf = 0xFFu; // split[] of a node that stays a leaf
constexpr int kFields = 13;            // count, box min xyz, box max xyz, centroid min xyz, centroid max xyz
constexpr int kThreads = 256;
constexpr int kScanItems = 4;          // items per thread in the prefix kernels
constexpr int kScanBlock = kThreads * kScanItems;
constexpr float kMax = 3.40282347e+38f;  // Vec3::MAX / Vec3::MIN of Bounds3::identity (scene.rs:67-72)

enum BuildError : uint32_t { kErrNanCost = 1u, kErrEmptySide = 2u, kErrNodeOverflow = 4u, kErrBadIndex = 8u };

// f32::min / f32::max without NaNs (the first argument stays on ties, -0.0 == +0.0 included).
__device__ __forceinline__ float min_rs(float a, float b) { return b < a ? b : a; }
__device__ __forceinline__ float max_rs(float a, float b) { return b > a ? b : a; }

// Order-preserving integer image of a float, for atomicMin / atomicMax.
__device__ __forceinline__ int fkey(float f) {
    const int b = __float_as_int(f);
    return b >= 0 ? b : b ^ 0x7FFFFFFF;
}
__device__ __forceinline__ float fkey_inv(int k) { return __int_as_float(k >= 0 ? k : k ^ 0x7FFFFFFF); }

struct Nodes {
    uint32_t *begin, *count, *left;  // left child id (right = left + 1), kNoNode for a leaf
    uint32_t *axis, *split, *good;   // split = chosen bucket boundary (kSplitLeaf: leaf), good = primitives that stay left
    uint32_t *slot;                  // accumulator slot of the node in its level
    float4 *bmin, *bmax;             // box of the node's primitives (set-exact; final, order-exact after finish)
    float4 *cmin, *cmax;             // box of their centroids
    uint32_t *size, *depth, *pre;    // subtree node count, tree_depth (bvh.rs:124-134), pre-order position
};

struct Build {
    uint32_t n;              // primitives
    uint32_t node_cap;
    uint32_t acc_cap;        // accumulator slots
    float4 *rmin[2], *rmax[2];  // records: (box min, bits type) (box max, bits index), double buffered
    uint32_t* node_of[2];    // node a position belongs to (kNoNode once it sits in a leaf)
    uint8_t* bucket;
    uint32_t *prefix;        // [n + 1] exclusive count of "stays left" flags
    uint32_t *frpos, *bgpos; // scatter tables of the partition (positions of front rejects / back goods by rank)
    uint32_t* blocksum;
    int* acc;                // [kFields][kBuckets][acc_cap]
    uint32_t* active[2];     // nodes that may split in this / the next level
    uint32_t* counters;      // [0] node count, [1] next level's active count, [2] error bits
    Nodes nodes;
};

__device__ __forceinline__ int* acc_at(const Build& b, int field, int bucket, uint32_t slot) {
    return b.acc + ((size_t)(field * kBuckets + bucket) * b.acc_cap + slot);
}

// PrimitiveInfo::build_array (bvh.rs:40-72): spheres, planes, triangles, with their bounds
// (scene.rs:173-180, 203-207, mesh.rs:143-147).
__global__ void init_records_kernel(Build b, const RtSphere* spheres, uint32_t sphere_count, const float* plane_bounds,
                                    uint32_t plane_count, const RtVec3* vertices, uint32_t vertex_count,
                                    const RtTriangle* triangles, uint32_t root_is_leaf) {
    const uint32_t p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= b.n) return;
    float lo[3], hi[3];
    uint32_t type, index;
    if (p < sphere_count) {
        const RtSphere s = spheres[p];
        for (int k = 0; k < 3; k++) {
            lo[k] = s.pos[k] - s.radius;
            hi[k] = s.pos[k] + s.radius;
        }
        type = 0u;
        index = p;
    } else if (p < sphere_count + plane_count) {
        index = p - sphere_count;
        for (int k = 0; k < 3; k++) {
            lo[k] = plane_bounds[6u * index + k];
            hi[k] = plane_bounds[6u * index + 3 + k];
        }
        type = 1u;
    } else {
        index = p - sphere_count - plane_count;
        const RtTriangle t = triangles[index];
        uint32_t vi[3] = {t.vertex_0, t.vertex_1, t.vertex_2};
        if (vi[0] >= vertex_count || vi[1] >= vertex_count || vi[2] >= vertex_count) {
            atomicOr(&b.counters[2], (uint32_t)kErrBadIndex);  // the build fails; stay in bounds
            vi[0] = vi[1] = vi[2] = 0u;
        }
        for (int k = 0; k < 3; k++) {
            lo[k] = kMax;
            hi[k] = -kMax;
        }
        for (int j = 0; j < 3; j++) {
            const RtVec3 v = vertices[vi[j]];
            const float c[3] = {v.x, v.y, v.z};
            for (int k = 0; k < 3; k++) {
                lo[k] = min_rs(lo[k], c[k]);
                hi[k] = max_rs(hi[k], c[k]);
            }
        }
        type = 2u;
    }
    b.rmin[0][p] = make_float4(lo[0], lo[1], lo[2], __uint_as_float(type));
    b.rmax[0][p] = make_float4(hi[0], hi[1], hi[2], __uint_as_float(index));
    b.node_of[0][p] = root_is_leaf ? kNoNode : 0u;
}

__global__ void acc_init_kernel(Build b, uint32_t slots) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const size_t per_field = (size_t)kBuckets * slots;
    if (i >= per_field * kFields) return;
    const int field = (int)(i / per_field);
    const size_t rest = i % per_field;
    const int bucket = (int)(rest / slots);
    const uint32_t slot = (uint32_t)(rest % slots);
    // count 0; minima start at +MAX, maxima at -MAX (Bounds3::identity)
    int v = 0;
    if (field >= 1 && field <= 3) v = fkey(kMax);
    else if (field >= 4 && field <= 6) v = fkey(-kMax);
    else if (field >= 7 && field <= 9) v = fkey(kMax);
    else if (field >= 10) v = fkey(-kMax);
    *acc_at(b, field, bucket, slot) = v;
}

// Bounds3::center (scene.rs:102-104): min * 0.5 + max * 0.5.
__device__ __forceinline__ float center_of(float lo, float hi) { return lo * 0.5f + hi * 0.5f; }

// bucket_index_of (bvh.rs:258-269): (12f32 * ((c - min) / (max - min))) as usize, 12 -> 11.
__device__ __forceinline__ int bucket_of(float c, float min_c, float max_c) {
    const float scaled = (float)kBuckets * ((c - min_c) / (max_c - min_c));
    int k = scaled > 0.0f ? (int)scaled : 0;  // `as usize` saturates negatives / NaN to 0
    if (k >= kBuckets) k = kBuckets - 1;
    return k;
}

// One pass over the primitives of a level: every primitive of a node that may split finds its
// bucket and adds itself to the node's bucket accumulators. ROOT: the pass that measures the whole
// input (one bucket of one pseudo node). A block whose positions all belong to one node first
// gathers in shared memory (the top levels: millions of primitives on a dozen buckets).
template <bool ROOT>
__global__ void __launch_bounds__(kThreads) bucket_kernel(Build b, int cur) {
    __shared__ int sacc[kFields * kBuckets];
    __shared__ uint32_t s_first, s_last;
    const uint32_t p = blockIdx.x * kThreads + threadIdx.x;
    const bool in = p < b.n;
    uint32_t node = kNoNode;
    if (in) node = ROOT ? 0u : b.node_of[cur][p];
    if (!ROOT && node != kNoNode && b.nodes.split[node] == kSplitLeaf) {
        // the node turned out to be a leaf (degenerate centroid extent, bvh.rs:238-244)
        b.node_of[cur][p] = kNoNode;
        node = kNoNode;
    }
    if (threadIdx.x == 0) s_first = node;
    const uint32_t last_thread = min((uint32_t)kThreads, b.n - blockIdx.x * kThreads) - 1u;
    if (threadIdx.x == last_thread) s_last = node;
    for (int i = threadIdx.x; i < kFields * kBuckets; i += kThreads) {
        const int field = i / kBuckets;
        sacc[i] = field == 0 ? 0 : ((field >= 1 && field <= 3) || (field >= 7 && field <= 9) ? fkey(kMax) : fkey(-kMax));
    }
    __syncthreads();
    // positions of a node are contiguous and nodes appear in position order: first == last means
    // the whole block is one node
    const bool uniform = s_first == s_last && s_first != kNoNode;
    int k = 0;
    float v[kFields];
    if (node != kNoNode) {
        const float4 lo = b.rmin[cur][p], hi = b.rmax[cur][p];
        const float c[3] = {center_of(lo.x, hi.x), center_of(lo.y, hi.y), center_of(lo.z, hi.z)};
        if (!ROOT) {
            const uint32_t axis = b.nodes.axis[node];
            const float4 cmin = b.nodes.cmin[node], cmax = b.nodes.cmax[node];
            const float min_c = axis == 0u ? cmin.x : (axis == 1u ? cmin.y : cmin.z);
            const float max_c = axis == 0u ? cmax.x : (axis == 1u ? cmax.y : cmax.z);
            k = bucket_of(c[axis], min_c, max_c);
            b.bucket[p] = (uint8_t)k;
        }
        v[1] = lo.x; v[2] = lo.y; v[3] = lo.z;
        v[4] = hi.x; v[5] = hi.y; v[6] = hi.z;
        v[7] = c[0]; v[8] = c[1]; v[9] = c[2];
        v[10] = c[0]; v[11] = c[1]; v[12] = c[2];
        if (uniform) {
            atomicAdd(&sacc[0 * kBuckets + k], 1);
            for (int f = 1; f < kFields; f++) {
                const bool is_min = (f >= 1 && f <= 3) || (f >= 7 && f <= 9);
                if (is_min) atomicMin(&sacc[f * kBuckets + k], fkey(v[f]));
                else atomicMax(&sacc[f * kBuckets + k], fkey(v[f]));
            }
        } else {
            const uint32_t slot = ROOT ? 0u : b.nodes.slot[node];
            atomicAdd(acc_at(b, 0, k, slot), 1);
            for (int f = 1; f < kFields; f++) {
                const bool is_min = (f >= 1 && f <= 3) || (f >= 7 && f <= 9);
                if (is_min) atomicMin(acc_at(b, f, k, slot), fkey(v[f]));
                else atomicMax(acc_at(b, f, k, slot), fkey(v[f]));
            }
        }
    } else if (in && !ROOT) {
        b.bucket[p] = 0xFFu;
    }
    if (uniform) {
        __syncthreads();
        const uint32_t slot = ROOT ? 0u : b.nodes.slot[s_first];
        for (int i = threadIdx.x; i < kFields * kBuckets; i += kThreads) {
            const int field = i / kBuckets, bucket = i % kBuckets;
            if (sacc[0 * kBuckets + bucket] == 0) continue;  // nothing of this block fell into the bucket
            const bool is_min = (field >= 1 && field <= 3) || (field >= 7 && field <= 9);
            if (field == 0) atomicAdd(acc_at(b, 0, bucket, slot), sacc[i]);
            else if (is_min) atomicMin(acc_at(b, field, bucket, slot), sacc[i]);
            else atomicMax(acc_at(b, field, bucket, slot), sacc[i]);
        }
    }
}

struct Box {
    float lo[3], hi[3];
};
__device__ __forceinline__ Box box_identity() {
    Box x;
    for (int k = 0; k < 3; k++) {
        x.lo[k] = kMax;
        x.hi[k] = -kMax;
    }
    return x;
}
__device__ __forceinline__ Box box_unite(const Box& a, const Box& o) {
    Box x;
    for (int k = 0; k < 3; k++) {
        x.lo[k] = min_rs(a.lo[k], o.lo[k]);
        x.hi[k] = max_rs(a.hi[k], o.hi[k]);
    }
    return x;
}
// Bounds3::surface_area (scene.rs:125-128)
__device__ __forceinline__ float box_area(const Box& x) {
    const float dx = x.hi[0] - x.lo[0], dy = x.hi[1] - x.lo[1], dz = x.hi[2] - x.lo[2];
    return 2.0f * (dx * dy + dx * dz + dy * dz);
}
// Bounds3::max_axis (scene.rs:113-122)
__device__ __forceinline__ uint32_t box_max_axis(const Box& x) {
    const float dx = x.hi[0] - x.lo[0], dy = x.hi[1] - x.lo[1], dz = x.hi[2] - x.lo[2];
    if (dz > dx && dz > dy) return 2u;
    if (dy > dx) return 1u;
    return 0u;
}

__device__ __forceinline__ Box acc_box(const Build& b, int first_field, int bucket, uint32_t slot) {
    Box x;
    for (int k = 0; k < 3; k++) {
        x.lo[k] = fkey_inv(*acc_at(b, first_field + k, bucket, slot));
        x.hi[k] = fkey_inv(*acc_at(b, first_field + 3 + k, bucket, slot));
    }
    return x;
}

// A node with its primitive set measured: leaf, or a candidate for splitting in the next level
// (bvh.rs:227-244).
__device__ __forceinline__ void node_create(const Build& b, uint32_t id, uint32_t begin, uint32_t count, const Box& box,
                                            const Box& cbox, uint32_t* next_active, uint32_t* next_count) {
    Nodes nd = b.nodes;
    nd.begin[id] = begin;
    nd.count[id] = count;
    nd.left[id] = kNoNode;
    nd.bmin[id] = make_float4(box.lo[0], box.lo[1], box.lo[2], 0.0f);
    nd.bmax[id] = make_float4(box.hi[0], box.hi[1], box.hi[2], 0.0f);
    nd.cmin[id] = make_float4(cbox.lo[0], cbox.lo[1], cbox.lo[2], 0.0f);
    nd.cmax[id] = make_float4(cbox.hi[0], cbox.hi[1], cbox.hi[2], 0.0f);
    nd.good[id] = 0u;
    nd.size[id] = 1u;
    nd.depth[id] = 0u;
    bool leaf = count <= kLeafMax;
    uint32_t axis = 0u;
    if (!leaf) {
        axis = box_max_axis(cbox);
        leaf = cbox.lo[axis] == cbox.hi[axis];  // degenerate along the axis
    }
    // A leaf's split_axis is 0 (bvh.rs:102-109); an interior node keeps the axis of its centroids.
    nd.axis[id] = leaf ? 0u : axis;
    nd.split[id] = kSplitLeaf;
    if (!leaf) {
        const uint32_t j = atomicAdd(next_count, 1u);
        next_active[j] = id;
        nd.slot[id] = j;
        nd.split[id] = 0u;
    }
}

__global__ void root_kernel(Build b) {
    if (blockIdx.x || threadIdx.x) return;
    b.counters[0] = 1u;
    b.counters[1] = 0u;
    node_create(b, 0u, 0u, b.n, acc_box(b, 1, 0, 0u), acc_box(b, 7, 0, 0u), b.active[0], &b.counters[1]);
}

// One thread per candidate node: the 11 SAH costs in the reference's evaluation order, the first
// minimum (bvh.rs:279-300), then the two children with the boxes of their bucket ranges.
__global__ void split_kernel(Build b, int cur, uint32_t active_count) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= active_count) return;
    const uint32_t node = b.active[cur][i];
    const uint32_t slot = i;  // == nodes.slot[node]
    Box boxes[kBuckets];
    uint32_t counts[kBuckets];
    for (int k = 0; k < kBuckets; k++) {
        counts[k] = (uint32_t)*acc_at(b, 0, k, slot);
        boxes[k] = acc_box(b, 1, k, slot);
    }
    Box whole;
    {
        const float4 lo = b.nodes.bmin[node], hi = b.nodes.bmax[node];
        whole.lo[0] = lo.x; whole.lo[1] = lo.y; whole.lo[2] = lo.z;
        whole.hi[0] = hi.x; whole.hi[1] = hi.y; whole.hi[2] = hi.z;
    }
    const float total_area = box_area(whole);
    int best = 0;
    float best_cost = 0.0f;
    bool nan = false;
    for (int split = 0; split < kBuckets - 1; split++) {
        Box b0 = box_identity(), b1 = box_identity();
        unsigned long long c0 = 0, c1 = 0;
        for (int k = 0; k <= split; k++) {
            b0 = box_unite(b0, boxes[k]);
            c0 += counts[k];
        }
        for (int k = split + 1; k < kBuckets; k++) {
            b1 = box_unite(b1, boxes[k]);
            c1 += counts[k];
        }
        const float cost = 0.125f + ((float)c0 * box_area(b0) + (float)c1 * box_area(b1)) / total_area;
        if (cost != cost) nan = true;  // NotNan::new(..).expect("bounds have non-zero surface area")
        if (split == 0 || cost < best_cost) {
            best = split;
            best_cost = cost;
        }
    }
    uint32_t good = 0;
    for (int k = 0; k <= best; k++) good += counts[k];
    const uint32_t count = b.nodes.count[node];
    if (nan) atomicOr(&b.counters[2], (uint32_t)kErrNanCost);
    if (good == 0u || good == count) atomicOr(&b.counters[2], (uint32_t)kErrEmptySide);
    if (nan || good == 0u || good == count) {
        b.nodes.split[node] = kSplitLeaf;  // keeps the later passes in bounds; the build fails
        return;
    }
    const uint32_t left = atomicAdd(&b.counters[0], 2u);
    if (left + 2u > b.node_cap) {
        atomicOr(&b.counters[2], (uint32_t)kErrNodeOverflow);
        b.nodes.split[node] = kSplitLeaf;
        return;
    }
    Box lb = box_identity(), lc = box_identity(), rb = box_identity(), rc = box_identity();
    for (int k = 0; k < kBuckets; k++) {
        if (counts[k] == 0u) continue;
        if (k <= best) {
            lb = box_unite(lb, boxes[k]);
            lc = box_unite(lc, acc_box(b, 7, k, slot));
        } else {
            rb = box_unite(rb, boxes[k]);
            rc = box_unite(rc, acc_box(b, 7, k, slot));
        }
    }
    const uint32_t begin = b.nodes.begin[node];
    b.nodes.left[node] = left;
    b.nodes.split[node] = (uint32_t)best;
    b.nodes.good[node] = good;
    node_create(b, left, begin, good, lb, lc, b.active[cur ^ 1], &b.counters[1]);
    node_create(b, left + 1u, begin + good, count - good, rb, rc, b.active[cur ^ 1], &b.counters[1]);
}

// "stays left" flag of a position (bvh.rs:307: bucket <= min_cost_index)
__device__ __forceinline__ uint32_t good_flag(const Build& b, int cur, uint32_t p) {
    if (p >= b.n) return 0u;
    const uint32_t node = b.node_of[cur][p];
    if (node == kNoNode) return 0u;
    const uint32_t split = b.nodes.split[node];
    return (split != kSplitLeaf && (uint32_t)b.bucket[p] <= split) ? 1u : 0u;
}

__device__ __forceinline__ uint32_t block_sum(uint32_t v, uint32_t* warp_sums) {
    for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(0xFFFFFFFFu, v, off);
    if ((threadIdx.x & 31u) == 0u) warp_sums[threadIdx.x >> 5] = v;
    __syncthreads();
    uint32_t total = 0;
    for (int w = 0; w < kThreads / 32; w++) total += warp_sums[w];
    __syncthreads();
    return total;
}

__global__ void __launch_bounds__(kThreads) flag_count_kernel(Build b, int cur) {
    __shared__ uint32_t warp_sums[kThreads / 32];
    const uint32_t base = blockIdx.x * kScanBlock + threadIdx.x * kScanItems;
    uint32_t v = 0;
    for (int j = 0; j < kScanItems; j++) v += good_flag(b, cur, base + j);
    const uint32_t total = block_sum(v, warp_sums);
    if (threadIdx.x == 0) b.blocksum[blockIdx.x] = total;
}

// Exclusive scan of the block sums by one block (N / 1024 entries).
__global__ void __launch_bounds__(1024) blocksum_scan_kernel(Build b, uint32_t blocks) {
    __shared__ uint32_t warp_sums[32];
    __shared__ uint32_t carry;
    if (threadIdx.x == 0) carry = 0u;
    __syncthreads();
    for (uint32_t base = 0; base < blocks; base += 1024u) {
        const uint32_t i = base + threadIdx.x;
        const uint32_t v = i < blocks ? b.blocksum[i] : 0u;
        uint32_t inc = v;
        for (int off = 1; off < 32; off <<= 1) {
            const uint32_t up = __shfl_up_sync(0xFFFFFFFFu, inc, off);
            if ((threadIdx.x & 31u) >= (uint32_t)off) inc += up;
        }
        if ((threadIdx.x & 31u) == 31u) warp_sums[threadIdx.x >> 5] = inc;
        __syncthreads();
        uint32_t before = carry;
        for (uint32_t w = 0; w < (threadIdx.x >> 5); w++) before += warp_sums[w];
        if (i < blocks) b.blocksum[i] = before + inc - v;
        __syncthreads();
        if (threadIdx.x == 1023u) carry = before + inc;
        __syncthreads();
    }
    if (threadIdx.x == 0) b.blocksum[blocks] = carry;  // grand total
}

__global__ void __launch_bounds__(kThreads) prefix_kernel(Build b, int cur) {
    __shared__ uint32_t warp_sums[kThreads / 32];
    const uint32_t base = blockIdx.x * kScanBlock + threadIdx.x * kScanItems;
    uint32_t f[kScanItems], v = 0;
    for (int j = 0; j < kScanItems; j++) {
        f[j] = good_flag(b, cur, base + j);
        v += f[j];
    }
    uint32_t inc = v;
    for (int off = 1; off < 32; off <<= 1) {
        const uint32_t up = __shfl_up_sync(0xFFFFFFFFu, inc, off);
        if ((threadIdx.x & 31u) >= (uint32_t)off) inc += up;
    }
    if ((threadIdx.x & 31u) == 31u) warp_sums[threadIdx.x >> 5] = inc;
    __syncthreads();
    uint32_t before = b.blocksum[blockIdx.x];
    for (uint32_t w = 0; w < (threadIdx.x >> 5); w++) before += warp_sums[w];
    uint32_t run = before + inc - v;
    for (int j = 0; j < kScanItems; j++) {
        if (base + j <= b.n) b.prefix[base + j] = run;  // prefix[n] = total
        run += f[j];
    }
}

// The reference's partition loop (bvh.rs:304-315)
//     while split < end { if good(a[split]) { split += 1 } else { end -= 1; swap(a[split], a[end]) } }
// as a permutation. With G good elements, the first G positions are the "front", the rest the
// "back". Good front elements stay. The loop meets the front rejects in position order and, for
// each, pulls back elements from the top down until a good one arrives, which takes the reject's
// place; every reject it meets lands one below the previous one, starting at the last position.
// Hence: back good number m (from the top) goes to front reject number m's position; front reject
// number 0 goes to the last position, number m > 0 to just below back good number m - 1; a back
// reject moves down by one — except the one at position G (if a reject), which is met last of its
// run: it goes just below the last back good (or to the last position if the back has no good).
struct PartitionCtx {
    uint32_t begin, count, good;  // node range, number of good elements
    uint32_t goods_before;        // good elements of the node before the position
    uint32_t front_rejects;       // = back goods
};
__device__ __forceinline__ uint32_t partition_dest(const PartitionCtx& c, uint32_t p, bool is_good, const uint32_t* frpos,
                                                   const uint32_t* bgpos) {
    const uint32_t rel = p - c.begin, end = c.begin + c.count;
    if (rel < c.good) {
        if (is_good) return p;
        const uint32_t m = rel - c.goods_before;  // rank among the front rejects
        return m == 0u ? end - 1u : bgpos[c.begin + m - 1u] - 1u;
    }
    if (is_good) return frpos[c.begin + (c.good - c.goods_before - 1u)];  // rank from the top among back goods
    if (rel == c.good) return c.front_rejects ? bgpos[c.begin + c.front_rejects - 1u] - 1u : end - 1u;
    return p - 1u;
}

__global__ void __launch_bounds__(kThreads) scatter_kernel(Build b, int cur) {
    const uint32_t p = blockIdx.x * kThreads + threadIdx.x;
    if (p >= b.n) return;
    const uint32_t node = b.node_of[cur][p];
    if (node == kNoNode) return;
    const uint32_t split = b.nodes.split[node];
    if (split == kSplitLeaf) return;
    const uint32_t begin = b.nodes.begin[node], good = b.nodes.good[node];
    const uint32_t goods_before = b.prefix[p] - b.prefix[begin];
    const bool is_good = (uint32_t)b.bucket[p] <= split;
    const uint32_t rel = p - begin;
    if (rel < good) {
        if (!is_good) b.frpos[begin + (rel - goods_before)] = p;
    } else if (is_good) {
        b.bgpos[begin + (good - goods_before - 1u)] = p;
    }
}

__global__ void __launch_bounds__(kThreads) move_kernel(Build b, int cur) {
    const uint32_t p = blockIdx.x * kThreads + threadIdx.x;
    if (p >= b.n) return;
    const int nxt = cur ^ 1;
    const uint32_t node = b.node_of[cur][p];
    const float4 lo = b.rmin[cur][p], hi = b.rmax[cur][p];
    uint32_t dest = p, owner = kNoNode;
    if (node != kNoNode) {
        const uint32_t split = b.nodes.split[node];
        if (split != kSplitLeaf) {
            PartitionCtx c;
            c.begin = b.nodes.begin[node];
            c.count = b.nodes.count[node];
            c.good = b.nodes.good[node];
            const uint32_t at_begin = b.prefix[c.begin];
            c.goods_before = b.prefix[p] - at_begin;
            c.front_rejects = c.good - (b.prefix[c.begin + c.good] - at_begin);
            dest = partition_dest(c, p, (uint32_t)b.bucket[p] <= split, b.frpos, b.bgpos);
            const uint32_t child = b.nodes.left[node] + (dest - c.begin < c.good ? 0u : 1u);
            owner = b.nodes.split[child] == kSplitLeaf ? kNoNode : child;
        }
    }
    b.rmin[nxt][dest] = lo;
    b.rmax[nxt][dest] = hi;
    b.node_of[nxt][dest] = owner;
}

// ---- finish: exact bounds, subtree sizes, pre-order positions, output ----
// Leaves fold their primitives in slot order (Bounds3::from_bounds, bvh.rs:225).
__global__ void leaf_bounds_kernel(Build b, int cur, uint32_t node_count) {
    const uint32_t id = blockIdx.x * blockDim.x + threadIdx.x;
    if (id >= node_count || b.nodes.left[id] != kNoNode) return;
    Box x = box_identity();
    const uint32_t begin = b.nodes.begin[id], end = begin + b.nodes.count[id];
    for (uint32_t p = begin; p < end; p++) {
        const float4 lo = b.rmin[cur][p], hi = b.rmax[cur][p];
        x.lo[0] = min_rs(x.lo[0], lo.x); x.lo[1] = min_rs(x.lo[1], lo.y); x.lo[2] = min_rs(x.lo[2], lo.z);
        x.hi[0] = max_rs(x.hi[0], hi.x); x.hi[1] = max_rs(x.hi[1], hi.y); x.hi[2] = max_rs(x.hi[2], hi.z);
    }
    b.nodes.bmin[id] = make_float4(x.lo[0], x.lo[1], x.lo[2], 0.0f);
    b.nodes.bmax[id] = make_float4(x.hi[0], x.hi[1], x.hi[2], 0.0f);
}

// Interior nodes of one level, children first: bounds = from_bounds([left, right]) (bvh.rs:197-203).
__global__ void up_kernel(Build b, uint32_t first, uint32_t last) {
    const uint32_t id = first + blockIdx.x * blockDim.x + threadIdx.x;
    if (id >= last) return;
    const uint32_t l = b.nodes.left[id];
    if (l == kNoNode) return;
    const uint32_t r = l + 1u;
    const float4 llo = b.nodes.bmin[l], lhi = b.nodes.bmax[l], rlo = b.nodes.bmin[r], rhi = b.nodes.bmax[r];
    b.nodes.bmin[id] = make_float4(min_rs(min_rs(kMax, llo.x), rlo.x), min_rs(min_rs(kMax, llo.y), rlo.y),
                                   min_rs(min_rs(kMax, llo.z), rlo.z), 0.0f);
    b.nodes.bmax[id] = make_float4(max_rs(max_rs(-kMax, lhi.x), rhi.x), max_rs(max_rs(-kMax, lhi.y), rhi.y),
                                   max_rs(max_rs(-kMax, lhi.z), rhi.z), 0.0f);
    b.nodes.size[id] = 1u + b.nodes.size[l] + b.nodes.size[r];
    const uint32_t dl = b.nodes.depth[l], dr = b.nodes.depth[r];
    b.nodes.depth[id] = (dl > dr ? dl : dr) + 1u;
}

// Pre-order positions, parents first (flatten, bvh.rs:155-178): left = parent + 1, right after the left subtree.
__global__ void down_kernel(Build b, uint32_t first, uint32_t last) {
    const uint32_t id = first + blockIdx.x * blockDim.x + threadIdx.x;
    if (id >= last) return;
    if (id == 0u) b.nodes.pre[0] = 0u;
    const uint32_t l = b.nodes.left[id];
    if (l == kNoNode) return;
    const uint32_t at = b.nodes.pre[id];
    b.nodes.pre[l] = at + 1u;
    b.nodes.pre[l + 1u] = at + 1u + b.nodes.size[l];
}

__global__ void emit_nodes_kernel(Build b, uint32_t node_count, RtBvhNode* out) {
    const uint32_t id = blockIdx.x * blockDim.x + threadIdx.x;
    if (id >= node_count) return;
    const uint32_t l = b.nodes.left[id];
    const float4 lo = b.nodes.bmin[id], hi = b.nodes.bmax[id];
    RtBvhNode n;
    n.min[0] = lo.x; n.min[1] = lo.y; n.min[2] = lo.z; n._pad0 = 0.0f;
    n.max[0] = hi.x; n.max[1] = hi.y; n.max[2] = hi.z; n._pad1 = 0.0f;
    n.primitives_or_second_child_index = l == kNoNode ? b.nodes.begin[id] : b.nodes.pre[l + 1u];
    n.primitives_len = l == kNoNode ? b.nodes.count[id] : 0u;
    n.split_axis = b.nodes.axis[id];
    n._pad2 = 0u;
    out[b.nodes.pre[id]] = n;
}

__global__ void emit_primitives_kernel(Build b, int cur, RtPrimitiveInfo* out) {
    const uint32_t p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= b.n) return;
    RtPrimitiveInfo info;
    info.primitive_type = __float_as_uint(b.rmin[cur][p].w);
    info.index = __float_as_uint(b.rmax[cur][p].w);
    out[p] = info;
}

}  // namespace rtb
