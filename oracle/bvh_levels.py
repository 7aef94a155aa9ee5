"""bvh_levels.py — numpy model of the LEVEL-SYNCHRONOUS BVH build the device builder runs
(rsoderh-raytracing_b200/csrc/rt_bvh_build.cuh). TEST INFRASTRUCTURE ONLY: only tests/ may import it.

It restates, pass for pass, what the CUDA kernels do — bucket accumulation over all primitives
of a level, one split decision per candidate node, the closed form of the reference's swap
partition (src/bvh.rs:304-315), sizes bottom-up, pre-order positions top-down — so the CPU suite
can check the ALGORITHM against the recursive host builder (host/bvh.cpp, itself checked against
scene_oracle.build_bvh and the goldens) byte for byte without a GPU. The GPU suite then checks the
kernels against the host builder directly.
"""
from __future__ import annotations

import numpy as np

f32 = np.float32
F32_MAX = f32(3.4028234663852886e38)
BUCKETS = 12
LEAF_MAX = 5
NO_NODE = 0xFFFFFFFF
SPLIT_LEAF = 0xFF


def partition_literal(flags):
    """The reference's loop (bvh.rs:304-315) on positions 0..n-1; returns (order, split): order[k]
    is the original position of the element that ends up at k."""
    a = list(range(len(flags)))
    split, end = 0, len(a)
    while split < end:
        if flags[a[split]]:
            split += 1
        else:
            end -= 1
            a[split], a[end] = a[end], a[split]
    return a, split


def partition_dest(flags):
    """Closed form of the same permutation (partition_dest in rt_bvh_build.cuh): dest[p] is where
    the element at p ends up."""
    good = np.asarray(flags, dtype=bool)
    n = len(good)
    g = int(good.sum())
    prefix = np.concatenate([[0], np.cumsum(good)]).astype(np.int64)   # goods before p
    pos = np.arange(n)
    front = pos < g
    front_rejects = g - int(prefix[g])
    frpos = np.zeros(n, dtype=np.int64)
    bgpos = np.zeros(n, dtype=np.int64)
    fr = front & ~good
    frpos[pos[fr] - prefix[:-1][fr]] = pos[fr]
    bg = ~front & good
    bgpos[g - prefix[:-1][bg] - 1] = pos[bg]
    dest = pos.copy()
    m = pos[fr] - prefix[:-1][fr]
    dest[fr] = np.where(m == 0, n - 1, bgpos[np.maximum(m - 1, 0)] - 1)
    dest[bg] = frpos[g - prefix[:-1][bg] - 1]
    br = ~front & ~good
    dest[br] = pos[br] - 1
    if g < n and not good[g]:
        dest[g] = bgpos[front_rejects - 1] - 1 if front_rejects else n - 1
    return dest, g


def _fold_min(values):
    """f32::min fold from +MAX: the first of equal values stays (so does the sign of a zero)."""
    acc = F32_MAX
    for v in values:
        if v < acc:
            acc = v
    return acc


def _fold_max(values):
    acc = -F32_MAX
    for v in values:
        if v > acc:
            acc = v
    return acc


def _area(lo, hi):
    d = hi - lo
    return f32(2.0) * (d[0] * d[1] + d[0] * d[2] + d[1] * d[2])


def _max_axis(lo, hi):
    d = hi - lo
    if d[2] > d[0] and d[2] > d[1]:
        return 2
    if d[1] > d[0]:
        return 1
    return 0


def build(lo: np.ndarray, hi: np.ndarray, types: np.ndarray, indices: np.ndarray):
    """lo / hi [N,3] f32 primitive boxes in build_array order. Returns (nodes [M,12] u32 words of
    RtBvhNode, primitives [N,2] u32, depth)."""
    n = len(lo)
    assert n > 0
    lo, hi = lo.astype(f32).copy(), hi.astype(f32).copy()
    types, indices = types.astype(np.uint32).copy(), indices.astype(np.uint32).copy()
    # node arrays (creation order)
    N = dict(begin=[], count=[], left=[], axis=[], split=[], good=[], bmin=[], bmax=[], cmin=[], cmax=[])
    node_of = np.full(n, NO_NODE, dtype=np.int64)
    active = []

    def node_create(begin, count, blo, bhi, clo, chi):
        nid = len(N["begin"])
        leaf = count <= LEAF_MAX
        axis = 0
        if not leaf:
            axis = _max_axis(clo, chi)
            leaf = clo[axis] == chi[axis]
        N["begin"].append(begin); N["count"].append(count); N["left"].append(NO_NODE)
        N["axis"].append(0 if leaf else axis); N["split"].append(SPLIT_LEAF if leaf else 0); N["good"].append(0)
        N["bmin"].append(blo); N["bmax"].append(bhi); N["cmin"].append(clo); N["cmax"].append(chi)
        if not leaf:
            active.append(nid)
        return nid

    with np.errstate(over="ignore", invalid="ignore", divide="ignore"):
        cen = lo * f32(0.5) + hi * f32(0.5)
        node_create(0, n, lo.min(axis=0), hi.max(axis=0), cen.min(axis=0), cen.max(axis=0))
        if active:
            node_of[:] = 0
        level_start = [0, 1]
        while active:
            cur_active, active = active, []
            slot_of = {nid: i for i, nid in enumerate(cur_active)}
            cen = lo * f32(0.5) + hi * f32(0.5)
            # ---- bucket pass over every position that belongs to a candidate node ----
            on = node_of != NO_NODE
            pos = np.nonzero(on)[0]
            nodes_p = node_of[pos]
            axis_p = np.array(N["axis"], dtype=np.int64)[nodes_p]
            cmin_p = np.array(N["cmin"], dtype=f32)[nodes_p, axis_p]
            cmax_p = np.array(N["cmax"], dtype=f32)[nodes_p, axis_p]
            c_p = cen[pos, axis_p]
            scaled = f32(BUCKETS) * ((c_p - cmin_p) / (cmax_p - cmin_p))
            bucket_p = np.where(scaled > 0, scaled, f32(0)).astype(np.int64)   # `as usize`
            bucket_p = np.minimum(bucket_p, BUCKETS - 1)
            bucket = np.full(n, 0xFF, dtype=np.int64)
            bucket[pos] = bucket_p
            slots = len(cur_active)
            slot_p = np.array([slot_of[x] for x in nodes_p], dtype=np.int64)
            key = slot_p * BUCKETS + bucket_p
            counts = np.zeros(slots * BUCKETS, dtype=np.int64)
            np.add.at(counts, key, 1)
            blo = np.full((slots * BUCKETS, 3), F32_MAX, dtype=f32)
            bhi = np.full((slots * BUCKETS, 3), -F32_MAX, dtype=f32)
            clo = np.full((slots * BUCKETS, 3), F32_MAX, dtype=f32)
            chi = np.full((slots * BUCKETS, 3), -F32_MAX, dtype=f32)
            np.minimum.at(blo, key, lo[pos]); np.maximum.at(bhi, key, hi[pos])
            np.minimum.at(clo, key, cen[pos]); np.maximum.at(chi, key, cen[pos])
            # ---- split pass: one decision per candidate ----
            for nid in cur_active:
                s = slot_of[nid] * BUCKETS
                total = _area(N["bmin"][nid], N["bmax"][nid])
                best, best_cost = 0, None
                for split in range(BUCKETS - 1):
                    b0l = np.full(3, F32_MAX, f32); b0h = np.full(3, -F32_MAX, f32)
                    b1l = np.full(3, F32_MAX, f32); b1h = np.full(3, -F32_MAX, f32)
                    c0 = c1 = 0
                    for k in range(split + 1):
                        b0l = np.minimum(b0l, blo[s + k]); b0h = np.maximum(b0h, bhi[s + k]); c0 += counts[s + k]
                    for k in range(split + 1, BUCKETS):
                        b1l = np.minimum(b1l, blo[s + k]); b1h = np.maximum(b1h, bhi[s + k]); c1 += counts[s + k]
                    cost = f32(0.125) + (f32(c0) * _area(b0l, b0h) + f32(c1) * _area(b1l, b1h)) / total
                    if cost != cost:
                        raise ValueError("bounds have non-zero surface area")
                    if best_cost is None or cost < best_cost:
                        best, best_cost = split, cost
                good = int(counts[s:s + best + 1].sum())
                count = N["count"][nid]
                assert 0 < good < count, "BVH partition produced an empty side"
                used_l = [k for k in range(best + 1) if counts[s + k]]
                used_r = [k for k in range(best + 1, BUCKETS) if counts[s + k]]
                begin = N["begin"][nid]
                left = node_create(begin, good, blo[[s + k for k in used_l]].min(axis=0), bhi[[s + k for k in used_l]].max(axis=0),
                                   clo[[s + k for k in used_l]].min(axis=0), chi[[s + k for k in used_l]].max(axis=0))
                node_create(begin + good, count - good, blo[[s + k for k in used_r]].min(axis=0),
                            bhi[[s + k for k in used_r]].max(axis=0), clo[[s + k for k in used_r]].min(axis=0),
                            chi[[s + k for k in used_r]].max(axis=0))
                N["left"][nid] = left
                N["split"][nid] = best
                N["good"][nid] = good
            # ---- flags, prefix, scatter tables, move: all positions at once ----
            split_arr = np.array(N["split"], dtype=np.int64)
            flag = np.zeros(n, dtype=np.int64)
            flag[pos] = (bucket_p <= split_arr[nodes_p]).astype(np.int64)
            prefix = np.concatenate([[0], np.cumsum(flag)])
            begin_arr = np.array(N["begin"], dtype=np.int64)
            count_arr = np.array(N["count"], dtype=np.int64)
            good_arr = np.array(N["good"], dtype=np.int64)
            left_arr = np.array(N["left"], dtype=np.int64)
            b_p, g_p, cnt_p = begin_arr[nodes_p], good_arr[nodes_p], count_arr[nodes_p]
            before = prefix[pos] - prefix[b_p]
            rel = pos - b_p
            is_good = flag[pos] == 1
            frpos = np.zeros(n, dtype=np.int64)
            bgpos = np.zeros(n, dtype=np.int64)
            fr = (rel < g_p) & ~is_good
            frpos[b_p[fr] + rel[fr] - before[fr]] = pos[fr]
            bg = (rel >= g_p) & is_good
            bgpos[b_p[bg] + g_p[bg] - before[bg] - 1] = pos[bg]
            front_rejects = g_p - (prefix[b_p + g_p] - prefix[b_p])
            end_p = b_p + cnt_p
            dest = pos.copy()
            m = rel - before
            dest_fr = np.where(m == 0, end_p - 1, bgpos[b_p + np.maximum(m - 1, 0)] - 1)
            dest_bg = frpos[b_p + np.maximum(g_p - before - 1, 0)]
            dest_t0 = np.where(front_rejects > 0, bgpos[b_p + np.maximum(front_rejects - 1, 0)] - 1, end_p - 1)
            dest = np.where(fr, dest_fr, dest)
            dest = np.where(bg, dest_bg, dest)
            br = (rel >= g_p) & ~is_good
            dest = np.where(br, np.where(rel == g_p, dest_t0, pos - 1), dest)
            assert len(np.unique(dest)) == len(dest)
            child = left_arr[nodes_p] + (dest - b_p >= g_p)
            owner = np.where(split_arr[child] == SPLIT_LEAF, NO_NODE, child)
            new_lo, new_hi, new_t, new_i = lo.copy(), hi.copy(), types.copy(), indices.copy()
            new_lo[dest], new_hi[dest], new_t[dest], new_i[dest] = lo[pos], hi[pos], types[pos], indices[pos]
            node_of[pos] = NO_NODE
            new_owner = node_of.copy()
            new_owner[dest] = owner
            lo, hi, types, indices, node_of = new_lo, new_hi, new_t, new_i, new_owner
            level_start.append(len(N["begin"]))

    # ---- finish ----
    M = len(N["begin"])
    fmin = np.zeros((M, 3), dtype=f32)
    fmax = np.zeros((M, 3), dtype=f32)
    size = np.ones(M, dtype=np.int64)
    depth = np.zeros(M, dtype=np.int64)
    for nid in range(M):
        if N["left"][nid] == NO_NODE:
            b, e = N["begin"][nid], N["begin"][nid] + N["count"][nid]
            for k in range(3):
                fmin[nid, k] = _fold_min(lo[b:e, k])
                fmax[nid, k] = _fold_max(hi[b:e, k])
    for level in range(len(level_start) - 2, -1, -1):
        for nid in range(level_start[level], level_start[level + 1]):
            l = N["left"][nid]
            if l == NO_NODE:
                continue
            r = l + 1
            for k in range(3):
                fmin[nid, k] = _fold_min([fmin[l, k], fmin[r, k]])
                fmax[nid, k] = _fold_max([fmax[l, k], fmax[r, k]])
            size[nid] = 1 + size[l] + size[r]
            depth[nid] = max(depth[l], depth[r]) + 1
    pre = np.zeros(M, dtype=np.int64)
    for level in range(len(level_start) - 1):
        for nid in range(level_start[level], level_start[level + 1]):
            l = N["left"][nid]
            if l == NO_NODE:
                continue
            pre[l] = pre[nid] + 1
            pre[l + 1] = pre[nid] + 1 + size[l]
    out = np.zeros((M, 12), dtype=np.uint32)
    for nid in range(M):
        row = out[pre[nid]]
        row[0:3] = fmin[nid].view(np.uint32)
        row[4:7] = fmax[nid].view(np.uint32)
        l = N["left"][nid]
        row[8] = N["begin"][nid] if l == NO_NODE else pre[l + 1]
        row[9] = N["count"][nid] if l == NO_NODE else 0
        row[10] = N["axis"][nid]
    prims = np.stack([types, indices], axis=1).astype(np.uint32)
    return out, prims, int(depth[0])
