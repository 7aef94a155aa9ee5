"""Sample-split partitioning across ranks and the accumulation reduce.

Every pixel-sample's RNG seed depends only on (pixel_index, sample_index)
(shader.wgsl:1309-1312), so disjoint sample-index ranges rendered on different GPUs and
summed give the single-GPU image up to float re-association. One process per GPU;
torch.distributed provides the plumbing (NCCL on GPUs, gloo in CPU tests).
"""
from __future__ import annotations


def split_range(sample_begin: int, sample_count: int, rank: int, world: int):
    """Contiguous, disjoint ranges whose union is [sample_begin, sample_begin + count):
    the same arithmetic as rt_group_render() in the C ABI. Returns (begin, count)."""
    if not (0 <= rank < world):
        raise ValueError("rank out of range")
    lo = sample_count * rank // world
    hi = sample_count * (rank + 1) // world
    return sample_begin + lo, hi - lo


def weak_range(spp_per_rank: int, rank: int):
    """Fixed work per rank (bench.py): rank r renders samples [r*spp, (r+1)*spp)."""
    return rank * spp_per_rank, spp_per_rank


def reduce_accumulation(accum, dst: int = 0):
    """Sum-reduce the W*H*4 accumulation tensors onto `dst` and restore alpha = 1
    (every rank wrote alpha 1, shader.wgsl:1371, so the sum holds the world size)."""
    import torch.distributed as dist
    if dist.is_initialized() and dist.get_world_size() > 1:
        dist.reduce(accum, dst=dst, op=dist.ReduceOp.SUM)
        if dist.get_rank() == dst:
            accum[..., 3] = 1.0
    return accum
