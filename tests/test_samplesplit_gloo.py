"""Multi-rank sample split on CPU: world_size 2 over gloo (SURVEY §8e host logic).

Each rank renders its sample range of the same frame (with the CPU oracle standing in for
the GPU kernels), the accumulation buffers are sum-reduced onto rank 0 through
rsoderh_raytracing_b200.samplesplit.reduce_accumulation, and rank 0 compares with the
single-rank render: equal up to float re-association, alpha restored to 1.
"""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import ASSETS, ROOT


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, spp, out_path):
    for p in (ROOT, os.path.join(ROOT, "oracle")):
        if p not in sys.path:
            sys.path.insert(0, p)
    import _pkg
    rt = _pkg.load()
    import oracle as orc
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        scene = rt.Scene.load_toml(os.path.join(ASSETS, "scenes", "house.toml"))
        env = rt.EnvironmentMaps([rt.Environment.synthetic(128, 64)])
        w, h = 64, 36
        ora = orc.Oracle(scene, env, w, h, max_bounces=6)
        begin, count = rt.samplesplit.split_range(0, spp, rank, world)
        local, counters = ora.render(begin, count, threads=2)
        accum = torch.from_numpy(local)
        rt.samplesplit.reduce_accumulation(accum, dst=0)
        rays = torch.tensor([counters["rays"] + counters["shadow_rays"]], dtype=torch.int64)
        dist.all_reduce(rays)
        if rank == 0:
            full, fc = ora.render(0, spp, threads=2)
            np.savez(out_path, reduced=accum.numpy(), full=full, rays=rays.numpy(),
                     full_rays=fc["rays"] + fc["shadow_rays"])
    finally:
        dist.destroy_process_group()


def test_split_ranges_partition_exactly(rt):
    ss = rt.samplesplit
    for total, world in [(256, 8), (10, 3), (7, 8), (1024, 4), (1, 2)]:
        covered = []
        for r in range(world):
            b, c = ss.split_range(5, total, r, world)
            covered += list(range(b, b + c))
        assert covered == list(range(5, 5 + total))
    assert ss.weak_range(256, 3) == (768, 256)
    with pytest.raises(ValueError):
        ss.split_range(0, 8, 2, 2)


def test_two_rank_sample_split_matches_single_rank(tmp_path):
    world, spp = 2, 6
    out = str(tmp_path / "result.npz")
    mp.spawn(_worker, args=(world, _free_port(), spp, out), nprocs=world, join=True)
    data = np.load(out)
    reduced, full = data["reduced"], data["full"]
    assert (reduced[..., 3] == 1.0).all()
    scale = np.maximum(1.0, np.abs(full[..., :3]))
    assert (np.abs(reduced[..., :3] - full[..., :3]) / scale).max() <= 1e-5 * spp
    assert not np.array_equal(reduced, np.zeros_like(reduced))
    # Disjoint seed ranges: the ranks traced exactly the rays of the single-rank render.
    assert int(data["rays"][0]) == int(data["full_rays"])
