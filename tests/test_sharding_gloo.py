"""
CPU tests of the N > 1 host logic with world_size-2 `gloo`: contiguous ray-range sharding and the all-reduce of the
reduced products.  The per-shard partial results are produced by the oracle's NumPy definitions (no GPU here); on
the GPU box the same helpers run over NCCL inside Reducer.allreduce / bench.py.
"""
import os
import socket
import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent


def test_shard_range_partitions_exactly():
    from ray_trace_pb_b200.sharding import shard_range
    for n in (0, 1, 7, 8, 9, 1000, 10**9):
        for world in (1, 2, 3, 4, 8):
            spans = [shard_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0
            for (f0, c0), (f1, _c1) in zip(spans, spans[1:]):
                assert f0 + c0 == f1
            assert spans[-1][0] + spans[-1][1] == n
            counts = [c for _, c in spans]
            assert max(counts) - min(counts) <= 1
    with pytest.raises(ValueError):
        shard_range(10, 2, 2)


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out_dir):
    sys.path.insert(0, str(ROOT))
    sys.path.insert(0, str(ROOT / "tests"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    import torch
    import torch.distributed as dist
    import systems
    import ray_trace_pb_b200.materials as rtm
    import ray_trace_pb_b200.raytrace as rt
    from oracle import oracle
    from ray_trace_pb_b200.sharding import allreduce_grid, allreduce_stats, shard_range

    dist.init_process_group("gloo", rank=rank, world_size=world)
    system, m_in, m_out, _ = systems.plano_convex(rt, rtm)
    mats = [m_in] + system.materials + [m_out]
    n_side = 96
    first, count = shard_range(n_side * n_side, rank, world)
    rays = oracle.source_rays("grid", n_side, n_side, 20.0, [0, 0, -5.0], (0, 0, 1), 0.5, b_max=20.0,
                              first=first, count=count)
    last = oracle.trace(system.surfaces, mats, rays, keep_all=False)
    stats = torch.from_numpy(oracle.reduce_stats(last, (0, 0, 0), (1, 0, 0), (0, 1, 0)))
    grid = torch.from_numpy(oracle.reduce_grid(last, (0, 0, 0), (1, 0, 0), (0, 1, 0), 32, 21.0))
    # the buckets of a sweep: (n, 12), merged row by row (second row: this rank's statistics with the x axis flipped)
    flipped = torch.from_numpy(oracle.reduce_stats(last, (0, 0, 0), (-1, 0, 0), (0, 1, 0)))
    buckets = torch.stack((stats, flipped)).clone()
    allreduce_stats(stats)
    allreduce_stats(buckets)
    allreduce_grid(grid)
    np.savez(Path(out_dir) / f"rank{rank}.npz", stats=stats.numpy(), grid=grid.numpy(), first=first, count=count,
             buckets=buckets.numpy())
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_gloo_allreduce_matches_single_process(tmp_path, rt, rtm, oracle):
    import torch.multiprocessing as mp
    import systems
    world = 2
    port = _free_port()
    mp.spawn(_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    parts = [np.load(tmp_path / f"rank{r}.npz") for r in range(world)]
    assert int(parts[0]["first"]) == 0 and int(parts[1]["first"]) == int(parts[0]["count"])
    assert int(parts[0]["count"]) + int(parts[1]["count"]) == 96 * 96
    # every rank ends with the same, complete answer
    assert np.array_equal(parts[0]["stats"], parts[1]["stats"])
    assert np.array_equal(parts[0]["buckets"][0], parts[0]["stats"])
    b1 = parts[0]["buckets"][1]
    assert b1[0] == parts[0]["stats"][0] and b1[8] == -parts[0]["stats"][9] and b1[9] == -parts[0]["stats"][8]
    assert np.array_equal(parts[0]["grid"], parts[1]["grid"])
    system, m_in, m_out, _ = systems.plano_convex(rt, rtm)
    mats = [m_in] + system.materials + [m_out]
    rays = oracle.source_rays("grid", 96, 96, 20.0, [0, 0, -5.0], (0, 0, 1), 0.5, b_max=20.0)
    last = oracle.trace(system.surfaces, mats, rays, keep_all=False)
    want_stats = oracle.reduce_stats(last, (0, 0, 0), (1, 0, 0), (0, 1, 0))
    want_grid = oracle.reduce_grid(last, (0, 0, 0), (1, 0, 0), (0, 1, 0), 32, 21.0)
    assert parts[0]["stats"][0] == want_stats[0]
    # sums over a symmetric bundle cancel to ~0, so compare against the scale of the summands (|u| <= 20, N ~ 9e3)
    np.testing.assert_allclose(parts[0]["stats"], want_stats, rtol=1e-12, atol=1e-9)
    assert np.array_equal(parts[0]["grid"][2], want_grid[2])
    np.testing.assert_allclose(parts[0]["grid"], want_grid, rtol=0, atol=1e-9)


def test_merge_stats_host_matches_definition():
    from ray_trace_pb_b200.sharding import merge_stats_host
    a = np.array([3, 1, 2, 3, 4, 5, 6, 7, -1.0, 2.0, -3.0, 4.0])
    b = np.array([2, 1, 1, 1, 1, 1, 1, 1, -2.0, 1.0, -1.0, 9.0])
    m = merge_stats_host([a, b])
    assert m[0] == 5 and m[8] == -2.0 and m[9] == 2.0 and m[10] == -3.0 and m[11] == 9.0
