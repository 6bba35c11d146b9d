"""N > 1 host logic on CPU: world_size-2 gloo.  Each rank takes its contiguous shard, computes it (the
CPU oracle stands in for the device here -- this tests the sharding / gather plumbing, not the kernels),
and the gathered statistics and concatenated shards must equal the single-process result."""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n_total, out_dir):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    import torch
    import torch.distributed as dist
    import oracle_lib
    import bench
    from simplediffeq_b200 import jl_range
    from simplediffeq_b200.sharding import shard_bounds, endpoint_stats, gather_endpoint_stats, reduce_max
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lo, hi = shard_bounds(n_total, world, rank)
    u0, p = bench.lorenz_inputs_np(lo, hi, n_total)
    r = oracle_lib.solve("lorenz", "Tsit5", u0.T, p.T, 0.0, 1.0, 0.01, tgrid=jl_range(0.0, 0.01, 1.0))
    u = torch.from_numpy(np.ascontiguousarray(r.u[:, 0, :].T))
    stats = gather_endpoint_stats(endpoint_stats(u), dist)
    tmax = reduce_max(10.0 + rank, dist)
    dist.barrier()
    np.save(os.path.join(out_dir, "u_%d.npy" % rank), u.numpy())
    if rank == 0:
        np.save(os.path.join(out_dir, "stats.npy"), np.stack([stats["mean"].numpy(), stats["min"].numpy(), stats["max"].numpy()]))
        np.save(os.path.join(out_dir, "tmax.npy"), np.array([tmax, float(stats["count"][0])]))
    dist.destroy_process_group()


def test_shard_bounds_cover_everything_once(sde):
    from simplediffeq_b200.sharding import shard_bounds
    for n in (0, 1, 7, 10_000_000, 10_000_001):
        for world in (1, 2, 4, 8):
            b = [shard_bounds(n, world, r) for r in range(world)]
            assert b[0][0] == 0 and b[-1][1] == n
            assert all(b[i][1] == b[i + 1][0] for i in range(world - 1))
            sizes = [hi - lo for lo, hi in b]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        shard_bounds(10, 2, 2)


def test_two_rank_gloo_sharded_solve_matches_single_process(sde, oracle, tmp_path):
    import torch.multiprocessing as mp
    n_total, world = 1001, 2
    port = _free_port()
    mp.start_processes(_worker, args=(world, port, n_total, str(tmp_path)), nprocs=world, join=True, start_method="spawn")
    import bench
    u0, p = bench.lorenz_inputs_np(0, n_total, n_total)
    full = oracle.solve("lorenz", "Tsit5", u0.T, p.T, 0.0, 1.0, 0.01, tgrid=sde.jl_range(0.0, 0.01, 1.0)).u[:, 0, :].T
    got = np.concatenate([np.load(tmp_path / ("u_%d.npy" % r)) for r in range(world)], axis=1)
    assert got.tobytes() == np.ascontiguousarray(full).tobytes()
    stats = np.load(tmp_path / "stats.npy")
    np.testing.assert_allclose(stats[0], full.mean(axis=1), rtol=1e-13)
    assert np.array_equal(stats[1], full.min(axis=1)) and np.array_equal(stats[2], full.max(axis=1))
    tmax, count = np.load(tmp_path / "tmax.npy")
    assert tmax == 11.0 and count == n_total
