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


def test_cost_weighted_bounds_equalise_a_sorted_sweep(sde):
    """contiguous ranges of equal estimated cost for adaptive sweeps.  Cost model of the sorted Van der Pol
    sweep (config 3): accepted steps rise ~8x along the index."""
    from simplediffeq_b200.sharding import cost_weighted_bounds, pilot_indices, shard_bounds
    n = 1 << 20
    true_cost = lambda i: 89.0 + (708.0 - 89.0) * (i / (n - 1.0)) ** 1.5      # noqa: E731
    idx = pilot_indices(n, 4096)
    assert idx[0] == 0 and idx[-1] == n - 1 and len(idx) == 4096 and np.all(np.diff(idx) > 0)
    allc = true_cost(np.arange(n, dtype=np.float64))
    for world in (1, 2, 4, 8):
        b = cost_weighted_bounds(n, world, idx, np.round(true_cost(idx)))
        assert b[0] == 0 and b[-1] == n and len(b) == world + 1 and all(b[i] <= b[i + 1] for i in range(world))
        per = np.array([allc[b[g]:b[g + 1]].sum() for g in range(world)])
        assert per.max() / per.mean() < 1.005, per / per.mean()          # balanced to 0.5 %
        static = np.array([allc[slice(*shard_bounds(n, world, g))].sum() for g in range(world)])
        if world == 8:
            assert static.max() / static.mean() > 1.7                      # what equal index ranges would cost
    # uniform cost -> (almost) the equal-count split; zero / NaN costs and tiny ensembles fall back gracefully
    b = cost_weighted_bounds(1000, 4, pilot_indices(1000, 64), np.full(64, 7.0))
    assert b == [0, 250, 500, 750, 1000]
    assert cost_weighted_bounds(1000, 4, pilot_indices(1000, 64), np.zeros(64)) == [0, 250, 500, 750, 1000]
    assert cost_weighted_bounds(0, 2, [], []) == [0, 0, 0]
    assert cost_weighted_bounds(1, 2, [0], [5.0]) in ([0, 0, 1], [0, 1, 1])
    b = cost_weighted_bounds(10, 2, [0, 9], [np.nan, 4.0])
    assert b[0] == 0 and b[-1] == 10 and b[1] >= 5
    with pytest.raises(ValueError):
        cost_weighted_bounds(10, 2, [3, 3], [1.0, 1.0])


def _worker_adaptive(rank, world, port, n_total, out_dir):
    """Cost-weighted shards of an adaptive Van der Pol sweep: every rank runs the same pilot (the oracle stands in for the
    device), derives the same bounds without any exchange, solves its range; rank 0 gathers the attempt totals."""
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    import torch
    import torch.distributed as dist
    import oracle_lib
    import common as C
    from simplediffeq_b200.sharding import pilot_weighted_bounds
    dist.init_process_group("gloo", rank=rank, world_size=world)
    u0, p = C.vdp_sweep(n_total)
    dt0 = float(np.float32(0.1))

    def solve(idx):
        return oracle_lib.solve("vanderpol", "ATsit5", u0[idx], p[idx], 0.0, 20.0, dt0, abstol=1e-6, reltol=1e-6)

    def pilot(idx):
        r = solve(idx)
        return r.naccept + r.nreject

    b = pilot_weighted_bounds(n_total, world, pilot, k=48)
    r = solve(np.arange(b[rank], b[rank + 1]))
    mine = torch.tensor([float((r.naccept + r.nreject).sum()), float(b[rank]), float(b[rank + 1])], dtype=torch.float64)
    parts = [torch.empty_like(mine) for _ in range(world)]
    dist.all_gather(parts, mine)
    np.save(os.path.join(out_dir, "a_%d.npy" % rank), np.ascontiguousarray(r.u[:, 0, :]))
    if rank == 0:
        np.save(os.path.join(out_dir, "parts.npy"), torch.stack(parts).numpy())
    dist.destroy_process_group()


def test_two_rank_gloo_cost_weighted_adaptive_shards(sde, oracle, tmp_path):
    import torch.multiprocessing as mp
    import common as C
    n_total, world = 1500, 2
    mp.start_processes(_worker_adaptive, args=(world, _free_port(), n_total, str(tmp_path)), nprocs=world, join=True,
                       start_method="spawn")
    parts = np.load(tmp_path / "parts.npy")                       # [rank] = (attempts, lo, hi)
    assert parts[0][1] == 0 and parts[0][2] == parts[1][1] and parts[1][2] == n_total     # same bounds on both ranks
    assert parts[0][2] > 0.55 * n_total                           # the cheap end gets more trajectories ...
    assert abs(parts[0][0] - parts[1][0]) <= 0.03 * parts[:, 0].sum()      # ... and the same work
    u0, p = C.vdp_sweep(n_total)
    full = oracle.solve("vanderpol", "ATsit5", u0, p, 0.0, 20.0, float(np.float32(0.1)), abstol=1e-6, reltol=1e-6)
    got = np.concatenate([np.load(tmp_path / ("a_%d.npy" % r)) for r in range(world)], axis=0)
    assert got.tobytes() == np.ascontiguousarray(full.u[:, 0, :]).tobytes()
    half = (full.naccept + full.nreject)[:n_total // 2].sum() / (full.naccept + full.nreject).sum()
    assert half < 0.42                                            # equal index ranges would be this unbalanced


def test_cost_weighted_bounds_properties_random_profiles(sde):
    """Random positive cost profiles (smooth, stepped, with zero stretches): bounds are monotone, cover [0, n) once,
    and no rank's true cost exceeds the mean by more than the profile's resolution allows."""
    from simplediffeq_b200.sharding import cost_weighted_bounds, pilot_indices
    rng = np.random.default_rng(8)
    for trial in range(40):
        n = int(rng.integers(1, 200_000))
        world = int(rng.choice([1, 2, 3, 4, 8]))
        k = int(rng.choice([2, 16, 256, 4096]))
        i = np.arange(n, dtype=np.float64) / max(n - 1, 1)
        kind = trial % 3
        if kind == 0:
            cost = 50 + 400 * i ** rng.uniform(0.3, 3.0)
        elif kind == 1:
            cost = 100 + 80 * np.sin(2 * np.pi * rng.uniform(0.5, 2.0) * i) + 300 * (i > rng.uniform(0.2, 0.8))
        else:
            cost = np.where(i < rng.uniform(0.1, 0.5), 0.0, 200.0 * i)
        idx = pilot_indices(n, k)
        b = cost_weighted_bounds(n, world, idx, np.round(cost[idx]))
        assert len(b) == world + 1 and b[0] == 0 and b[-1] == n
        assert all(0 <= b[g] <= b[g + 1] <= n for g in range(world))
        if kind == 0 and n >= 50_000 and k >= 256:
            per = np.array([cost[b[g]:b[g + 1]].sum() for g in range(world)])
            assert per.max() <= 1.02 * per.mean(), (n, world, k, per / per.mean())
