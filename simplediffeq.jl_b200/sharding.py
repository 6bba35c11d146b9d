"""Multi-GPU plumbing: trajectories shard by contiguous index range, one process (or host thread)
per device, NO collective on the data path (each trajectory depends only on its own u0, p).
torch.distributed is used only for the barrier / max-over-ranks timing of the benchmark and for the
optional final gather of endpoint statistics (NCCL on GPUs, gloo in the CPU tests)."""


def shard_bounds(n_total, world, rank):
    """[lo, hi) of rank's contiguous index range: floor(g*N/G) .. floor((g+1)*N/G)  (same rule as
    sde_solve's in-library sharder)."""
    if not 0 <= rank < world:
        raise ValueError("rank out of range")
    return n_total * rank // world, n_total * (rank + 1) // world


def pilot_indices(n_total, k=4096):
    """k evenly spaced global trajectory indices (first and last included): the pilot sample whose step counts
    estimate how the cost of an adaptive solve varies along the index (same rule on every rank, no exchange)."""
    import numpy as np
    k = int(min(max(k, 2), n_total)) if n_total > 1 else n_total
    if k <= 1:
        return np.zeros(k, dtype=np.int64)
    return (np.arange(k, dtype=np.int64) * (n_total - 1)) // (k - 1)


def cost_weighted_bounds(n_total, world, sample_idx, sample_cost):
    """Contiguous split of [0, n_total) into `world` ranges of (estimated) EQUAL COST instead of equal count, for
    adaptive solves whose step counts vary along the index (a sorted Van der Pol sweep costs 8x more at
    one end than at the other, so equal index ranges leave most GPUs idle while the last one works).

    sample_idx / sample_cost: sorted global indices of a pilot sample (pilot_indices) and their cost (accepted +
    rejected attempts of an endpoint-only solve).  The cost density is taken piecewise linear between the sample
    points, integrated (trapezoids), and the cumulative cost inverted at g/world.  Returns world+1 non-decreasing
    integer bounds with bounds[0] = 0 and bounds[-1] = n_total; rank g owns [bounds[g], bounds[g+1]).  Integer-valued
    inputs and a fixed operation order: every rank that evaluates this on the same pilot gets the same bounds."""
    import numpy as np
    if world < 1:
        raise ValueError("world must be >= 1")
    idx = np.asarray(sample_idx, dtype=np.float64)
    cost = np.asarray(sample_cost, dtype=np.float64)
    if idx.shape != cost.shape or idx.ndim != 1:
        raise ValueError("sample_idx and sample_cost must be 1-d arrays of the same length")
    if n_total <= 0 or len(idx) == 0:
        return [n_total * g // world for g in range(world + 1)]
    if np.any(np.diff(idx) <= 0) or idx[0] < 0 or idx[-1] > n_total - 1:
        raise ValueError("sample_idx must be strictly increasing global indices")
    cost = np.where(np.isfinite(cost) & (cost > 0), cost, 0.0)
    # knots at trajectory edges: extend the first / last density to the ends of [0, n_total]
    x = np.concatenate([[0.0], idx + 0.5, [float(n_total)]])
    d = np.concatenate([[cost[0]], cost, [cost[-1]]])
    keep = np.concatenate([[True], np.diff(x) > 0])
    x, d = x[keep], d[keep]
    cum = np.concatenate([[0.0], np.cumsum(0.5 * (d[1:] + d[:-1]) * np.diff(x))])
    if not cum[-1] > 0:
        return [n_total * g // world for g in range(world + 1)]
    bounds = [0]
    for g in range(1, world):
        target = cum[-1] * g / world
        j = int(np.searchsorted(cum, target, side="right")) - 1
        j = min(max(j, 0), len(x) - 2)
        # invert the quadratic cum(x) on segment j: density d[j] + s (x - x[j])
        w, a, b = x[j + 1] - x[j], d[j], d[j + 1]
        r = target - cum[j]
        s = (b - a) / w
        if abs(s) * w <= 1e-12 * max(a, b, 1e-300):
            t = r / a if a > 0 else w
        else:
            t = (-a + (a * a + 2.0 * s * r) ** 0.5) / s
        cut = int(round(x[j] + min(max(t, 0.0), w)))
        bounds.append(min(max(cut, bounds[-1]), n_total))
    bounds.append(n_total)
    return bounds


def pilot_weighted_bounds(n_total, world, solve_pilot, k=4096):
    """cost_weighted_bounds from a pilot run: `solve_pilot(idx)` solves the trajectories with the given global
    indices endpoint-only on the caller's device and returns their attempt counts (naccept + nreject).  The step
    sequence of a trajectory is deterministic, so every rank computes the same pilot and no exchange is needed."""
    idx = pilot_indices(n_total, k)
    return cost_weighted_bounds(n_total, world, idx, solve_pilot(idx))


def endpoint_stats(u_soa):
    """Per-component (sum, min, max, count) of the final states of one shard; u_soa: [n_state, n]
    torch tensor.  Sums (not means) so that shards combine exactly."""
    import torch
    n = u_soa.shape[1]
    return torch.stack([u_soa.sum(dim=1), u_soa.amin(dim=1), u_soa.amax(dim=1),
                        torch.full_like(u_soa[:, 0], float(n))])


def gather_endpoint_stats(local, dist=None):
    """all_gather of the per-shard statistics and their combination (mean, min, max) on every rank."""
    import torch
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        parts = [local]
    else:
        parts = [torch.empty_like(local) for _ in range(dist.get_world_size())]
        dist.all_gather(parts, local.contiguous())
    s = torch.stack(parts)                       # [world, 4, n_state]
    count = s[:, 3].sum(dim=0)
    return {"mean": s[:, 0].sum(dim=0) / count, "min": s[:, 1].amin(dim=0), "max": s[:, 2].amax(dim=0),
            "count": count}


def reduce_max(value, dist=None, device="cpu"):
    """max over ranks of a scalar (the benchmark's timing rule)."""
    import torch
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())
