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
