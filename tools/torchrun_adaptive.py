"""One process per GPU (torchrun) on a SORTED adaptive sweep (BASELINE config 3, Van der Pol mu = 0.1..50: the cost
rises ~8x along the index): equal index ranges against the cost-weighted contiguous split of
simplediffeq_b200.sharding (pilot of 4096 trajectories, no exchange).  Device-timed, max over ranks.  Never run so far
(written after round 1's GPU minutes were spent):
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 \\
        tools/torchrun_adaptive.py [log2(n_total), default 23]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402
import simplediffeq_b200 as S  # noqa: E402
from simplediffeq_b200.sharding import shard_bounds, pilot_weighted_bounds, reduce_max  # noqa: E402


def main():
    rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    dev = torch.device("cuda", int(os.environ.get("LOCAL_RANK", 0)))
    torch.cuda.set_device(dev)
    if world > 1:
        fd = os.dup(1); os.dup2(2, 1)                       # NCCL's banner goes to stderr
        dist.init_process_group("nccl", device_id=dev)
        dist.barrier(); os.dup2(fd, 1); os.close(fd)
    n = 1 << (int(sys.argv[1]) if len(sys.argv) > 1 else 23)
    alg, tspan, dt0, tol = S.GPUSimpleATsit5(), (0.0, 20.0), float(np.float32(0.1)), 1e-6

    def inputs(idx):
        m = len(idx)
        u0 = torch.zeros(2, m, dtype=torch.float64, device=dev); u0[0] = 2
        mu = (0.1 + 49.9 * idx.to(torch.float64) / (n - 1)).reshape(1, m).contiguous()
        return u0, mu

    def run(lo, hi):
        u0, mu = inputs(torch.arange(lo, hi, device=dev))
        best, out = None, None
        for rep in range(4):
            if world > 1:
                dist.barrier()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize(); e0.record()
            out = S.solve_device(S.systems.vanderpol, alg, u0, mu, tspan, dt=dt0, abstol=tol, reltol=tol, sync=False)
            e1.record(); torch.cuda.synchronize()
            ms = reduce_max(e0.elapsed_time(e1), dist if world > 1 else None, dev)
            best = ms if rep and (best is None or ms < best) else best
        return best, int(out["naccept"].sum().item())

    def pilot(idx):
        u0, mu = inputs(torch.from_numpy(idx).to(dev))
        o = S.solve_device(S.systems.vanderpol, alg, u0, mu, tspan, dt=dt0, abstol=tol, reltol=tol)
        return (o["naccept"] + o["nreject"]).cpu().numpy()

    for name, (lo, hi) in (("equal index ranges", shard_bounds(n, world, rank)),
                           ("cost-weighted ranges", (lambda b: (b[rank], b[rank + 1]))(pilot_weighted_bounds(n, world, pilot)))):
        ms, acc = run(lo, hi)
        t = torch.tensor([float(acc), float(hi - lo)], dtype=torch.float64, device=dev)
        if world > 1:
            parts = [torch.empty_like(t) for _ in range(world)]
            dist.all_gather(parts, t)
        else:
            parts = [t]
        if rank == 0:
            total = sum(float(q[0]) for q in parts)
            print("%-22s n=2^%d GPUs=%d: %.3f ms  %.4g accepted steps/s  shard sizes %s" % (
                name, int(np.log2(n)), world, ms, total / ms * 1e3, [int(q[1]) for q in parts]), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
