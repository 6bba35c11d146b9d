"""Development aid: persistent CTAs per SM of the adaptive kernels (SDE_TUNE_CTAS_PER_SM) on ensembles at or below the
number of resident lanes -- does a smaller grid (more trajectories per lane, better balance inside a warp) beat full occupancy?"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import simplediffeq_b200 as S
dev = torch.device("cuda:0")
dt0 = float(np.float32(0.1))


def problem(kind, n, shuffled):
    idx = torch.arange(n, dtype=torch.int64)
    if shuffled: idx = (idx * 2654435761) % n
    if kind == "vdp":
        u0 = torch.zeros(2, n, dtype=torch.float64); u0[0] = 2
        p = (0.1 + 49.9 * idx.double() / (n - 1)).reshape(1, n)
        return S.systems.vanderpol, u0.to(dev), p.contiguous().to(dev), (0.0, 20.0)
    u0 = torch.zeros(3, n, dtype=torch.float64); u0[0] = 1
    p = torch.empty(3, n, dtype=torch.float64); p[0] = 10; p[1] = 21.0 * idx.double() / (n - 1); p[2] = 8.0 / 3.0
    return S.systems.lorenz, u0.to(dev), p.to(dev), (0.0, 10.0)


def timed(fn, reps=5):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    best = 1e30
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best


for kind, alg, tol in (("vdp", S.GPUSimpleATsit5(), 1e-6), ("lorenz", S.GPUSimpleATsit5(), 1e-8), ("lorenz", S.GPUSimpleAVern9(), 1e-12)):
    for shuffled in (True, False):
        for n in (1 << 16, 1 << 17, 1 << 18, 1 << 19):
            sysm, u0, p, tspan = problem(kind, n, shuffled)
            row = []
            for cap in (0, 6, 4, 3, 2):
                if cap: os.environ["SDE_TUNE_CTAS_PER_SM"] = str(cap)
                else: os.environ.pop("SDE_TUNE_CTAS_PER_SM", None)
                row.append(timed(lambda: S.solve_device(sysm, alg, u0, p, tspan, dt=dt0, abstol=tol, reltol=tol, sync=False, stats=False)))
            print("%-6s %-16s %-8s n=2^%d  CTAs/SM full/6/4/3/2: %s ms" % (kind, type(alg).__name__, "shuffled" if shuffled else "sorted", int(np.log2(n)),
                  "  ".join("%.3f" % x for x in row)), flush=True)
os.environ.pop("SDE_TUNE_CTAS_PER_SM", None)
