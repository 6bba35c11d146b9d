"""Development aid: config 5f (Lorenz, fixed-step Tsit5 dt = 0.01, saveat = 0:0.01:10, SoA) and every-step outputs with and without the
opt-in fast flags (SDE_COMPAT_FAST_RHS | SDE_COMPAT_FAST_STAGES)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import simplediffeq_b200 as S
dev = torch.device("cuda:0")
n = 1 << 20
u0 = torch.zeros(3, n, dtype=torch.float64, device=dev); u0[0] = 1
p = torch.empty(3, n, dtype=torch.float64, device=dev); p[0] = 10; p[1] = 21.0 * torch.arange(n, dtype=torch.float64, device=dev) / (n - 1); p[2] = 8.0 / 3.0
sa = S.jl_range(0.0, 0.01, 10.0)


def timed(fn, reps=5):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    best = 1e30
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best


for name, kw in (("saveat dt=0.01 SoA (config 5f at 2^20)", dict(dt=0.01, saveat=sa, save_mode=1, layout=1)),
                 ("saveat dt=0.01 trajectory-major", dict(dt=0.01, saveat=sa, save_mode=1, layout=0)),
                 ("saveat dt=0.1 SoA (config 5 at 2^20)", dict(dt=0.1, saveat=sa, save_mode=1, layout=1)),
                 ("endpoint dt=0.001", dict(dt=0.001))):
    out = None
    res = []
    for compat in (0, 8, 24):
        r = S.solve_device(S.systems.lorenz, S.GPUSimpleTsit5(), u0, p, (0.0, 10.0), compat=compat, sync=True, **kw)
        out = r["u"] if compat == 0 else out
        dev_max = float((r["u"] - out).abs().max().item()) if compat else 0.0
        buf = r["u"]
        ms = timed(lambda: S.solve_device(S.systems.lorenz, S.GPUSimpleTsit5(), u0, p, (0.0, 10.0), compat=compat, sync=False, out=buf, **kw))
        res.append((compat, ms, dev_max))
        del r
    print("%-42s " % name + "  ".join("compat %2d: %7.3f ms (max |diff| %.2e)" % x for x in res), flush=True)
    del out, buf
