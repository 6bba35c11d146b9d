"""Development aid: SimpleEM keeping every state (the reference's behaviour), both layouts."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import simplediffeq_b200 as S
dev = torch.device("cuda:0")


def timed(fn, reps=4):
    fn(); fn(); torch.cuda.synchronize()
    best = 1e30
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best


for name, sysm, nstate, npar, dtype in (("gbm f64", S.sde_systems.gbm, 1, 2, torch.float64), ("gbm f32", S.sde_systems.gbm, 1, 2, torch.float32),
                                        ("nondiag2x4 f64", S.sde_systems.nondiag2x4, 2, 1, torch.float64)):
    n, steps = 1 << 22, 255
    u0 = torch.ones((nstate, n), dtype=dtype, device=dev)
    p = torch.empty((npar, n), dtype=dtype, device=dev)
    if npar == 2: p[0] = 0.1; p[1] = 0.2
    else: p[0] = 1.01
    for layout, nm in ((1, "SoA"), (0, "trajectory-major")):
        out = torch.empty((steps + 1, nstate, n) if layout == 1 else (n, steps + 1, nstate), dtype=dtype, device=dev)
        ms = timed(lambda: S.solve_em_device(sysm, u0, p, 0.0, 1 / 256, steps, seed=1, save_mode=2, layout=layout, out=out, sync=False))
        gb = out.numel() * out.element_size() / 1e9
        print("EM %-15s %d paths x %d steps, every state, %-16s: %7.2f ms  %6.0f GB/s  %.3g steps/s" % (name, n, steps, nm, ms, gb / ms * 1e3, n * steps / ms * 1e3), flush=True)
        del out
