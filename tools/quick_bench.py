"""Ad-hoc device-timed throughput probe (development aid; the contract benchmark is bench.py)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import simplediffeq_b200 as S

def probe(system, alg, n, tspan, dt, dtype=torch.float64, reps=3, **kw):
    sysm = getattr(S.systems, system)
    dev = torch.device("cuda:0")
    if system == "lorenz":
        u0 = torch.zeros(3, n, dtype=dtype, device=dev); u0[0] = 1
        p = torch.empty(3, n, dtype=dtype, device=dev); p[0] = 10; p[2] = 8.0/3.0
        p[1] = 21.0 * torch.arange(n, dtype=torch.float64, device=dev) / max(n - 1, 1)
    elif system == "vanderpol":
        u0 = torch.zeros(2, n, dtype=dtype, device=dev); u0[0] = 2
        p = (0.1 + 49.9 * torch.arange(n, dtype=torch.float64, device=dev) / max(n - 1, 1)).to(dtype).reshape(1, n).contiguous()
    times = []
    for r in range(reps + 1):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        out = S.solve_device(sysm, alg, u0, p, tspan, dt=dt, sync=False, **kw)
        e1.record(); torch.cuda.synchronize()
        times.append(e0.elapsed_time(e1))
    ms = min(times[1:])
    if alg.adaptive:
        steps = int(out["naccept"].sum().item()); att = steps + int(out["nreject"].sum().item())
    else:
        steps = n * out["n_steps"]; att = steps
    print("%-12s %-16s %-8s n=%-9d ms=%9.3f  steps/s=%.4g attempts/s=%.4g" % (system, type(alg).__name__, str(dtype)[6:], n, ms, steps / ms * 1e3, att / ms * 1e3), flush=True)
    return out

if __name__ == "__main__":
    print(torch.cuda.get_device_name(0))
    for n in (1 << 18, 1 << 20, 1 << 22):
        probe("lorenz", S.GPUSimpleTsit5(), n, (0.0, 10.0), 1e-3)
    probe("lorenz", S.GPUSimpleTsit5(), 10_000_000, (0.0, 10.0), 1e-3, reps=2)
    probe("lorenz", S.GPUSimpleTsit5(), 1 << 22, (0.0, 10.0), 1e-3, dtype=torch.float32)
    probe("lorenz", S.GPUSimpleRK4(), 1 << 20, (0.0, 10.0), 1e-3)
    probe("lorenz", S.GPUSimpleVern7(), 1 << 20, (0.0, 10.0), 1e-3)
    probe("lorenz", S.GPUSimpleVern9(), 1 << 20, (0.0, 10.0), 1e-3)
    probe("lorenz", S.GPUSimpleATsit5(), 1 << 20, (0.0, 10.0), float(np.float32(0.1)), abstol=1e-8, reltol=1e-8)
    probe("vanderpol", S.GPUSimpleATsit5(), 1 << 20, (0.0, 20.0), float(np.float32(0.1)), abstol=1e-6, reltol=1e-6)
    probe("lorenz", S.GPUSimpleAVern9(), 1 << 20, (0.0, 10.0), float(np.float32(0.1)), abstol=1e-12, reltol=1e-12)
