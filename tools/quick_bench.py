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

LORENZ_SRC = """
__device__ void rhs(real* du, const real* u, const real* p, real t) {
  du[0] = p[0] * (u[1] - u[0]);
  du[1] = u[0] * (p[1] - u[2]) - u[1];
  du[2] = u[0] * u[1] - p[2] * u[2];
}"""


def probe_saveat(n, dt, layout, dtype=torch.float64, reps=2, alg=None, user=False):
    alg = alg or S.GPUSimpleTsit5()
    sysm = S.CudaRHS(LORENZ_SRC, 3, 3) if user else S.systems.lorenz
    dev = torch.device("cuda:0")
    u0 = torch.zeros(3, n, dtype=dtype, device=dev); u0[0] = 1
    p = torch.empty(3, n, dtype=dtype, device=dev); p[0] = 10; p[2] = 8.0/3.0
    p[1] = 21.0 * torch.arange(n, dtype=torch.float64, device=dev) / max(n - 1, 1)
    saveat = S.jl_range(0.0, 0.01, 10.0)
    shape = (n, 1001, 3) if layout == 0 else (1001, 3, n)
    out = torch.empty(shape, dtype=dtype, device=dev)
    times = []
    for r in range(reps + 1):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        res = S.solve_device(sysm, alg, u0, p, (0.0, 10.0), dt=dt, saveat=saveat, save_mode=1, layout=layout, out=out, stats=False, sync=False)
        e1.record(); torch.cuda.synchronize()
        times.append(e0.elapsed_time(e1))
    ms = min(times[1:])
    nbytes = n * (48 + 1001 * 3 * out.element_size())
    print("saveat %-14s %s layout=%s n=%d dt=%g: ms=%.3f  GB/s=%.1f  steps/s=%.4g" % (type(alg).__name__, str(dtype)[6:], "traj-major(staged)" if layout == 0 else "soa", n, dt, ms, nbytes / ms / 1e6, n * res["n_steps"] / ms * 1e3), flush=True)
    return out


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "tune":
        print("SDE_TUNE_STAGE_ELEMS =", os.environ.get("SDE_TUNE_STAGE_ELEMS"))
        probe_saveat(1 << 21, 0.1, 0, user=True)
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == "saveat":
        print(torch.cuda.get_device_name(0))
        for dt in (0.1, 0.01):
            for layout in (1, 0):
                probe_saveat(1 << 20, dt, layout)
        a = probe_saveat(1 << 18, 0.1, 0); b = probe_saveat(1 << 18, 0.1, 1)
        print("layouts agree:", torch.equal(a, b.permute(2, 0, 1)))
        probe_saveat(4_000_000, 0.1, 0, reps=1)
        probe_saveat(4_000_000, 0.1, 1, reps=1)
        probe_saveat(1 << 20, 0.1, 0, dtype=torch.float32)
        probe_saveat(1 << 20, 0.1, 1, dtype=torch.float32)
        sys.exit(0)
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
