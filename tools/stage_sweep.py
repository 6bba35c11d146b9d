"""Development aid: stage size of the trajectory-major series writer (SDE_TUNE_STAGE_ELEMS, NVRTC systems only)
on BASELINE config 5 (Lorenz, Tsit5, saveat = 0:0.01:10, dt = 0.1), 2 M trajectories."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
import simplediffeq_b200 as S
SRC = """
__device__ void rhs(real* du, const real* u, const real* p, real t) {
  du[0] = p[0] * (u[1] - u[0]);
  du[1] = u[0] * (p[1] - u[2]) - u[1];
  du[2] = u[0] * u[1] - p[2] * u[2];
}"""
dev = torch.device("cuda:0")
n = 2_000_000
u0 = torch.zeros(3, n, dtype=torch.float64, device=dev); u0[0] = 1
p = torch.empty(3, n, dtype=torch.float64, device=dev); p[0] = 10; p[2] = 8.0 / 3.0
p[1] = 21.0 * torch.arange(n, dtype=torch.float64, device=dev) / (n - 1)
saveat = S.jl_range(0.0, 0.01, 10.0)
out = torch.empty((n, 1001, 3), dtype=torch.float64, device=dev)
for elems in [int(x) for x in sys.argv[1:]] or [30, 36, 45, 48, 60]:
    os.environ["SDE_TUNE_STAGE_ELEMS"] = str(elems)
    user = S.CudaRHS(SRC + "\n// stage %d\n" % elems, 3, 3)
    run = lambda: S.solve_device(user, S.GPUSimpleTsit5(), u0, p, (0.0, 10.0), dt=0.1, saveat=saveat, save_mode=1, layout=0, out=out, stats=False, sync=False)
    run(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3): run()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 3
    print("stage %3d elements (%2d slots, %4d B runs): %.2f ms  %.0f GB/s" % (elems, elems // 3, elems // 3 * 24, ms, n * 24072 / ms / 1e6), flush=True)
