"""Development aid: A/B of fixed-step saveat kernel variants on BASELINE config 5 (Lorenz, Tsit5, dt = 0.1,
saveat = 0:0.01:10).  Each variant is the NVRTC twin of the built-in kernel compiled with extra macro definitions
(SDE_TUNE_DEFINES, SDE_TUNE_STAGE_ELEMS); the built-in kernel runs first for reference.

    python tools/tm_variants.py [n_traj] [layout: 0 trajectory-major | 1 SoA] [dt] ["-DSDE_STEP_UNROLL=2" "stage=64 -DX=1" ...]
"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import simplediffeq_b200 as S
SRC = """
__device__ void rhs(real* du, const real* u, const real* p, real t) {
  du[0] = p[0] * (u[1] - u[0]);
  du[1] = u[0] * (p[1] - u[2]) - u[1];
  du[2] = u[0] * u[1] - p[2] * u[2];
}"""
args = sys.argv[1:]
n = int(args.pop(0)) if args and args[0].isdigit() else 2_000_000
LAYOUT = int(args.pop(0)) if args and args[0] in ("0", "1") else 0      # 0 trajectory-major (staged), 1 SoA
DT = float(args.pop(0)) if args and args[0].replace(".", "").isdigit() else 0.1
variants = args or [""]
dev = torch.device("cuda:0")
u0 = torch.zeros(3, n, dtype=torch.float64, device=dev); u0[0] = 1
p = torch.empty(3, n, dtype=torch.float64, device=dev); p[0] = 10; p[2] = 8.0 / 3.0
p[1] = 21.0 * torch.arange(n, dtype=torch.float64, device=dev) / (n - 1)
saveat = S.jl_range(0.0, 0.01, 10.0)
out = torch.empty((n, 1001, 3) if LAYOUT == 0 else (1001, 3, n), dtype=torch.float64, device=dev)
ref = None


def timed(sysm, label):
    global ref
    run = lambda: S.solve_device(sysm, S.GPUSimpleTsit5(), u0, p, (0.0, 10.0), dt=DT, saveat=saveat, save_mode=1, layout=LAYOUT, out=out, stats=False, sync=False)
    out.zero_()
    run(); torch.cuda.synchronize()
    chk = (out[:: max(1, n // 4096)] if LAYOUT == 0 else out[:, :, :: max(1, n // 4096)]).clone()
    if ref is None: ref = chk
    same = bool(torch.equal(chk.view(torch.int64), ref.view(torch.int64)))
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3): run()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 3
    print("layout %d dt %-5g %-50s %8.2f ms  %6.0f GB/s  %.4g steps/s  bit-identical to built-in: %s"
          % (LAYOUT, DT, label, ms, n * 24072 / ms / 1e6, n * round(10 / DT) / ms * 1e3, same), flush=True)


timed(S.systems.lorenz, "built-in")
for i, v in enumerate(variants):
    words = v.split()
    stage = [w for w in words if w.startswith("stage=")]
    os.environ.pop("SDE_TUNE_STAGE_ELEMS", None)
    if stage: os.environ["SDE_TUNE_STAGE_ELEMS"] = stage[0][6:]
    mb = [w for w in words if w.startswith("minblocks=")]
    os.environ.pop("SDE_TUNE_MIN_BLOCKS", None)
    if mb: os.environ["SDE_TUNE_MIN_BLOCKS"] = mb[0][10:]
    os.environ["SDE_TUNE_DEFINES"] = " ".join(w for w in words if w.startswith("-D"))
    timed(S.CudaRHS(SRC + "\n// variant %d\n" % i, 3, 3), "nvrtc [%s]" % v)
