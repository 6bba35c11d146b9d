"""Development aid: host-buffer (sde_solve, pinned memory) vs device-resident timing of an adaptive sweep."""
import sys, os, time, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import simplediffeq_b200 as S
from simplediffeq_b200 import _lib
n = 1 << (int(sys.argv[1]) if len(sys.argv) > 1 else 20)
SORTED = len(sys.argv) > 2 and sys.argv[2] == "sorted"
ndev = _lib.device_count()
u0 = torch.zeros((2, n), dtype=torch.float64).pin_memory(); u0[0] = 2
idx = torch.arange(n, dtype=torch.int64) if SORTED else (torch.arange(n, dtype=torch.int64) * 2654435761) % n
p = (0.1 + 49.9 * idx.to(torch.float64) / (n - 1)).reshape(1, n).contiguous().pin_memory()
out = torch.empty((2, n), dtype=torch.float64).pin_memory()
tf = torch.empty(n, dtype=torch.float64).pin_memory()
na = torch.zeros(n, dtype=torch.int32).pin_memory(); nr = torch.zeros(n, dtype=torch.int32).pin_memory(); rc = torch.zeros(n, dtype=torch.int32).pin_memory()
alg = S.GPUSimpleATsit5(); keep = []
o = S.api.make_options(alg, np.dtype(np.float64), n, (0.0, 20.0), float(np.float32(0.1)), 1e-6, 1e-6, None, 0, 0, 0, 0, keep)
for devs in ([0], list(range(ndev))) if ndev > 1 else ([0],):
    d = (ctypes.c_int * len(devs))(*devs)
    for env in ("dynamic", "static"):
        if env == "static":
            os.environ["SDE_TUNE_STATIC_SHARDS"] = "1"
        else:
            os.environ.pop("SDE_TUNE_STATIC_SHARDS", None)
        best = 1e9
        for rep in range(4):
            t0 = time.perf_counter()
            r = _lib.lib().sde_solve(S.systems.vanderpol._handle, ctypes.byref(o), u0.data_ptr(), p.data_ptr(), out.data_ptr(), tf.data_ptr(),
                                     na.data_ptr(), nr.data_ptr(), rc.data_ptr(), d, len(devs))
            _lib.check(r)
            if rep: best = min(best, time.perf_counter() - t0)
        print("sde_solve VdP 2^%d %s, devices=%s %s: %.2f ms  %.4g accepted steps/s" % (int(np.log2(n)), "sorted" if SORTED else "shuffled", devs, env, best * 1e3, int(na.sum()) / best), flush=True)
        if len(devs) == 1: break
du0, dp = u0.cuda(), p.cuda()
S.solve_device(S.systems.vanderpol, alg, du0, dp, (0.0, 20.0), dt=float(np.float32(0.1)), abstol=1e-6, reltol=1e-6)
torch.cuda.synchronize(); t0 = time.perf_counter()
r = S.solve_device(S.systems.vanderpol, alg, du0, dp, (0.0, 20.0), dt=float(np.float32(0.1)), abstol=1e-6, reltol=1e-6)
print("device-resident: %.2f ms" % ((time.perf_counter() - t0) * 1e3))
