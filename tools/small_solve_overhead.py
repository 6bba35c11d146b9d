"""BASELINE config 1 at its own size (10 k Lorenz trajectories, GPUSimpleATsit5 tol 1e-8) through the host-buffer C ABI call:
where does the time between the 0.59 ms kernel and the wall clock of sde_solve go?  Pageable and pinned host buffers, the
SDE_TRACE phases of one call, and the device-resident launch for scale.
    python tools/small_solve_overhead.py [n_traj]"""
import ctypes, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import simplediffeq_b200 as S
from simplediffeq_b200 import _lib

n = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000
alg = S.GPUSimpleATsit5(); keep = []
dt0 = float(np.float32(0.1))
o = S.api.make_options(alg, np.dtype(np.float64), n, (0.0, 10.0), dt0, 1e-8, 1e-8, None, 0, 0, 0, 0, keep)
d = (ctypes.c_int * 1)(0)
L = _lib.lib()


def buffers(pinned):
    mk = (lambda *s, dtype=torch.float64: torch.zeros(*s, dtype=dtype).pin_memory()) if pinned else (lambda *s, dtype=torch.float64: torch.zeros(*s, dtype=dtype))
    u0 = mk(3, n); u0[0] = 1
    p = mk(3, n); p[0] = 10; p[1] = 21.0 * torch.arange(n, dtype=torch.float64) / (n - 1); p[2] = 8.0 / 3.0
    return u0, p, mk(3, n), mk(n), mk(n, dtype=torch.int32), mk(n, dtype=torch.int32), mk(n, dtype=torch.int32)


def call(b):
    u0, p, out, tf, na, nr, rc = b
    _lib.check(L.sde_solve(S.systems.lorenz._handle, ctypes.byref(o), u0.data_ptr(), p.data_ptr(), out.data_ptr(), tf.data_ptr(),
                           na.data_ptr(), nr.data_ptr(), rc.data_ptr(), d, 1))


for general in (True, False):      # SDE_TUNE_NO_SMALL: the general (pipelined pieces) path instead of the cached small-solve context
    if general: os.environ["SDE_TUNE_NO_SMALL"] = "1"
    else: os.environ.pop("SDE_TUNE_NO_SMALL", None)
    for pinned in (False, True):
        b = buffers(pinned)
        for _ in range(5): call(b)
        ts = []
        for _ in range(50):
            t0 = time.perf_counter(); call(b); ts.append(time.perf_counter() - t0)
        ts = np.array(ts) * 1e3
        print("sde_solve, %d trajectories, %s path, %s host buffers: best %.3f ms, median %.3f ms (50 calls)"
              % (n, "general" if general else "small-solve", "pinned" if pinned else "pageable", ts.min(), np.median(ts)), flush=True)
        if general: ref = b[2].clone()
        else: assert torch.equal(b[2], ref), "small-solve and general path differ"
du0, dp = b[0].cuda(), b[1].cuda()
for _ in range(5): r = S.solve_device(S.systems.lorenz, alg, du0, dp, (0.0, 10.0), dt=dt0, abstol=1e-8, reltol=1e-8, sync=False)
ev = [torch.cuda.Event(enable_timing=True) for _ in range(21)]
ev[0].record()
for k in range(20):
    S.solve_device(S.systems.lorenz, alg, du0, dp, (0.0, 10.0), dt=dt0, abstol=1e-8, reltol=1e-8, sync=False); ev[k + 1].record()
torch.cuda.synchronize()
print("device-resident launch (CUDA events): %.3f ms" % (ev[0].elapsed_time(ev[20]) / 20))
assert torch.equal(r["u"].cpu(), ref), "host-buffer and device-resident results differ"
t0 = time.perf_counter()
g = S.solve_arrays(S.systems.lorenz, alg, b[0].numpy(), b[1].numpy(), (0.0, 10.0), dt=dt0, abstol=1e-8, reltol=1e-8)
print("solve_arrays (python mirror, allocates its outputs): %.3f ms" % ((time.perf_counter() - t0) * 1e3))
os.environ["SDE_TRACE"] = "1"
sys.stderr.flush()
call(buffers(False))
os.environ["SDE_TUNE_NO_SMALL"] = "1"
call(buffers(False))
