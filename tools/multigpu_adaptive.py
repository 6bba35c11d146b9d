"""Development aid: in-library multi-device sharding of an adaptive sweep (host buffers, wall clock)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import simplediffeq_b200 as S
from simplediffeq_b200 import _lib
n = 1 << 22
u0 = np.zeros((2, n)); u0[0] = 2
mu = (0.1 + 49.9 * np.arange(n) / (n - 1)).reshape(1, n)
ndev = _lib.device_count()
for devs in ([0], list(range(ndev))):
    for rep in range(3):
        t0 = time.perf_counter()
        r = S.solve_arrays(S.systems.vanderpol, S.GPUSimpleATsit5(), u0, mu, (0.0, 20.0), dt=float(np.float32(0.1)), abstol=1e-6, reltol=1e-6, devices=devs)
        dt = time.perf_counter() - t0
    print("%s devices=%s: %.1f ms  accepted steps/s = %.4g" % (os.environ.get("SDE_TUNE_STATIC_SHARDS", "dynamic"), devs, dt * 1e3, r["naccept"].sum() / dt), flush=True)
