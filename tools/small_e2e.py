"""Development aid: end-to-end latency of small ensembles (config 1 size) through the Python mirror."""
import sys, time, numpy as np
import os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import simplediffeq_b200 as S
n = 10000
u0 = np.zeros((3, n)); u0[0] = 1
p = np.empty((3, n)); p[0] = 10; p[1] = 21.0 * np.arange(n) / (n - 1); p[2] = 8 / 3
for alg, kw in ((S.GPUSimpleATsit5(), dict(dt=float(np.float32(0.1)), abstol=1e-8, reltol=1e-8)), (S.GPUSimpleTsit5(), dict(dt=0.01))):
    best = 1e9
    for r in range(6):
        t0 = time.perf_counter(); g = S.solve_arrays(S.systems.lorenz, alg, u0, p, (0.0, 10.0), devices=[0], **kw); dt = time.perf_counter() - t0
        if r: best = min(best, dt)
    print(type(alg).__name__, "10k e2e (python + sde_solve, pageable): %.3f ms" % (best * 1e3))
