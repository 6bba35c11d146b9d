"""Development aid: wall clock of a small SimpleEM ensemble through the host-buffer call (sde_em_solve), with and without the
cudaMemGetInfo query the launcher used to make on every call (SDE_TUNE_EM_MEMINFO=1 brings it back)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import simplediffeq_b200 as S
n, steps = 10_000, 100
u0 = np.ones((1, n)); p = np.tile([[0.1], [0.2]], (1, n))
for query in (True, False):
    if query: os.environ["SDE_TUNE_EM_MEMINFO"] = "1"
    else: os.environ.pop("SDE_TUNE_EM_MEMINFO", None)
    for save_mode, name in ((0, "endpoint"), (2, "every state")):
        for _ in range(5): r = S.solve_em_arrays(S.sde_systems.gbm, u0, p, 0.0, 0.01, steps, seed=3, save_mode=save_mode)
        ts = []
        for _ in range(30):
            t0 = time.perf_counter(); r = S.solve_em_arrays(S.sde_systems.gbm, u0, p, 0.0, 0.01, steps, seed=3, save_mode=save_mode); ts.append(time.perf_counter() - t0)
        print("solve_em_arrays GBM %d paths x %d steps, %-11s, %s: best %.3f ms, median %.3f ms" % (n, steps, name, "with cudaMemGetInfo" if query else "without", min(ts) * 1e3, np.median(ts) * 1e3), flush=True)
