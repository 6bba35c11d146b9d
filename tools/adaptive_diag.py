"""Development aid: adaptive-parity statistics CUDA vs oracle (fraction of identical step counts etc.)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
import simplediffeq_b200 as S
import oracle_lib as O
import common as C

def diag(system, algname, dtype, n, tspan, tol, compat=0, seed=7, sweep=False):
    if sweep:
        u0, p = (C.lorenz_sweep(n, dtype) if system == "lorenz" else C.vdp_sweep(n, dtype))
        system_label = system + "-sweep"
    else:
        u0, p = C.random_problem(system, n, dtype, seed)
    dt0 = float(np.float32(0.1))
    g = S.solve_arrays(getattr(S.systems, system), getattr(S, algname)(), np.ascontiguousarray(u0.T), np.ascontiguousarray(p.T),
                       tspan, dt=dt0, abstol=tol, reltol=tol, compat=compat)
    o = O.solve(system, C.ALG_NAMES[algname], u0, p, tspan[0], tspan[1], dt0, dtype=dtype, abstol=tol, reltol=tol, n_threads=16)
    same_acc = np.mean(g["naccept"] == o.naccept); same_rej = np.mean(g["nreject"] == o.nreject)
    ou = o.u[:, 0, :]; gu = g["u"].T
    err = np.abs(gu - ou) / (tol + tol * np.abs(ou))
    d = g["naccept"].astype(int) - o.naccept
    print("%-13s %-16s %-8s tol=%g compat=%d n=%d: same naccept %.4f%% same nreject %.4f%%  max|dacc|=%d  err(max,99.9%%)=%.3g,%.3g  mean acc=%.1f rej=%.1f ret!=0: %d/%d"
          % (system, algname, np.dtype(dtype).name, tol, compat, n, 100 * same_acc, 100 * same_rej, np.abs(d).max(), np.nanmax(err), np.nanquantile(err, 0.999),
             o.naccept.mean(), o.nreject.mean(), (g["retcode"] != 0).sum(), (o.retcode != 0).sum()), flush=True)

if __name__ == "__main__":
    compats = [int(x) for x in sys.argv[1:]] or [0]
    for compat in compats:
        for dtype, tol in ((np.float64, 1e-8), (np.float32, 1e-4)):
            for alg in ("GPUSimpleATsit5", "GPUSimpleAVern7", "GPUSimpleAVern9"):
                diag("lorenz", alg, dtype, 10000, (0.0, 10.0), tol, compat, sweep=True)
        diag("lorenz", "GPUSimpleAVern9", np.float64, 10000, (0.0, 10.0), 1e-12, compat, sweep=True)
        diag("lorenz", "GPUSimpleAVern7", np.float64, 10000, (0.0, 10.0), 1e-10, compat, sweep=True)
        for dtype, tol in ((np.float64, 1e-6), (np.float32, 1e-4)):
            for alg in ("GPUSimpleATsit5", "GPUSimpleAVern7", "GPUSimpleAVern9"):
                diag("vanderpol", alg, dtype, 20000, (0.0, 20.0), tol, compat, sweep=True)
    sys.exit(0)
    for compat in compats:
        for system, tspan in (("lorenz", (0.0, 2.0)), ("vanderpol", (0.0, 2.0)), ("nonautonomous", (0.0, 2.0))):
            for alg in ("GPUSimpleATsit5", "GPUSimpleAVern7", "GPUSimpleAVern9"):
                diag(system, alg, np.float64, 20000, tspan, 1e-8, compat)
                diag(system, alg, np.float32, 20000, tspan, 1e-4, compat)
        diag("lorenz", "GPUSimpleATsit5", np.float64, 20000, (0.0, 10.0), 1e-8, compat)
        diag("lorenz", "GPUSimpleAVern9", np.float64, 20000, (0.0, 10.0), 1e-12, compat)
        diag("lorenz", "GPUSimpleATsit5", np.float32, 20000, (0.0, 10.0), 1e-4, compat)
