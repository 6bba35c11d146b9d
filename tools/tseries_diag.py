"""Development aid (cited by tests/test_gpu_parity.py::test_adaptive_everystep_variable_length): how far the time series of an adaptive
every-step solve moves when `pow` differs by one ulp -- the oracle against its own 1-ulp-pow twin and against the GPU."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
import simplediffeq_b200 as S, oracle_lib as O, common as C
n = 64
u0, p = C.lorenz_sweep(n)
tol, tspan, dt0 = 1e-8, (0.0, 3.0), float(np.float32(0.1))
o = O.solve("lorenz", "ATsit5", u0, p, 0.0, 3.0, dt0, abstol=tol, reltol=tol, save_mode=2, max_out=400, want_t=True)
o1 = O.solve("lorenz", "ATsit5", u0, p, 0.0, 3.0, dt0, abstol=tol, reltol=tol, save_mode=2, max_out=400, want_t=True, compat=16)
res = {}
for compat in (0, 2):
    g = S.solve_arrays(S.systems.lorenz, S.GPUSimpleATsit5(), np.ascontiguousarray(u0.T), np.ascontiguousarray(p.T), tspan, dt=dt0,
                       abstol=tol, reltol=tol, save_mode=2, compat=compat, out_capacity=400)
    res[compat] = g["t_series"]
i = 40
k = int(o.n[i])
print("naccept", k - 1)
for name, a in (("gpu fast", res[0][i]), ("gpu strict", res[2][i]), ("oracle pow+1ulp", o1.t[i])):
    d = np.abs(a[:k] - o.t[i, :k]) / np.maximum(np.abs(o.t[i, :k]), 1e-300)
    print("%-16s vs oracle: max rel diff in t = %.3g; first 6: %s" % (name, d[1:].max(), " ".join("%.2g" % x for x in d[1:7])))
d = np.abs(res[0][i][:k] - res[2][i][:k]) / np.maximum(np.abs(res[2][i][:k]), 1e-300)
print("gpu fast vs gpu strict: %.3g" % d[1:].max())
dts = np.diff(o.t[i, :k]); print("dts first:", dts[:5])
