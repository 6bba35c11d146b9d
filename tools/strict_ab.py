"""Literal controller (SDE_COMPAT_STRICT_CONTROLLER, own pow = sde_pow_glibc) against the default log2-domain one on
the adaptive BASELINE configs: device time, identical-step-count share against the CPU oracle on a 4096-trajectory
sample.  Development aid for the first GPU call after the strict pow was introduced (never run so far):
    python tools/strict_ab.py > gpurun_out/strict_ab.txt"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, os.path.join(ROOT, "tools"))
import numpy as np  # noqa: E402
import torch  # noqa: E402
import simplediffeq_b200 as S  # noqa: E402
from simplediffeq_b200 import _lib  # noqa: E402
import common as C  # noqa: E402
import oracle_lib  # noqa: E402
from quick_bench import probe  # noqa: E402

CASES = [("lorenz", S.GPUSimpleATsit5, "ATsit5", 1 << 20, (0.0, 10.0), 1e-8),
         ("vanderpol", S.GPUSimpleATsit5, "ATsit5", 1 << 20, (0.0, 20.0), 1e-6),
         ("lorenz", S.GPUSimpleAVern7, "AVern7", 1 << 20, (0.0, 10.0), 1e-10),
         ("lorenz", S.GPUSimpleAVern9, "AVern9", 1000000, (0.0, 10.0), 1e-12)]


def main():
    dt0 = float(np.float32(0.1))
    oracle_lib.build()
    for system, alg, oname, n, tspan, tol in CASES:
        for compat, label in ((0, "log2"), (_lib.COMPAT_STRICT_CONTROLLER, "literal")):
            print("[%s]" % label, end=" ")
            probe(system, alg(), n, tspan, dt0, abstol=tol, reltol=tol, compat=compat)
            m = 4096
            u0, p = (C.lorenz_sweep(m) if system == "lorenz" else C.vdp_sweep(m))
            g = S.solve_arrays(getattr(S.systems, system), alg(), np.ascontiguousarray(u0.T), np.ascontiguousarray(p.T),
                               tspan, dt=dt0, abstol=tol, reltol=tol, save_mode=0, compat=compat)
            o = oracle_lib.solve(system, oname, u0, p, tspan[0], tspan[1], dt0, abstol=tol, reltol=tol, n_threads=8)
            same = np.mean((g["naccept"] == o.naccept) & (g["nreject"] == o.nreject))
            bits = np.mean(np.all(g["u"].T.view(np.uint64) == o.u[:, 0, :].view(np.uint64), axis=1))
            print("        vs oracle on %d: identical step counts %.3f %%, bit-identical final states %.3f %%"
                  % (m, 100 * same, 100 * bits), flush=True)


if __name__ == "__main__":
    if not torch.cuda.is_available():
        sys.exit("needs a GPU")
    main()
