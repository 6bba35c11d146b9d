"""Generate tests/golden/golden_v1.json from the CPU oracle (oracle/oracle.cpp).

IMPORTANT: these vectors pin the ORACLE (regression protection + a fixture the GPU tests can be
checked against without recomputing); they are NOT outputs of the Julia reference, which cannot
run in the build container (no Julia).  Values are stored as hex strings of the IEEE bits.

    python tests/golden/make_golden.py
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import common as C  # noqa: E402
import oracle_lib as O  # noqa: E402
from simplediffeq_b200 import jl_range  # noqa: E402

CASES = []
for dtype in ("float64", "float32"):
    for alg in ("Tsit5", "RK4", "Vern7", "Vern9", "Euler"):
        CASES.append(dict(system="lorenz", alg=alg, dtype=dtype, n=6, tspan=[0.0, 1.0], dt=0.01, mode="endpoint"))
        CASES.append(dict(system="nonautonomous", alg=alg, dtype=dtype, n=4, tspan=[0.0, 1.0], dt=0.05, mode="endpoint"))
    for alg in ("Tsit5", "Vern7", "Vern9"):
        CASES.append(dict(system="lorenz", alg=alg, dtype=dtype, n=3, tspan=[0.0, 1.0], dt=0.05, mode="saveat",
                          saveat=[0.0, 0.01, 0.05, 0.07, 0.5, 1.0]))
    for alg, tol in (("ATsit5", 1e-6), ("AVern7", 1e-7), ("AVern9", 1e-8)):
        CASES.append(dict(system="lorenz", alg=alg, dtype=dtype, n=4, tspan=[0.0, 2.0], dt=float(np.float32(0.1)),
                          tol=tol if dtype == "float64" else 1e-4, mode="endpoint"))
CASES.append(dict(system="lorenz", alg="Tsit5", dtype="float64", n=2, tspan=[0.0, 10.0], dt=0.001, mode="endpoint", kat=True))


def inputs(case):
    dtype = np.dtype(case["dtype"]).type
    if case.get("kat"):
        # SURVEY.md 8c provisional KAT inputs: u0 = (1,0,0), p = (10, rho, 8/3), rho in {28, 21}
        u0 = np.array([[1, 0, 0], [1, 0, 0]], dtype=dtype)
        p = np.array([[10, 28, 8.0 / 3.0], [10, 21, 8.0 / 3.0]], dtype=dtype)
        return u0, p
    return C.random_problem(case["system"], case["n"], dtype, seed=1234)


def run(case):
    dtype = np.dtype(case["dtype"]).type
    u0, p = inputs(case)
    t0, tf = case["tspan"]
    kw = {}
    if case["alg"] in ("ATsit5", "AVern7", "AVern9"):
        kw.update(abstol=case["tol"], reltol=case["tol"])
    else:
        kw.update(tgrid=jl_range(dtype(t0), dtype(case["dt"]), dtype(tf), dtype))
    if case["mode"] == "saveat":
        kw.update(saveat=np.array(case["saveat"], dtype=dtype))
    r = O.solve(case["system"], case["alg"], u0, p, t0, tf, case["dt"], dtype=dtype, **kw)
    return r


def hexbits(a):
    a = np.ascontiguousarray(a)
    it = np.uint64 if a.dtype == np.float64 else np.uint32
    return [format(int(x), "x") for x in a.view(it).ravel()]


def main():
    out = []
    for case in CASES:
        r = run(case)
        e = dict(case)
        e["u_shape"] = list(r.u.shape)
        e["u_hex"] = hexbits(r.u)
        e["naccept"] = [int(x) for x in r.naccept]
        e["nreject"] = [int(x) for x in r.nreject]
        out.append(e)
    path = os.path.join(HERE, "golden_v1.json")
    with open(path, "w") as f:
        json.dump(out, f, indent=0)
    print("wrote", path, len(out), "cases", os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
