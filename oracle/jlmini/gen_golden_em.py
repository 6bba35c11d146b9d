#!/usr/bin/env python3
"""Generate tests/golden/golden_jlmini_em_v1.json by EXECUTING THE REFERENCE'S OWN SimpleEM SOURCE.

    python oracle/jlmini/gen_golden_em.py         # needs /root/reference (build container only)

The out-of-place `solve(prob::SDEProblem{uType,tType,false}, alg::SimpleEM; dt)` method of
`src/euler_maruyama.jl:46-94` is parsed and executed by the jlmini interpreter (its `@muladd` rewriting
included).  The one thing that is NOT the reference's: `randn`.  The reference draws from Julia's
task-local RNG, which nothing outside that Julia process can reproduce, so the interpreter's `randn`
hands out normals from a list stored in the fixture (component order = the order Julia fills
`randn(SVector{N,T})`: 1..N).  Everything else -- the step count `Int((tspan[2]-tspan[1])/dt) + 1`, the
time grid, `sqrt(dt)`, where the macro puts the FMAs in the scalar and the diagonal-vector update -- is
the reference's text.  The non-diagonal branch (`sqdt * g(...) * randn(m)`, a Matrix-Vector `muladd`
whose rounding order is BLAS-dependent) is outside the subset: assumption A11 stays an assumption.

Drift / diffusion functions: the reference's tests define them as one-liners at top level
(test/simpleem_tests.jl:4-5 `f(u,p,t) = 2u`, `g(u,p,t) = 1`; docstring src/euler_maruyama.jl:27-28
`f(u,p,t) = 0.1u`, `g(u,p,t) = 0.2u`); they are restated below as Python callables over jlmini values
with the parameters in `p` (operation order of oracle_em.cpp:58-80).

TEST INFRASTRUCTURE ONLY.  tests/test_oracle_em_jlmini.py checks oracle_em.cpp and the host layer's
step count / time grid against this file; tests/test_gpu_em.py checks the CUDA kernels against it.
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, HERE)

import jlmini as M  # noqa: E402
from jlmini import F32, F64, JlError, JlVector, SVec, jl_binop  # noqa: E402

REF = os.environ.get("SDE_REFERENCE", "/root/reference")
OUT = os.path.join(ROOT, "tests", "golden", "golden_jlmini_em_v1.json")


def _mul(a, b):
    return jl_binop("*", a, b)


def _sub(a, b):
    return jl_binop("-", a, b)


def _bc(fn, u):
    return SVec(fn(x) for x in u.v) if isinstance(u, SVec) else fn(u)


# name -> (f, g); p is a JlVector of T
SYSTEMS = {
    "gbm": (lambda u, p, t: _bc(lambda x: _mul(p.items[0], x), u), lambda u, p, t: _bc(lambda x: _mul(p.items[1], x), u)),
    "linadd1": (lambda u, p, t: _mul(p.items[0], u), lambda u, p, t: p.items[1]),
    # g returns the scalar p2; `sqdt * g .* randn(SVector)` broadcasts it (test/simpleem_tests.jl:16-18)
    "linadd2": (lambda u, p, t: _bc(lambda x: _mul(p.items[0], x), u), lambda u, p, t: p.items[1]),
    "ou": (lambda u, p, t: _mul(p.items[0], _sub(p.items[1], u)), lambda u, p, t: p.items[2]),
}

CASES = []


def case(name, system, u0, p, tspan, dt, dtype="float64", seed=1):
    CASES.append(dict(name=name, system=system, u0=u0, p=p, tspan=list(tspan), dt=dt, dtype=dtype, seed=seed))


for dt_ in ("float64", "float32"):
    sfx = "_" + dt_[-2:]
    case("reftest_scalar" + sfx, "linadd1", 0.5, [2.0, 1.0], (0.0, 1.0), 0.25, dt_)          # test/simpleem_tests.jl:4-10
    case("reftest_svector" + sfx, "linadd2", [0.1, 0.2], [2.0, 1.0], (0.0, 1.0), 0.25, dt_)   # :16-18
    case("docstring_gbm" + sfx, "gbm", 0.5, [0.1, 0.2], (0.0, 1.0), 0.125, dt_)               # src/euler_maruyama.jl:27-31
    case("gbm_64steps" + sfx, "gbm", 1.0, [0.1, 0.2], (0.0, 1.0), 1.0 / 64, dt_, seed=2)
    case("gbm_t0_offset" + sfx, "gbm", 1.5, [-0.3, 0.45], (0.5, 2.0), 0.0625, dt_, seed=3)      # time grid muladd(i, dt, t0)
    case("ou_100steps" + sfx, "ou", 0.25, [1.5, 0.7, 0.3], (0.0, 1.5625), 0.015625, dt_, seed=4)
    case("linadd2_40steps" + sfx, "linadd2", [-0.4, 1.25], [0.75, 0.5], (1.0, 6.0), 0.125, dt_, seed=5)
    case("scalar_zero_steps" + sfx, "linadd1", 0.5, [2.0, 1.0], (1.0, 1.0), 0.25, dt_)        # n = 1: only u0
    # Int((tspan[2] - tspan[1]) / dt) is an InexactError when dt does not divide the span
    case("inexact_error" + sfx, "linadd1", 0.5, [2.0, 1.0], (0.0, 1.0), 0.3, dt_)


def hexbits(x, T):
    a = np.asarray(x, dtype=T)
    it = np.uint64 if T is np.float64 else np.uint32
    return [format(int(v), "x") for v in a.view(it).ravel()]


_interp = None


def interp():
    global _interp
    if _interp is None:
        _interp = M.Interp()
        _interp.load(os.path.join(REF, "src/euler_maruyama.jl"), wanted={"solve"}, first_only=True)
        fn = _interp.globals.lookup("solve")
        assert len(fn.methods) == 1
    return _interp


def run(c):
    T = F64 if c["dtype"] == "float64" else F32
    npT = np.float64 if T is F64 else np.float32
    scalar = not isinstance(c["u0"], list)
    n_comp = 1 if scalar else len(c["u0"])
    u0 = T(c["u0"]) if scalar else SVec(T(x) for x in c["u0"])
    p = JlVector([T(x) for x in c["p"]])
    f, g = SYSTEMS[c["system"]]
    # a generous supply of normals, rounded to T once (what the kernel / oracle are fed as well)
    rng = np.random.default_rng(20261017 + c["seed"])
    supply = rng.standard_normal(4096).astype(npT)
    used = []

    def randn(ty):
        k = len(used)
        if isinstance(ty, M.TypeApp):            # randn(SVector{N,T})
            n = int(ty.params[0])
            used.extend(supply[k:k + n])
            return SVec(T(x) for x in supply[k:k + n])
        used.append(supply[k])
        return T(supply[k])

    it = interp()
    it.randn_hook = randn
    prob = M.SDEProblem(f, g, u0, (T(c["tspan"][0]), T(c["tspan"][1])), p)
    out = dict(c)
    try:
        sol = it.call(it.globals.lookup("solve"), [prob, M.Struct("SimpleEM", [], [])], {"dt": T(c["dt"])})
    except JlError as e:
        out["error"] = str(e).split(":")[0]
        return out
    ts, us = list(sol.t.items), list(sol.u.items)
    assert all(isinstance(x, T) for x in ts)
    out["n_out"] = len(us)
    out["t"] = hexbits(ts, npT)
    out["u"] = hexbits([[x for x in u.v] if isinstance(u, SVec) else [u] for u in us], npT)
    out["noise"] = hexbits(used, npT)            # [n_steps][n_comp]
    assert len(used) == (len(us) - 1) * n_comp
    return out


def main():
    if not os.path.isfile(os.path.join(REF, "src/euler_maruyama.jl")):
        print("reference tree not present at %s; nothing generated" % REF)
        return 1
    res = [run(c) for c in CASES]
    doc = {
        "generator": "oracle/jlmini/gen_golden_em.py (jlmini interpreter over src/euler_maruyama.jl; randn supplied)",
        "reference": "SciML/SimpleDiffEq.jl v1.16.3 at /root/reference",
        "cases": res,
    }
    with open(OUT, "w") as fh:
        json.dump(doc, fh, indent=0, separators=(",", ":"))
    print("wrote %s: %d cases, %.1f KB" % (OUT, len(res), os.path.getsize(OUT) / 1e3))
    return 0


if __name__ == "__main__":
    sys.exit(main())
