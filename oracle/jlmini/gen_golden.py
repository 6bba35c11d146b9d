#!/usr/bin/env python3
"""Generate tests/golden/golden_jlmini_v1.json by EXECUTING THE REFERENCE'S OWN SOURCE (jlmini.py).

    python oracle/jlmini/gen_golden.py            # needs /root/reference (build container only)

Each case records its inputs and the outputs (`sol.t`, `sol.u`, number of `f` calls) of
`solve(ODEProblem{false}(f, u0, tspan, p), alg(); kw...)` as the reference's `solve` method computes
them; floating-point values are stored as hex strings of their IEEE bits.  `f` is either a function
parsed from the reference's tests ("loop" = Lorenz, test/gpusimpleatsit5_tests.jl:3-13; "test" =
-u, test/gpu_ode_regression.jl:2-4) or one of the registry systems restated below with the
operation order of DESIGN.md / oracle.cpp (those have no definition in the reference).

TEST INFRASTRUCTURE ONLY.  tests/test_oracle_jlmini.py checks the C++ oracle against this file,
tests/test_gpu_parity.py checks the CUDA path against it.
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, HERE)

import refsolve as R  # noqa: E402
from jlmini import F32, F64, JlError, JlVector, SVec, jl_binop, MULADD_NOTES  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden", "golden_jlmini_v1.json")


# ---- registry systems that the reference does not define (operation order = oracle.cpp:118-191)
def _mul(a, b):
    return jl_binop("*", a, b)


def _sub(a, b):
    return jl_binop("-", a, b)


def _add(a, b):
    return jl_binop("+", a, b)


def vanderpol(u, p, t):
    u1, u2 = u.v
    one = type(u1)(1)
    return SVec([u2, _sub(_mul(_mul(p.items[0], _sub(one, _mul(u1, u1))), u2), u1)])


def robertson(u, p, t):
    u1, u2, u3 = u.v
    k1, k2, k3 = p.items
    return SVec([_add(_mul(-k1, u1), _mul(_mul(k3, u2), u3)),
                 _sub(_sub(_mul(k1, u1), _mul(_mul(k2, u2), u2)), _mul(_mul(k3, u2), u3)),
                 _mul(_mul(k2, u2), u2)])


def nonautonomous(u, p, t):
    u1, u2 = u.v
    a, b = p.items
    T = type(u1)
    t = T(t)
    return SVec([_add(u2, t), _add(_mul(-a, u1), _mul(_mul(b, t), t))])


def scalargrowth(u, p, t):
    return _mul(p.items[0], u)


PY_RHS = dict(vanderpol=vanderpol, robertson=robertson, nonautonomous=nonautonomous, scalargrowth=scalargrowth)
JL_RHS = dict(lorenz="loop", lineardecay="test")

LORENZ = ([1.0, 0.0, 0.0], [10.0, 28.0, 8.0 / 3.0])
LORENZ_T = ([10.0, 10.0, 10.0], [10.0, 28.0, 8.0 / 3.0])     # the reference test's own problem
F01 = float(np.float32(0.1))                                  # the default dt = 0.1f0

CASES = []


def case(name, alg, system, u0, p, tspan, dtype="float64", **kw):
    CASES.append(dict(name=name, alg=alg, system=system, u0=u0, p=p, tspan=list(tspan), dtype=dtype, kw=kw))


for dt_ in ("float64", "float32"):
    sfx = "_" + dt_[-2:]
    for alg in ("GPUSimpleTsit5", "GPUSimpleVern7", "GPUSimpleVern9"):
        case(alg + "_lorenz_endpoint" + sfx, alg, "lorenz", *LORENZ, (0.0, 1.0), dt_, dt=0.01, save_everystep=False)
        case(alg + "_lorenz_everystep" + sfx, alg, "lorenz", *LORENZ, (0.0, 0.5), dt_, dt=0.01)
        case(alg + "_lorenz_saveat" + sfx, alg, "lorenz", *LORENZ, (0.0, 1.0), dt_, dt=0.05,
             saveat=[0.0, 0.01, 0.05, 0.07, 0.5, 1.0])
        # several save points per step, first save point after t0 (quirk Q8), dt not dividing the span and a
        # save point beyond the last step (quirk Q5: `undef`)
        case(alg + "_lorenz_saveat_q5q8" + sfx, alg, "lorenz", *LORENZ, (0.0, 1.0), dt_, dt=0.03,
             saveat=[0.005, 0.01, 0.015, 0.03, 0.6, 0.85, 0.95, 0.99, 1.0])
        # unstable step size: the state overflows to Inf / NaN and keeps being stepped (no error in the reference)
        case(alg + "_lorenz_saveat_blowup" + sfx, alg, "lorenz", *LORENZ, (0.0, 3.0), dt_, dt=0.3,
             saveat=[0.05, 0.1, 0.15, 0.3, 0.6, 0.85, 0.95, 1.0, 2.5, 3.0])
        case(alg + "_nonauto_saveat" + sfx, alg, "nonautonomous", [0.3, -0.2], [1.5, 0.7], (0.5, 1.5), dt_, dt=0.125,
             saveat=[0.5, 0.6, 0.75, 1.0, 1.4375, 1.5])
        case(alg + "_decay_reftest" + sfx, alg, "lineardecay", [1.0, 1.0, 1.0], [10.0, 28.0, 8.0 / 3.0], (0.0, 1.0), dt_,
             dt=0.01, saveat=[0.0, 0.4])
    for alg in ("GPUSimpleRK4", "GPUSimpleEuler"):
        case(alg + "_lorenz" + sfx, alg, "lorenz", *LORENZ, (0.0, 1.0), dt_, dt=0.01)
        case(alg + "_nonauto" + sfx, alg, "nonautonomous", [0.3, -0.2], [1.5, 0.7], (0.5, 1.5), dt_, dt=0.125)
        case(alg + "_lorenz_ragged" + sfx, alg, "lorenz", *LORENZ, (0.0, 1.0), dt_, dt=0.03)
        case(alg + "_lorenz_blowup" + sfx, alg, "lorenz", *LORENZ, (0.0, 3.0), dt_, dt=0.3)
    for alg, tol in (("GPUSimpleATsit5", 1e-8), ("GPUSimpleAVern7", 1e-9), ("GPUSimpleAVern9", 1e-10)):
        tl = tol if dt_ == "float64" else 1e-5
        case(alg + "_lorenz_endpoint" + sfx, alg, "lorenz", *LORENZ, (0.0, 2.0), dt_, dt=F01, abstol=tl, reltol=tl,
             save_everystep=False)
        case(alg + "_lorenz_everystep" + sfx, alg, "lorenz", *LORENZ, (0.0, 1.0), dt_, dt=F01, abstol=tl, reltol=tl)
        case(alg + "_lorenz_saveat" + sfx, alg, "lorenz", *LORENZ, (0.0, 2.0), dt_, dt=F01, abstol=tl, reltol=tl,
             saveat=[0.0, 0.3, 0.31, 0.32, 1.0, 2.0])
        case(alg + "_nonauto_saveat" + sfx, alg, "nonautonomous", [0.3, -0.2], [1.5, 0.7], (0.5, 2.5), dt_, dt=F01,
             abstol=tl, reltol=tl, saveat=[0.6, 0.75, 1.0, 2.4375, 2.5])
        case(alg + "_vdp_endpoint" + sfx, alg, "vanderpol", [2.0, 0.0], [3.0], (0.0, 3.0), dt_, dt=F01, abstol=tl,
             reltol=tl, save_everystep=False)
        case(alg + "_robertson_endpoint" + sfx, alg, "robertson", [1.0, 0.0, 0.0], [0.04, 30.0, 10.0], (0.0, 1.0), dt_,
             dt=F01, abstol=tl, reltol=tl, save_everystep=False)
        case(alg + "_scalar_everystep" + sfx, alg, "scalargrowth", 0.5, [1.01], (0.0, 1.0), dt_, dt=F01, abstol=tl,
             reltol=tl)
        case(alg + "_decay_reftest" + sfx, alg, "lineardecay", [1.0, 1.0, 1.0], [10.0, 28.0, 8.0 / 3.0], (0.0, 1.0), dt_,
             dt=0.01, abstol=1e-7, reltol=1e-7, save_everystep=False)

# the reference's own test problem (test/gpusimpleatsit5_tests.jl:24-33): Lorenz, u0 = 10 ones(3), dt = 1e-2,
# abstol 1e-6, reltol 1e-3 -- first steps of tspan (0, 100) (the test compares sol.u[5] / sol.t[5])
case("reftest_atsit5_lorenz_t5", "GPUSimpleATsit5", "lorenz", *LORENZ_T, (0.0, 5.0), dt=1e-2, abstol=1e-6, reltol=1e-3)
case("reftest_tsit5_lorenz_dt0.1", "GPUSimpleTsit5", "lorenz", *LORENZ_T, (0.0, 5.0), dt=0.1, saveat=[2.5, 5.0])
# BASELINE configs[0]-style trajectories: rho on the 21-sweep, tol 1e-8, tspan (0, 10), default dt
case("config0_atsit5_rho21", "GPUSimpleATsit5", "lorenz", [1.0, 0.0, 0.0], [10.0, 21.0, 8.0 / 3.0], (0.0, 10.0),
     dt=F01, abstol=1e-8, reltol=1e-8, save_everystep=False)
case("config0_atsit5_rho10.5", "GPUSimpleATsit5", "lorenz", [1.0, 0.0, 0.0], [10.0, 10.5, 8.0 / 3.0], (0.0, 10.0),
     dt=F01, abstol=1e-8, reltol=1e-8, save_everystep=False)
case("config3_atsit5_vdp_mu25", "GPUSimpleATsit5", "vanderpol", [2.0, 0.0], [25.0], (0.0, 20.0),
     dt=F01, abstol=1e-6, reltol=1e-6, save_everystep=False)
case("config4_avern9_rho21_tol1e-12", "GPUSimpleAVern9", "lorenz", [1.0, 0.0, 0.0], [10.0, 21.0, 8.0 / 3.0], (0.0, 10.0),
     dt=F01, abstol=1e-12, reltol=1e-12, save_everystep=False)
# SURVEY.md 8c known answer: fixed Tsit5, dt = 1e-3, 10 000 steps
case("kat_tsit5_rho28_10000steps", "GPUSimpleTsit5", "lorenz", [1.0, 0.0, 0.0], [10.0, 28.0, 8.0 / 3.0], (0.0, 10.0),
     dt=1e-3, save_everystep=False)
# failure path: error("dt<dtmin") when the controller drives dt below 1e-14
case("atsit5_dtmin_error", "GPUSimpleATsit5", "scalargrowth", 1.0, [-1e17], (0.0, 1.0), dt=1e-3, abstol=1e-10,
     reltol=1e-10, save_everystep=False)
# default keyword arguments (dt = 0.1f0, abstol = 1f-6, reltol = 1f-3, save_everystep = true) on a Float64 problem
case("atsit5_defaults", "GPUSimpleATsit5", "lorenz", *LORENZ, (0.0, 1.0))
case("tsit5_defaults", "GPUSimpleTsit5", "lorenz", *LORENZ, (0.0, 1.0))


def hexbits(x, T):
    a = np.asarray(x, dtype=T)
    it = np.uint64 if T is np.float64 else np.uint32
    return [format(int(v), "x") for v in a.view(it).ravel()]


def run(c):
    T = F64 if c["dtype"] == "float64" else F32
    npT = np.float64 if T is F64 else np.float32
    scalar = not isinstance(c["u0"], list)
    u0 = T(c["u0"]) if scalar else R.svec(c["u0"], T)
    p = JlVector([T(x) for x in c["p"]])
    kw = {}
    for k, v in c["kw"].items():
        if k == "saveat":
            kw[k] = JlVector([T(x) for x in v])
        elif k in ("dt", "abstol", "reltol"):
            kw[k] = T(v)
        else:
            kw[k] = v
    f = JL_RHS.get(c["system"]) or PY_RHS[c["system"]]
    out = dict(c)
    try:
        ts, us, nf = R.solve(c["alg"], f, u0, (T(c["tspan"][0]), T(c["tspan"][1])), p, **kw)
    except JlError as e:
        out["error"] = str(e)
        return out
    n = 1 if scalar else len(c["u0"])
    tT = np.float32 if (len(ts) and isinstance(ts[-1], F32)) else np.float64
    out["t_dtype"] = np.dtype(tT).name
    out["t"] = hexbits([tT(x) for x in ts], tT)
    out["u"] = hexbits(R.to_array(us, n, npT), npT)
    out["n_out"] = len(us)
    out["f_calls"] = nf
    return out


def main():
    if not R.available():
        print("reference tree not present at %s; nothing generated" % R.REF)
        return 1
    res = [run(c) for c in CASES]
    notes = sorted(set(MULADD_NOTES))
    doc = {
        "generator": "oracle/jlmini/gen_golden.py (jlmini interpreter over the reference's own source files)",
        "reference": "SciML/SimpleDiffEq.jl v1.16.3 at /root/reference",
        "muladd_notes": notes,
        "cases": res,
    }
    with open(OUT, "w") as fh:
        json.dump(doc, fh, indent=0, separators=(",", ":"))
    print("wrote %s: %d cases, %.1f KB, muladd notes: %s" % (OUT, len(res), os.path.getsize(OUT) / 1e3, notes))
    return 0


if __name__ == "__main__":
    sys.exit(main())
