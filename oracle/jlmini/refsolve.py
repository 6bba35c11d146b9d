"""Run the reference's GPU-style `solve` methods from their own source text (see jlmini.py).

TEST INFRASTRUCTURE ONLY.  Needs the reference tree (default /root/reference, the build container);
nothing on the GPU box imports this -- the committed fixtures under tests/golden/ travel instead.
"""
import os

import numpy as np

from jlmini import (F32, F64, Interp, JlVector, Problem, Struct, SVec, MULADD_NOTES)  # noqa: F401

REF = os.environ.get("SDE_REFERENCE", "/root/reference")

SOLVER_FILES = [
    "src/tsit5/gpuatsit5.jl", "src/rk4/gpurk4.jl", "src/euler/gpueuler.jl",
    "src/verner/gpuvern7.jl", "src/verner/gpuvern9.jl",
]
ALGS = ["GPUSimpleTsit5", "GPUSimpleATsit5", "GPUSimpleRK4", "GPUSimpleEuler", "GPUSimpleVern7",
        "GPUSimpleAVern7", "GPUSimpleVern9", "GPUSimpleAVern9"]


def available():
    return os.path.isfile(os.path.join(REF, "src/tsit5/gpuatsit5.jl"))


_interp = None


def interp():
    """One interpreter with every definition of the hot path loaded from the reference tree."""
    global _interp
    if _interp is not None:
        return _interp
    it = Interp()
    it.load(os.path.join(REF, "src/SimpleDiffEq.jl"), wanted={"build_adaptive_controller_cache"})
    it.load(os.path.join(REF, "src/tsit5/atsit5_cache.jl"))
    it.load(os.path.join(REF, "src/tsit5/tsit5.jl"), wanted={"bθs"})
    it.load(os.path.join(REF, "src/verner/verner_tableaus.jl"))
    for f in SOLVER_FILES:
        it.load(os.path.join(REF, f), wanted={"solve"})
    # right-hand sides the reference's own tests define
    it.load(os.path.join(REF, "test/gpusimpleatsit5_tests.jl"), wanted={"loop"})
    it.load(os.path.join(REF, "test/gpu_ode_regression.jl"), wanted={"test"})
    _interp = it
    return it


def _count_f(it, f):
    def wrapped(u, p, t):
        it.stats["f_calls"] += 1
        return it.call(f, [u, p, t]) if not callable(f) else f(u, p, t)
    return wrapped


def svec(vals, T):
    return SVec(T(v) for v in vals)


def solve(alg, f, u0, tspan, p, **kw):
    """solve(ODEProblem{false}(f, u0, tspan, p), alg(); kw...) through the reference's source.

    f: name of a Julia function loaded from the reference tests ("loop" = Lorenz, "test" = -u) or a
    Python callable (u::SVec, p, t) -> SVec working on jlmini values.  Returns (ts, us, f_calls)
    with ts a list of scalars and us a list of SVec / scalars / None (`undef`)."""
    it = interp()
    fn = it.globals.lookup(f) if isinstance(f, str) else f
    it.stats["f_calls"] = 0
    prob = Problem(_count_f(it, fn), u0, tspan, p)
    algv = Struct(alg, [], [])
    sol = it.call(it.globals.lookup("solve"), [prob, algv], kw)
    ts = sol.t
    if isinstance(ts, JlVector):
        ts = list(ts.items)
    elif hasattr(ts, "vals"):
        ts = [ts.T(x) for x in ts.vals]
    us = list(sol.u.items) if isinstance(sol.u, JlVector) else list(sol.u)
    return ts, us, it.stats["f_calls"]


def to_array(us, n, dtype):
    out = np.full((len(us), n), np.nan, dtype=dtype)
    for i, u in enumerate(us):
        if u is None:
            continue            # `undef` slot (quirk Q5)
        out[i] = [x for x in u.v] if isinstance(u, SVec) else [u]
    return out
