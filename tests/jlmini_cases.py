"""Loader for tests/golden/golden_jlmini_v1.json (outputs of the reference's own source text run by
oracle/jlmini) and the adapters that run the same case on the CPU oracle.  Test infrastructure."""
import json
import os

import numpy as np

import common as C
import oracle_lib as O

HERE = os.path.dirname(os.path.abspath(__file__))
PATH = os.path.join(HERE, "golden", "golden_jlmini_v1.json")
ADAPTIVE = ("GPUSimpleATsit5", "GPUSimpleAVern7", "GPUSimpleAVern9")
ALWAYS_EVERYSTEP = ("GPUSimpleRK4", "GPUSimpleEuler")
F01 = np.float32(0.1)


RANDOM_PATH = os.path.join(HERE, "golden", "golden_jlmini_random_v1.json")
# the adaptive BASELINE configurations (whole accepted-step sequences): oracle/jlmini/gen_golden_configs.py
CONFIGS_PATH = os.path.join(HERE, "golden", "golden_jlmini_configs_v1.json")


def load_cases(path=PATH):
    with open(path) as fh:
        return json.load(fh)["cases"]


def _unhex(words, dtype):
    it = np.uint64 if np.dtype(dtype) == np.float64 else np.uint32
    return np.array([int(w, 16) for w in words], dtype=it).view(dtype)


def expected(case):
    """(t [n_out], u [n_out, N]) as the reference's solve returned them."""
    dtype = np.dtype(case["dtype"]).type
    n = case["n_out"]
    u = _unhex(case["u"], dtype).reshape(n, -1)
    t = _unhex(case["t"], np.dtype(case["t_dtype"]).type)
    return t, u


def case_inputs(case):
    dtype = np.dtype(case["dtype"]).type
    u0 = np.atleast_1d(np.asarray(case["u0"], dtype=dtype))
    p = np.asarray(case["p"], dtype=dtype)
    kw = case["kw"]
    # reference defaults: dt = 0.1f0, abstol = 1f-6, reltol = 1f-3, save_everystep = true
    dt = dtype(kw["dt"]) if "dt" in kw else dtype(F01)
    abstol = dtype(kw["abstol"]) if "abstol" in kw else dtype(np.float32(1e-6))
    reltol = dtype(kw["reltol"]) if "reltol" in kw else dtype(np.float32(1e-3))
    saveat = np.asarray(kw["saveat"], dtype=dtype) if "saveat" in kw else None
    if case["alg"] in ALWAYS_EVERYSTEP:
        kind = "everystep"
    elif saveat is not None:
        kind = "saveat"
    elif kw.get("save_everystep", True):
        kind = "everystep"
    else:
        kind = "endpoint"
    t0, tf = dtype(case["tspan"][0]), dtype(case["tspan"][1])
    return dict(dtype=dtype, u0=u0, p=p, dt=dt, abstol=abstol, reltol=reltol, saveat=saveat, kind=kind, t0=t0, tf=tf)


def oracle_run(case):
    from simplediffeq_b200 import jl_range
    a = case_inputs(case)
    dtype = a["dtype"]
    alg = C.ALG_NAMES[case["alg"]]
    kw = {}
    if case["alg"] in ADAPTIVE:
        kw.update(abstol=float(a["abstol"]), reltol=float(a["reltol"]), want_t=True)
    else:
        kw.update(tgrid=jl_range(a["t0"], a["dt"], a["tf"], dtype))
    if a["kind"] == "saveat":
        kw.update(saveat=a["saveat"])
    elif a["kind"] == "everystep":
        kw.update(save_mode=O.SAVE_EVERYSTEP, want_t=True)
        if case["alg"] in ADAPTIVE:
            kw.update(max_out=case.get("n_out", 1) + 2)     # capacity; the oracle reports how many it wrote
    r = O.solve(case["system"], alg, a["u0"][None, :], a["p"][None, :], float(a["t0"]), float(a["tf"]), float(a["dt"]),
                dtype=dtype, **kw)
    return r, a["kind"]
