"""Loader for tests/golden/golden_jlmini_em_v1.json (the reference's own SimpleEM source executed by
oracle/jlmini with supplied normals).  Test infrastructure."""
import json
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
PATH = os.path.join(HERE, "golden", "golden_jlmini_em_v1.json")


def load_cases():
    with open(PATH) as fh:
        return json.load(fh)["cases"]


def _unhex(words, dtype):
    it = np.uint64 if np.dtype(dtype) == np.float64 else np.uint32
    return np.array([int(w, 16) for w in words], dtype=it).view(dtype)


def inputs(case):
    """dtype, u0 [N], p [NP], t0, tf, dt (as dtype scalars)."""
    T = np.dtype(case["dtype"]).type
    u0 = np.atleast_1d(np.asarray(case["u0"], dtype=T))
    p = np.asarray(case["p"], dtype=T)
    return T, u0, p, T(case["tspan"][0]), T(case["tspan"][1]), T(case["dt"])


def expected(case):
    """t [n], u [n, N], noise [n-1, N] as the reference's solve produced / consumed them."""
    T = np.dtype(case["dtype"]).type
    n = case["n_out"]
    N = 1 if not isinstance(case["u0"], list) else len(case["u0"])
    return _unhex(case["t"], T), _unhex(case["u"], T).reshape(n, N), _unhex(case["noise"], T).reshape(n - 1, N)
