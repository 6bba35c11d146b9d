"""Pin the SimpleEM oracle (oracle/oracle_em.cpp) and the host layer's step count / time grid against
outputs of the reference's OWN SOURCE TEXT (src/euler_maruyama.jl:46-94 executed by oracle/jlmini with the
normals supplied; tests/golden/golden_jlmini_em_v1.json, generator oracle/jlmini/gen_golden_em.py).
Bit for bit: every state, every time, the number of outputs, and Julia's InexactError."""
import os

import numpy as np
import pytest

import common as C
import oracle_lib as O
from jlmini_em_cases import load_cases, inputs, expected

CASES = load_cases()


@pytest.mark.parametrize("case", CASES, ids=[c["name"] for c in CASES])
def test_em_oracle_reproduces_reference_source_execution(case, oracle, sde):
    T, u0, p, t0, tf, dt = inputs(case)
    if "error" in case:
        assert case["error"] == "InexactError"
        with pytest.raises(ValueError, match="InexactError"):
            sde.em.em_steps((t0, tf), dt, T)
        return
    exp_t, exp_u, z = expected(case)
    n_steps = sde.em.em_steps((t0, tf), dt, T)
    assert n_steps + 1 == case["n_out"]                       # n = Int((tspan[2] - tspan[1]) / dt) + 1
    assert C.bits_equal(sde.em.em_times((t0, tf), dt, T), exp_t)
    got = O.em_solve(case["system"], u0[:, None], p[:, None], float(t0), float(dt), n_steps,
                     np.ascontiguousarray(z[:, :, None]))      # noise [n_steps][M][n_traj = 1]
    assert got.shape == (1, n_steps + 1, u0.size)
    assert C.bits_equal(np.ascontiguousarray(got[0]), exp_u), \
        "max ulp diff %d" % C.max_ulp_diff(np.ascontiguousarray(got[0]), exp_u)


def test_em_fixture_covers_scalar_vector_both_dtypes_and_the_error():
    names = {c["name"] for c in CASES}
    assert {"reftest_scalar_64", "reftest_svector_64", "reftest_scalar_32", "inexact_error_64"} <= names
    assert {c["system"] for c in CASES} == {"gbm", "linadd1", "linadd2", "ou"}
    assert len(CASES) >= 16


@pytest.mark.skipif(not os.path.isfile("/root/reference/src/euler_maruyama.jl"), reason="reference tree not present")
def test_em_fixture_is_reproducible_from_the_reference_tree():
    import importlib.util
    import json
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    spec = importlib.util.spec_from_file_location("gen_golden_em", os.path.join(root, "oracle", "jlmini", "gen_golden_em.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    fresh = [mod.run(c) for c in mod.CASES]
    assert json.loads(json.dumps(fresh)) == CASES
