"""CPU checks that pin the SimpleEM oracle (oracle/oracle_em.cpp).  The reference's own SimpleEM tests
(test/simpleem_tests.jl) assert only the time grid and the number of states; those are restated
here, and the step arithmetic / noise specification are pinned by known answers and closed forms."""
import numpy as np
import pytest

import oracle_lib as O


def test_philox4x32_10_random123_known_answers():
    # Random123 kat_vectors, philox4x32 10 rounds
    kat = [(([0] * 4, [0] * 2), "6627e8d5 e169c58d bc57ac4c 9b00dbd8"),
           (([0xffffffff] * 4, [0xffffffff] * 2), "408f276d 41c83b0e a20bc7c6 6d5451fd"),
           (([0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344], [0xa4093822, 0x299f31d0]),
            "d16cfe09 94fdcceb 5001e420 24126ea1")]
    for (ctr, key), want in kat:
        assert " ".join("%08x" % x for x in O.philox4x32_10(ctr, key)) == want


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
def test_noise_spec_is_standard_normal_and_partition_independent(dtype):
    from scipy import stats
    z = O.em_normals(dtype, seed=20261017, traj_offset=0, n_traj=4096, n_steps=64, M=1).astype(np.float64)
    flat = z.ravel()
    assert abs(flat.mean()) < 4 / np.sqrt(flat.size)
    assert abs(flat.var() - 1) < 4 * np.sqrt(2 / flat.size)
    assert stats.kstest(flat, "norm").pvalue > 1e-3
    # independence across steps / trajectories (lag-1 correlations)
    assert abs(np.corrcoef(z[:-1, 0].ravel(), z[1:, 0].ravel())[0, 1]) < 0.01
    assert abs(np.corrcoef(z[:, 0, :-1].ravel(), z[:, 0, 1:].ravel())[0, 1]) < 0.01
    # counter = GLOBAL trajectory index: a shard starting at trajectory 1000 sees the same numbers
    part = O.em_normals(dtype, 20261017, 1000, 96, 64, 1)
    assert np.array_equal(part, O.em_normals(dtype, 20261017, 0, 4096, 64, 1)[:, :, 1000:1096])
    # the linear index q = step*M + m: M = 3 re-slices the same stream
    z3 = O.em_normals(dtype, 20261017, 0, 8, 64, 3)
    z1 = O.em_normals(dtype, 20261017, 0, 8, 192, 1)
    assert np.array_equal(z3.reshape(192, 8), z1.reshape(192, 8))
    # another seed is another stream
    assert not np.array_equal(O.em_normals(dtype, 1, 0, 8, 8, 1), O.em_normals(dtype, 2, 0, 8, 8, 1))


def test_reference_simpleem_tests_restated():
    """test/simpleem_tests.jl:10-14,18-20,51-54: dt = 0.25 on (0,1) gives 5 states at t = 0:0.25:1.0 for the
    scalar, SVector{2} and 2x4 non-diagonal problems."""
    import simplediffeq_b200 as S
    assert S.em_steps((0.0, 1.0), 0.25) == 4
    assert S.em_times((0.0, 1.0), 0.25).tolist() == [0.0, 0.25, 0.5, 0.75, 1.0]
    with pytest.raises(ValueError):          # Int(3.3333333333333335) is an InexactError in Julia
        S.em_steps((0.0, 1.0), 0.3)
    rng = np.random.default_rng(0)
    for system, u0 in (("linadd1", [0.5]), ("linadd2", [0.1, 0.2]), ("nondiag2x4", [1.0, 1.0])):
        N, NP, M, diag = O.em_dims(system)
        p = {"linadd1": [2.0, 1.0], "linadd2": [2.0, 1.0], "nondiag2x4": [1.01]}[system]
        z = rng.standard_normal((4, M, 1))
        out = O.em_solve(system, np.array(u0).reshape(N, 1), np.array(p).reshape(NP, 1), 0.0, 0.25, 4, z)
        assert out.shape == (1, 5, N) and np.all(np.isfinite(out))
        assert out[0, 0].tolist() == u0


def test_noise_free_limit_is_forward_euler_bit_for_bit():
    """With dW = 0 the scalar EM step muladd(sqdt*g, 0, muladd(f, dt, uprev)) is GPUSimpleEuler's
    muladd(dt, f, uprev): compare with the ODE oracle on f = p1*u."""
    n, steps, dt = 64, 100, 0.01
    mu = np.linspace(-2, 2, n)
    u0 = np.full((1, n), 0.75)
    p = np.stack([mu, np.full(n, 0.3)])
    em = O.em_solve("gbm", u0, p, 0.0, dt, steps, np.zeros((steps, 1, n)))
    import simplediffeq_b200 as S
    eu = O.solve("scalargrowth", "Euler", u0.T.copy(), mu.reshape(n, 1).copy(), 0.0, 1.0, dt,
                 tgrid=S.jl_range(0.0, dt, 1.0), save_mode=O.SAVE_EVERYSTEP)
    assert em[:, :, 0].tobytes() == np.ascontiguousarray(eu.u[:, :steps + 1, 0]).tobytes()


def test_strong_order_one_half_against_exact_gbm_path():
    """dX = mu X dt + sigma X dW has X_T = X_0 exp((mu - sigma^2/2) T + sigma W_T); Euler-Maruyama
    converges strongly with order 1/2 (the reference's docstring, src/euler_maruyama.jl:41)."""
    rng = np.random.default_rng(5)
    n, T, mu, sigma = 4000, 1.0, 1.5, 1.0
    fine = 2 ** 10
    zf = rng.standard_normal((fine, 1, n))
    WT = np.sqrt(T / fine) * zf.sum(axis=0)[0]
    exact = np.exp((mu - 0.5 * sigma ** 2) * T + sigma * WT)
    u0 = np.ones((1, n)); p = np.stack([np.full(n, mu), np.full(n, sigma)])
    errs = []
    for k in (4, 5, 6, 7, 8):
        steps = 2 ** k
        # coarse increments = sums of fine ones, renormalised to unit variance
        z = zf.reshape(steps, fine // steps, 1, n).sum(axis=1) / np.sqrt(fine // steps)
        out = O.em_solve("gbm", u0, p, 0.0, T / steps, steps, z)
        errs.append(np.mean(np.abs(out[:, -1, 0] - exact)))
    slope = np.polyfit(np.log(T / 2.0 ** np.arange(4, 9)), np.log(errs), 1)[0]
    assert 0.4 < slope < 0.65, (slope, errs)


def test_weak_mean_and_one_step_covariance():
    rng = np.random.default_rng(11)
    n, steps, dt = 200_000, 16, 1 / 16
    u0 = np.ones((1, n)); p = np.stack([np.full(n, 0.1), np.full(n, 0.2)])
    out = O.em_solve("gbm", u0, p, 0.0, dt, steps, rng.standard_normal((steps, 1, n)), n_threads=8)
    xT = out[:, -1, 0]
    assert abs(xT.mean() - (1 + 0.1 * dt) ** steps) < 4 * xT.std() / np.sqrt(n)      # E[EM] is exact
    # non-diagonal: one step from u0 has covariance dt * G G^T
    u0 = np.ones((2, n)); p = np.full((1, n), 1.01)
    one = O.em_solve("nondiag2x4", u0, p, 0.0, dt, 1, rng.standard_normal((1, 4, n)), n_threads=8)[:, 1, :]
    G = np.array([[0.3, 0.6, 0.9, 0.12], [1.2, 0.2, 0.3, 1.8]])
    cov = np.cov(one.T)
    assert np.allclose(cov, dt * G @ G.T, rtol=0.03)
    assert np.allclose(one.mean(axis=0), 1 + 1.01 * dt, atol=4 * np.sqrt(cov.diagonal() / n))


def test_vector_diagonal_and_scalar_forms_agree_to_rounding():
    """(:76-77) vs (:79-80): same mathematics, different @muladd placement."""
    rng = np.random.default_rng(2)
    n, steps, dt = 50, 64, 1 / 64
    z = rng.standard_normal((steps, 2, n))
    p = np.stack([np.full(n, 2.0), np.full(n, 1.0)])
    v = O.em_solve("linadd2", np.full((2, n), 0.5), p, 0.0, dt, steps, z)
    s0 = O.em_solve("linadd1", np.full((1, n), 0.5), p, 0.0, dt, steps, z[:, :1])
    s1 = O.em_solve("linadd1", np.full((1, n), 0.5), p, 0.0, dt, steps, z[:, 1:])
    assert np.allclose(v[:, :, 0], s0[:, :, 0], rtol=1e-13, atol=1e-14)
    assert np.allclose(v[:, :, 1], s1[:, :, 0], rtol=1e-13, atol=1e-14)
    assert not np.array_equal(v[:, :, 0], s0[:, :, 0])      # but not the same rounding


def _golden_em():
    import json, os
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "golden_em_v1.json")
    return json.load(open(path))["cases"]


def _unhex(h, dtype, shape):
    return np.frombuffer(bytes.fromhex("".join(h)), dtype=dtype).reshape(shape).copy()


def golden_em_arrays(case):
    T = np.dtype(case["dtype"])
    N, NP, M, _ = O.em_dims(case["system"])
    n, steps = case["n"], case["n_steps"]
    return (_unhex(case["u0"], T, (N, n)), _unhex(case["p"], T, (NP, n)), _unhex(case["noise"], T, (steps, M, n)),
            _unhex(case["out"], T, (n, steps + 1, N)))


def test_oracle_em_reproduces_committed_golden_vectors():
    cases = _golden_em()
    assert len(cases) == 10
    for case in cases:
        u0, p, z, want = golden_em_arrays(case)
        got = O.em_solve(case["system"], u0, p, case["t0"], case["dt"], case["n_steps"], z)
        assert got.tobytes() == want.tobytes(), (case["system"], case["dtype"])
