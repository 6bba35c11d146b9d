"""GPU parity tests of the SimpleEM path (src/euler_maruyama.jl:48-94): the CUDA kernels through the
C ABI vs oracle/oracle_em.cpp.  Given the same increments the two must agree BIT FOR BIT (every
operation is IEEE and identically ordered); the Philox/Box-Muller stream is checked against the
oracle's independent restatement of the noise specification and statistically."""
import numpy as np
import pytest

import oracle_lib as O

pytestmark = pytest.mark.gpu

SYSTEMS = {"gbm": ([1.0], [0.1, 0.2]), "linadd1": ([0.5], [2.0, 1.0]), "linadd2": ([0.1, 0.2], [2.0, 1.0]),
           "ou": ([0.3], [1.5, 1.0, 0.4]), "nondiag2x4": ([1.0, 1.0], [1.01])}


def _problem(system, n, dtype, seed):
    rng = np.random.default_rng(seed)
    u0, p = SYSTEMS[system]
    u0 = (np.array(u0)[:, None] * (1 + 0.2 * rng.uniform(-1, 1, (len(u0), n)))).astype(dtype)
    p = (np.array(p)[:, None] * (1 + 0.2 * rng.uniform(-1, 1, (len(p), n)))).astype(dtype)
    return u0, p


def _bits(a, b):
    a, b = np.ascontiguousarray(a), np.ascontiguousarray(b)
    return a.shape == b.shape and a.dtype == b.dtype and a.tobytes() == b.tobytes()


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
@pytest.mark.parametrize("system", list(SYSTEMS))
def test_em_provided_noise_bit_exact(sde, system, dtype):
    n, steps, dt = 777, 40, 1 / 32
    sysm = getattr(sde.sde_systems, system)
    u0, p = _problem(system, n, dtype, 3)
    z = np.random.default_rng(4).standard_normal((steps, sysm.n_noise, n)).astype(dtype)
    want = O.em_solve(system, u0, p, 0.25, dt, steps, z)                      # [n][steps+1][N]
    tm = sde.solve_em_arrays(sysm, u0, p, 0.25, dt, steps, noise=z, layout=0)
    soa = sde.solve_em_arrays(sysm, u0, p, 0.25, dt, steps, noise=z, layout=1)
    end = sde.solve_em_arrays(sysm, u0, p, 0.25, dt, steps, noise=z, save_mode=0)
    assert _bits(tm, want)
    assert _bits(soa.transpose(2, 0, 1), want)
    assert _bits(end.T, want[:, -1, :])
    assert np.all(np.isfinite(want))


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
def test_em_philox_stream_matches_noise_specification(sde, dtype):
    """sde_em_noise (what the kernels consume) vs the oracle's restatement of the spec (libm log / sincos)."""
    from scipy import stats
    n, steps, M = 1000, 33, 3
    got = sde.em_noise(dtype, 20261017, n, steps, M, traj_offset=12345)
    want = O.em_normals(dtype, 20261017, 12345, n, steps, M)
    tol = 1e-13 if dtype == np.float64 else 3e-6
    assert np.max(np.abs(got.astype(np.float64) - want.astype(np.float64))) < tol
    big = sde.em_noise(dtype, 7, 1 << 16, 16, 1).astype(np.float64).ravel()
    assert abs(big.mean()) < 4 / np.sqrt(big.size) and abs(big.var() - 1) < 4 * np.sqrt(2 / big.size)
    assert stats.kstest(big, "norm").pvalue > 1e-3
    assert abs(stats.skew(big)) < 0.02 and abs(stats.kurtosis(big)) < 0.04


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
@pytest.mark.parametrize("system", ["gbm", "linadd2", "nondiag2x4"])
def test_em_philox_solve_equals_oracle_on_dumped_noise(sde, system, dtype):
    n, steps, dt, seed = 515, 25, 0.04, 99
    sysm = getattr(sde.sde_systems, system)
    u0, p = _problem(system, n, dtype, 8)
    z = sde.em_noise(dtype, seed, n, steps, sysm.n_noise)
    want = O.em_solve(system, u0, p, 0.0, dt, steps, z)
    got = sde.solve_em_arrays(sysm, u0, p, 0.0, dt, steps, seed=seed)
    assert _bits(got, want)
    other = sde.solve_em_arrays(sysm, u0, p, 0.0, dt, steps, seed=seed + 1)
    assert not _bits(other, want)


def test_em_results_do_not_depend_on_partitioning(sde, monkeypatch):
    """The Philox counter is the GLOBAL trajectory index: pieces, offsets and devices cannot change a bit."""
    from simplediffeq_b200 import _lib
    n, steps, dt, seed = 1003, 20, 0.05, 5
    sysm = sde.sde_systems.gbm
    u0, p = _problem("gbm", n, np.float64, 1)
    whole = sde.solve_em_arrays(sysm, u0, p, 0.0, dt, steps, seed=seed, devices=[0])
    monkeypatch.setenv("SDE_TUNE_PIECE", "96")
    pieces = sde.solve_em_arrays(sysm, u0, p, 0.0, dt, steps, seed=seed, devices=[0])
    zp = np.random.default_rng(0).standard_normal((steps, 1, n))
    pieces_prov = sde.solve_em_arrays(sysm, u0, p, 0.0, dt, steps, noise=zp, layout=1, devices=[0])
    monkeypatch.delenv("SDE_TUNE_PIECE")
    assert _bits(whole, pieces)
    assert _bits(pieces_prov, sde.solve_em_arrays(sysm, u0, p, 0.0, dt, steps, noise=zp, layout=1, devices=[0]))
    # a shard solved on its own with traj_offset = its first global index
    part = sde.solve_em_arrays(sysm, np.ascontiguousarray(u0[:, 400:700]), np.ascontiguousarray(p[:, 400:700]), 0.0, dt,
                               steps, seed=seed, traj_offset=400)
    assert _bits(part, whole[400:700])
    ndev = _lib.device_count()
    if ndev >= 2:
        many = sde.solve_em_arrays(sysm, u0, p, 0.0, dt, steps, seed=seed, devices=list(range(min(ndev, 8))))
        assert _bits(many, whole)


def test_em_user_sde_nvrtc_matches_builtin(sde):
    src = """
    __device__ void rhs(real* f, const real* u, const real* p, real t) { f[0] = p[0] * u[0]; }
    __device__ void noise(real* g, const real* u, const real* p, real t) { g[0] = p[1] * u[0]; }
    """
    user = sde.CudaSDE(src, 1, 2)
    n, steps = 300, 16
    for dtype in (np.float64, np.float32):
        u0, p = _problem("gbm", n, dtype, 21)
        a = sde.solve_em_arrays(user, u0, p, 0.0, 1 / 16, steps, seed=3)
        b = sde.solve_em_arrays(sde.sde_systems.gbm, u0, p, 0.0, 1 / 16, steps, seed=3)
        assert _bits(a, b)
    src2 = """
    __device__ void rhs(real* f, const real* u, const real* p, real t) { f[0] = p[0] * u[0]; f[1] = p[0] * u[1]; }
    __device__ void noise(real* g, const real* u, const real* p, real t) {
      g[0] = real(0.3) * u[0]; g[1] = real(0.6) * u[0]; g[2] = real(0.9) * u[0]; g[3] = real(0.12) * u[0];
      g[4] = real(1.2) * u[1]; g[5] = real(0.2) * u[1]; g[6] = real(0.3) * u[1]; g[7] = real(1.8) * u[1];
    }
    """
    user2 = sde.CudaSDE(src2, 2, 1, n_noise=4, diagonal=False)
    u0, p = _problem("nondiag2x4", n, np.float64, 22)
    assert _bits(sde.solve_em_arrays(user2, u0, p, 0.0, 0.25, 4, seed=1),
                 sde.solve_em_arrays(sde.sde_systems.nondiag2x4, u0, p, 0.0, 0.25, 4, seed=1))
    with pytest.raises(Exception):
        sde.CudaSDE("__device__ void rhs(real* f, const real* u, const real* p, real t) { f[0] = 1; }", 1, 0)


def test_em_ensemble_moments_full_size(sde):
    """1 Mi GBM paths, device resident, endpoint only: E[X_n] = x0 (1 + mu dt)^n exactly for Euler-Maruyama,
    Var[X_n] = x0^2 (((1 + mu dt)^2 + sigma^2 dt)^n - (1 + mu dt)^(2n)); and the OU stationary spread."""
    import torch
    n, steps, dt, mu, sigma = 1 << 20, 64, 1 / 64, 0.1, 0.2
    dev = torch.device("cuda:0")
    u0 = torch.ones((1, n), dtype=torch.float64, device=dev)
    p = torch.empty((2, n), dtype=torch.float64, device=dev); p[0] = mu; p[1] = sigma
    x = sde.solve_em_device(sde.sde_systems.gbm, u0, p, 0.0, dt, steps, seed=2026).cpu().numpy()[0]
    m = (1 + mu * dt) ** steps
    v = ((1 + mu * dt) ** 2 + sigma ** 2 * dt) ** steps - m ** 2
    assert abs(x.mean() - m) < 4 * np.sqrt(v / n)
    assert abs(x.var() - v) < 0.01 * v
    # different seeds are independent streams
    y = sde.solve_em_device(sde.sde_systems.gbm, u0, p, 0.0, dt, steps, seed=2027).cpu().numpy()[0]
    assert abs(np.corrcoef(x, y)[0, 1]) < 5 / np.sqrt(n)
    # host path == device path
    h = sde.solve_em_arrays(sde.sde_systems.gbm, np.ones((1, 4096)), np.tile([[mu], [sigma]], (1, 4096)), 0.0, dt, steps,
                            seed=2026, save_mode=0)
    assert _bits(h[0], x[:4096])


def test_em_python_mirror_reference_tests(sde):
    """test/simpleem_tests.jl restated through the mirrored interface."""
    prob = sde.SDEProblem(sde.sde_systems.linadd1, 0.5, (0.0, 1.0), p=[2.0, 1.0])      # f = 2u, g = 1
    sol = sde.solve(prob, sde.SimpleEM(), dt=0.25)
    assert sol.t.tolist() == [0.0, 0.25, 0.5, 0.75, 1.0] and len(sol.u) == 5 and sol.u[0, 0] == 0.5
    with pytest.raises(ValueError, match="dt required"):
        sde.solve(prob, sde.SimpleEM())
    prob2 = sde.SDEProblem(sde.sde_systems.linadd2, [0.1, 0.2], (0.0, 1.0), p=[2.0, 1.0])
    sol2 = sde.solve(prob2, sde.SimpleEM(), dt=0.25)
    assert sol2.u.shape == (5, 2) and sol2.u.dtype == np.float64
    prob3 = sde.SDEProblem(sde.sde_systems.nondiag2x4, [1.0, 1.0], (0.0, 1.0), p=[1.01])
    assert len(sde.solve(prob3, sde.SimpleEM(), dt=0.25).u) == 5
    ens = sde.EnsembleProblem(prob, prob_func=lambda pr, i, rep: sde.SDEProblem(pr.f, pr.u0 * i, pr.tspan, pr.p))
    es = sde.solve(ens, sde.SimpleEM(), dt=0.25, trajectories=7, seed=11)
    assert len(es) == 7 and es[2].u[0, 0] == 1.5 and es[2].u.shape == (5, 1)
    again = sde.solve(ens, sde.SimpleEM(), dt=0.25, trajectories=7, seed=11)
    assert _bits(es.u_raw, again.u_raw)                                              # reproducible


def test_em_committed_golden_vectors(sde):
    """tests/golden/golden_em_v1.json (oracle outputs with their increments): the CUDA path reproduces them bit for bit."""
    from test_oracle_em import _golden_em, golden_em_arrays
    for case in _golden_em():
        u0, p, z, want = golden_em_arrays(case)
        got = sde.solve_em_arrays(getattr(sde.sde_systems, case["system"]), u0, p, case["t0"], case["dt"], case["n_steps"],
                                  noise=z, layout=0)
        assert _bits(got, want), (case["system"], case["dtype"])


# ---- CUDA path vs the reference's OWN SOURCE TEXT (src/euler_maruyama.jl executed by oracle/jlmini with
# the normals supplied; tests/golden/golden_jlmini_em_v1.json).  Reads like test/simpleem_tests.jl:
# solve(SDEProblem(f, g, u0, tspan, p), SimpleEM(); dt) -> sol.t, sol.u.
import jlmini_em_cases as JE  # noqa: E402

_JE_CASES = JE.load_cases()


@pytest.mark.parametrize("case", _JE_CASES, ids=[c["name"] for c in _JE_CASES])
def test_em_cuda_path_vs_reference_source_execution(sde, case):
    T, u0, p, t0, tf, dt = JE.inputs(case)
    prob = sde.SDEProblem(getattr(sde.sde_systems, case["system"]), u0, (t0, tf), p)
    if "error" in case:
        with pytest.raises(ValueError, match="InexactError"):      # Julia: Int(3.3333333333333335)
            sde.solve_em(prob, sde.SimpleEM(), dt=dt)
        return
    exp_t, exp_u, z = JE.expected(case)
    sol = sde.solve_em(prob, sde.SimpleEM(), dt=dt, noise=z[:, :, None] if len(z) else None)
    assert len(sol.u) == case["n_out"] and len(sol.t) == case["n_out"]
    assert _bits(np.asarray(sol.t, dtype=T), exp_t)
    assert _bits(np.asarray(sol.u, dtype=T).reshape(exp_u.shape), exp_u)
