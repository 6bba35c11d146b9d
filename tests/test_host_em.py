"""CPU tests of the SimpleEM host side: registry, option validation, NVRTC compile of user SDEs (no device
needed), the mirrored interface's error behaviour, and that nothing is computed without a GPU."""
import ctypes

import numpy as np
import pytest


def test_em_builtin_registry(sde):
    dims = dict(gbm=(1, 2, 1, True), linadd1=(1, 2, 1, True), linadd2=(2, 2, 2, True), ou=(1, 3, 1, True),
                nondiag2x4=(2, 1, 4, False))
    for name, d in dims.items():
        s = getattr(sde.sde_systems, name)
        assert (s.n_state, s.n_param, s.n_noise, s.diagonal) == d
    with pytest.raises(Exception):
        sde.em.builtin_sde_system("no-such-sde")


def test_em_every_builtin_kernel_exists_and_options_are_validated(sde):
    from simplediffeq_b200 import _lib, em
    L = _lib.lib()
    for name in sde.sde_systems.names():
        h = getattr(sde.sde_systems, name)._handle
        for dtype in (np.float64, np.float32):
            for save in (0, 2):
                for noise in (0, 1):
                    o = em._options(dtype, 8, 0.0, 0.1, 10, save, 0, noise, 1, 0)
                    assert L.sde_em_system_prepare(h, ctypes.byref(o)) == 0, L.sde_last_error()
    h = sde.sde_systems.gbm._handle
    bad = em._options(np.float64, 8, 0.0, 0.1, 10, 1, 0, 0, 1, 0)            # saveat does not exist for SimpleEM
    assert L.sde_em_system_prepare(h, ctypes.byref(bad)) == -4
    bad = em._options(np.float64, 8, 0.0, 0.1, 10, 2, 0, 5, 1, 0)            # unknown noise mode
    assert L.sde_em_system_prepare(h, ctypes.byref(bad)) == -1
    bad = em._options(np.float64, 8, 0.0, 0.1, -1, 2, 0, 0, 1, 0)
    assert L.sde_em_system_prepare(h, ctypes.byref(bad)) == -1
    assert L.sde_em_system_prepare(None, ctypes.byref(bad)) == -1
    assert L.sde_trim() == 0                                                # no pools yet: nothing to do


def test_em_user_sde_compiles_for_sm100a(sde):
    from simplediffeq_b200 import _lib, em
    L = _lib.lib()
    src = """
    __device__ void rhs(real* f, const real* u, const real* p, real t) { f[0] = p[0] * u[0]; f[1] = -u[1] + t; }
    __device__ void noise(real* g, const real* u, const real* p, real t) {
      g[0] = p[1]; g[1] = 0; g[2] = p[1] * u[0]; g[3] = real(0.5); g[4] = 0; g[5] = u[1]; }
    """
    user = sde.CudaSDE(src, 2, 2, n_noise=3, diagonal=False)
    assert (user.n_state, user.n_param, user.n_noise, user.diagonal) == (2, 2, 3, False)
    for dtype in (np.float64, np.float32):
        for save, noise in ((0, 0), (2, 0), (2, 1)):
            o = em._options(dtype, 8, 0.0, 0.1, 10, save, 1, noise, 1, 0)
            assert L.sde_em_system_prepare(user._handle, ctypes.byref(o)) == 0, L.sde_last_error()
    with pytest.raises(_lib.SdeError) as e:                                  # `noise` missing
        sde.CudaSDE("__device__ void rhs(real* f, const real* u, const real* p, real t) { f[0] = u[0]; }", 1, 0)
    assert e.value.code == -3
    with pytest.raises(_lib.SdeError):                                       # diagonal noise needs n_noise == n_state
        sde.CudaSDE(src, 2, 2, n_noise=3, diagonal=True)


def test_em_interface_errors_mirror_the_reference(sde):
    prob = sde.SDEProblem(sde.sde_systems.gbm, 1.0, (0.0, 1.0), p=[0.1, 0.2])
    with pytest.raises(ValueError, match="dt required for SimpleEM"):        # src/euler_maruyama.jl:51
        sde.solve(prob, sde.SimpleEM())
    with pytest.raises(ValueError, match="InexactError"):                    # Int((1-0)/0.3), :66
        sde.solve(prob, sde.SimpleEM(), dt=0.3)
    with pytest.raises(TypeError):
        sde.SDEProblem(lambda u, p, t: u, 1.0, (0.0, 1.0))
    with pytest.raises(ValueError):
        sde.SDEProblem(sde.sde_systems.gbm, [1.0, 2.0], (0.0, 1.0), p=[0.1, 0.2])
    assert sde.em_steps((0.0, 10.0), 0.001) == 10000
    t = sde.em_times((0.5, 1.5), 0.125)
    assert t[0] == 0.5 and t[-1] == 1.5 and len(t) == 9
    t32 = sde.em_times((0.0, 1.0), 0.25, np.float32)
    assert t32.dtype == np.float32 and t32.tolist() == [0.0, 0.25, 0.5, 0.75, 1.0]


def test_em_no_cpu_fallback(sde):
    from simplediffeq_b200 import _lib
    if _lib.device_count() > 0:
        pytest.skip("a GPU is present")
    with pytest.raises(_lib.SdeError) as e:
        sde.solve_em_arrays(sde.sde_systems.gbm, np.ones((1, 8)), np.ones((2, 8)), 0.0, 0.1, 4, seed=1)
    assert e.value.code == -2
    with pytest.raises(_lib.SdeError):
        sde.em_noise(np.float64, 1, 8, 4, 1)
