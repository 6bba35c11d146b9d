"""Host-side mirror of the reference's SimpleEM interface (src/euler_maruyama.jl:48-94):

    solve(prob::SDEProblem, SimpleEM(); dt = error("dt required for SimpleEM"))
    solve(EnsembleProblem(prob; prob_func), SimpleEM(); dt, trajectories)        [SciMLBase ensemble driver]

`prob.f` / `prob.g` must live on the device: a built-in SDE system (`sde_systems.gbm`, ...) or a `CudaSDE`
(CUDA-C source with `rhs` and `noise`, compiled by NVRTC).  The whole ensemble crosses the C ABI in one
`sde_em_solve` call.  The reference draws randn() from Julia's task-local RNG; here the increments are a
counter-based Philox stream selected by `seed` (reproducible, independent of how the ensemble is
sharded), or an explicit `noise` array of standard normals.
"""
import ctypes

import os

import numpy as np

from . import _lib
from .api import EnsembleProblem


class SDESystem:
    """Drift f(u,p,t) and diffusion g(u,p,t) available on the device."""

    def __init__(self, handle, name, n_state, n_param, n_noise, diagonal, owned):
        self._handle, self.name, self._owned = handle, name, owned
        self.n_state, self.n_param, self.n_noise, self.diagonal = n_state, n_param, n_noise, diagonal

    def __repr__(self):
        return "SDESystem(%s, n_state=%d, n_param=%d, n_noise=%d, %s)" % (
            self.name, self.n_state, self.n_param, self.n_noise, "diagonal" if self.diagonal else "non-diagonal")

    def __del__(self):
        try:
            if self._owned and self._handle:
                _lib.lib().sde_em_system_free(self._handle)
        except Exception:
            pass


def _dims(h):
    a, b, c, d = ctypes.c_int(), ctypes.c_int(), ctypes.c_int(), ctypes.c_int()
    _lib.check(_lib.lib().sde_em_system_dims(h, ctypes.byref(a), ctypes.byref(b), ctypes.byref(c), ctypes.byref(d)))
    return a.value, b.value, c.value, bool(d.value)


def builtin_sde_system(name):
    h = ctypes.c_void_p()
    _lib.check(_lib.lib().sde_em_system_builtin(name.encode(), ctypes.byref(h)))
    return SDESystem(h, name, *_dims(h), False)


class _SdeSystems:
    """Built-in SDE registry: sde_systems.gbm, .linadd1, .linadd2, .ou, .nondiag2x4."""
    _names = ("gbm", "linadd1", "linadd2", "ou", "nondiag2x4")

    def __init__(self):
        self._cache = {}

    def __getattr__(self, name):
        if name.startswith("_") or name not in self._names:
            raise AttributeError(name)
        if name not in self._cache:
            self._cache[name] = builtin_sde_system(name)
        return self._cache[name]

    def names(self):
        return self._names


sde_systems = _SdeSystems()


def CudaSDE(src, n_state, n_param, n_noise=None, diagonal=True):
    """User SDE as CUDA C++ source defining `__device__ void rhs(real* f, const real* u, const real* p, real t)`
    and `__device__ void noise(real* g, const real* u, const real* p, real t)` (diagonal: n_state values; else an
    n_state x n_noise row-major matrix = the reference's noise_rate_prototype shape)."""
    n_noise = n_state if n_noise is None else n_noise
    h = ctypes.c_void_p()
    log = ctypes.create_string_buffer(16384)
    rc = _lib.lib().sde_em_system_nvrtc(src.encode(), n_state, n_param, n_noise, 1 if diagonal else 0,
                                        ctypes.byref(h), log, len(log))
    if rc != _lib.SDE_OK:
        raise _lib.SdeError(rc, _lib.lib().sde_last_error().decode("utf-8", "replace"))
    return SDESystem(h, "user", n_state, n_param, n_noise, diagonal, True)


class SimpleEM:
    """Fixed-step Euler-Maruyama (src/euler_maruyama.jl:45)."""

    def __repr__(self):
        return "SimpleEM()"


class SDEProblem:
    """SDEProblem{false}(f, g, u0, tspan, p): `system` carries f and g."""

    def __init__(self, system, u0, tspan, p=None):
        if not isinstance(system, SDESystem):
            raise TypeError("f/g must be a built-in SDE system or a CudaSDE (no host callables on the GPU path)")
        u0 = np.atleast_1d(np.asarray(u0))
        self.dtype = np.dtype(np.float32) if u0.dtype == np.float32 else np.dtype(np.float64)
        self.f = system
        self.u0 = u0.astype(self.dtype)
        self.tspan = (tspan[0], tspan[1])
        self.p = np.zeros(0, self.dtype) if p is None else np.atleast_1d(np.asarray(p, dtype=self.dtype))
        if self.u0.shape != (system.n_state,):
            raise ValueError("u0 must have %d components" % system.n_state)
        if self.p.shape != (system.n_param,):
            raise ValueError("p must have %d components" % system.n_param)


def em_steps(tspan, dt, dtype=np.float64):
    """n - 1 with n = Int((tspan[2] - tspan[1]) / dt) + 1 (src/euler_maruyama.jl:66).  Julia's Int() of a
    non-integer float is an InexactError: ValueError here."""
    T = np.dtype(dtype).type
    q = (T(tspan[1]) - T(tspan[0])) / T(dt)
    if not np.isfinite(q) or q != np.floor(q):
        raise ValueError("InexactError: Int(%r)" % float(q))
    if q < 0:
        raise ValueError("tspan[2] < tspan[1]")
    return int(q)


def _round_fraction(fr, T):
    """Round an exact rational to T (float64 / float32), nearest-even, in ONE rounding."""
    x = float(fr)                      # CPython rounds int/int correctly to double
    if T is np.float64:
        return np.float64(x)
    c = np.float32(x)                  # may be double-rounded: pick the best of c and its neighbours exactly
    from fractions import Fraction
    best, best_err = None, None
    for cand in (np.nextafter(c, np.float32(-np.inf)), c, np.nextafter(c, np.float32(np.inf))):
        if not np.isfinite(cand):
            continue
        err = abs(Fraction(float(cand)) - fr)
        even = (int(np.float32(cand).view(np.uint32)) & 1) == 0
        if best is None or err < best_err or (err == best_err and even):
            best, best_err = cand, err
    return np.float32(best)


def em_times(tspan, dt, dtype=np.float64):
    """t = [tspan[1] + i*dt for i in 0:n-1] under @muladd = muladd(i, dt, tspan[1]) (src/euler_maruyama.jl:68):
    one fused multiply-add per element, evaluated exactly in rational arithmetic and rounded once."""
    from fractions import Fraction
    T = np.dtype(dtype).type
    n = em_steps(tspan, dt, dtype) + 1
    fdt, ft0 = Fraction(float(T(dt))), Fraction(float(T(tspan[0])))
    return np.array([_round_fraction(Fraction(i) * fdt + ft0, T) for i in range(n)], dtype=T)


class EMEnsembleSolution:
    """`.u_raw` (layout-dependent), `.t` (shared by all trajectories); indexing gives (t, u[n, n_state])."""

    def __init__(self, **kw):
        self.__dict__.update(kw)

    def __len__(self):
        return self.n_traj

    def u_of(self, i):
        if self.save_mode == _lib.SAVE_ENDPOINT:
            return np.stack([self.u0_soa[:, i], self.u_raw[:, i]])
        return self.u_raw[i] if self.layout == _lib.LAYOUT_TRAJ_MAJOR else self.u_raw[:, :, i]

    def __getitem__(self, i):
        t = self.t if self.save_mode != _lib.SAVE_ENDPOINT else self.t[[0, -1]]
        return EMSolution(t, self.u_of(i))


class EMSolution:
    def __init__(self, t, u):
        self.t, self.u = t, u
        self.retcode = "Default"

    def __len__(self):
        return len(self.t)


def _options(dtype, n, t0, dt, n_steps, save_mode, layout, noise_mode, seed, traj_offset):
    o = _lib.SdeEmOptions()
    o.dtype = _lib.SDE_F64 if np.dtype(dtype) == np.float64 else _lib.SDE_F32
    o.save_mode, o.layout, o.noise_mode = save_mode, layout, noise_mode
    o.n_traj, o.t0, o.dt, o.n_steps = n, float(t0), float(dt), int(n_steps)
    o.seed, o.traj_offset = int(seed) & 0xFFFFFFFFFFFFFFFF, int(traj_offset)
    return o


def solve_em_arrays(system, u0_soa, p_soa, t0, dt, n_steps, *, seed=0, noise=None, save_mode=_lib.SAVE_EVERYSTEP,
                    layout=_lib.LAYOUT_TRAJ_MAJOR, devices=None, traj_offset=0):
    """Array-level entry: SoA host arrays in, raw array out, ONE sde_em_solve call.
    noise: None (Philox stream of `seed`) or standard normals [n_steps, n_noise, n_traj]."""
    dtype = u0_soa.dtype
    N, n = u0_soa.shape
    o = _options(dtype, n, t0, dt, n_steps, save_mode, layout,
                 _lib.NOISE_PROVIDED if noise is not None else _lib.NOISE_PHILOX, seed, traj_offset)
    if save_mode == _lib.SAVE_ENDPOINT:
        out = np.empty((N, n), dtype=dtype)
    else:
        out = np.empty((n, n_steps + 1, N) if layout == _lib.LAYOUT_TRAJ_MAJOR else (n_steps + 1, N, n), dtype=dtype)
    u0c = np.ascontiguousarray(u0_soa)
    pc = np.ascontiguousarray(p_soa, dtype=dtype)
    zc = None
    if noise is not None:
        zc = np.ascontiguousarray(noise, dtype=dtype)
        if zc.shape != (n_steps, system.n_noise, n):
            raise ValueError("noise must have shape (n_steps, n_noise, n_traj)")
    dev, ndev = None, 0
    if devices is not None:
        dev, ndev = (ctypes.c_int * len(devices))(*devices), len(devices)
    rc = _lib.lib().sde_em_solve(system._handle, ctypes.byref(o), u0c.ctypes.data, pc.ctypes.data if pc.size else None,
                                 zc.ctypes.data if zc is not None and zc.size else None, out.ctypes.data, dev, ndev)
    _lib.check(rc)
    return out


def em_noise(dtype, seed, n_traj, n_steps, n_noise, traj_offset=0):
    """The standard normals a Philox solve consumes: [n_steps, n_noise, n_traj]."""
    o = _options(dtype, n_traj, 0.0, 0.0, n_steps, 0, 0, _lib.NOISE_PHILOX, seed, traj_offset)
    out = np.empty((n_steps, n_noise, n_traj), dtype=dtype)
    _lib.check(_lib.lib().sde_em_noise(ctypes.byref(o), n_noise, out.ctypes.data))
    return out


def solve_em_device(system, d_u0, d_p, t0, dt, n_steps, *, seed=0, d_noise=None, save_mode=_lib.SAVE_ENDPOINT,
                    layout=_lib.LAYOUT_SOA, out=None, traj_offset=0, stream=None, sync=True):
    """Device-resident entry (torch CUDA tensors, SoA [n_state, n]); torch only provides memory / stream."""
    import torch
    assert d_u0.is_cuda and d_u0.is_contiguous()
    dtype = np.dtype(np.float64) if d_u0.dtype == torch.float64 else np.dtype(np.float32)
    N, n = d_u0.shape
    o = _options(dtype, n, t0, dt, n_steps, save_mode, layout,
                 _lib.NOISE_PROVIDED if d_noise is not None else _lib.NOISE_PHILOX, seed, traj_offset)
    if out is None:
        shape = (N, n) if save_mode == _lib.SAVE_ENDPOINT else (
            (n, n_steps + 1, N) if layout == _lib.LAYOUT_TRAJ_MAJOR else (n_steps + 1, N, n))
        out = torch.empty(shape, dtype=d_u0.dtype, device=d_u0.device)
    st = torch.cuda.current_stream(d_u0.device).cuda_stream if stream is None else stream
    with torch.cuda.device(d_u0.device):
        rc = _lib.lib().sde_em_solve_device(system._handle, ctypes.byref(o), d_u0.data_ptr(),
                                            d_p.data_ptr() if d_p is not None and d_p.numel() else None, n,
                                            d_noise.data_ptr() if d_noise is not None else None, n,
                                            out.data_ptr(), n, st, 0 if sync else 1)
    _lib.check(rc)
    return out


def solve_em(prob, alg=None, *, dt=None, trajectories=None, seed=None, noise=None, save_everystep=True,
             layout="traj_major", devices=None, **kwargs):
    """solve(prob, SimpleEM(); dt) / solve(EnsembleProblem(prob; prob_func), SimpleEM(); dt, trajectories).
    Unknown keywords are swallowed like the reference's `kwargs...`.  `seed`, `noise`, `save_everystep=False`
    (endpoint only), `layout` and `devices` are extensions.
    seed=None (default) draws a fresh 64-bit Philox key from the OS for every call, like the reference's `randn` on
    the task-local RNG: repeated solves give independent paths.  The key that was used is kept on the solution
    (`sol.seed`) so that a run can be reproduced by passing it back."""
    if seed is None and noise is None:
        seed = int.from_bytes(os.urandom(8), "little")
    elif seed is None:
        seed = 0
    if dt is None:
        raise ValueError("dt required for SimpleEM")        # src/euler_maruyama.jl:51
    single = isinstance(prob, SDEProblem)
    ens = EnsembleProblem(prob) if single else prob
    base = ens.prob
    if not isinstance(base, SDEProblem):
        raise TypeError("SimpleEM needs an SDEProblem")
    n = 1 if single else trajectories
    if n is None:
        raise ValueError("trajectories=... is required for an EnsembleProblem")
    sysm, dtype = base.f, base.dtype
    u0_soa = np.empty((sysm.n_state, n), dtype=dtype)
    p_soa = np.empty((sysm.n_param, n), dtype=dtype)
    if ens.u0s is not None or ens.ps is not None:
        u0_soa[:] = np.asarray(ens.u0s, dtype=dtype).T if ens.u0s is not None else base.u0[:, None]
        p_soa[:] = np.asarray(ens.ps, dtype=dtype).T if ens.ps is not None else base.p[:, None]
    else:
        for i in range(n):
            pi = ens.prob_func(base, i + 1, 1) if ens.prob_func is not None else base
            if pi.f is not sysm or tuple(pi.tspan) != tuple(base.tspan):
                raise ValueError("prob_func may change u0 and p only (one kernel per ensemble)")
            u0_soa[:, i] = pi.u0
            p_soa[:, i] = pi.p
    n_steps = em_steps(base.tspan, dt, dtype)
    save_mode = _lib.SAVE_EVERYSTEP if save_everystep else _lib.SAVE_ENDPOINT
    lay = _lib.LAYOUT_TRAJ_MAJOR if layout == "traj_major" else _lib.LAYOUT_SOA
    raw = solve_em_arrays(sysm, u0_soa, p_soa, base.tspan[0], dt, n_steps, seed=seed, noise=noise,
                          save_mode=save_mode, layout=lay, devices=devices)
    sol = EMEnsembleSolution(n_traj=n, dtype=dtype, save_mode=save_mode, layout=lay, u0_soa=u0_soa, u_raw=raw,
                             t=em_times(base.tspan, dt, dtype), prob=ens, alg=alg)
    sol.seed = seed
    if single:
        one = sol[0]
        try:
            one.seed = seed
        except AttributeError:
            pass
        return one
    return sol
