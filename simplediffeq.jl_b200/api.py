"""Host-side mirror of the reference's interface for the GPUSimple* path.

Mirrors (names, argument meaning, defaults, error behaviour) of SciML/SimpleDiffEq.jl v1.16.3:

    solve(prob::ODEProblem, GPUSimpleTsit5();  saveat=nothing, save_everystep=true, dt=0.1f0)
        src/tsit5/gpuatsit5.jl:55-61
    solve(prob::ODEProblem, GPUSimpleATsit5(); dt=0.1f0, saveat=nothing, save_everystep=true,
          abstol=1f-6, reltol=1f-3)                       src/tsit5/gpuatsit5.jl:205-212
    solve(prob::ODEProblem, GPUSimpleRK4(); dt=<required>)  src/rk4/gpurk4.jl:53-58
    GPUSimpleVern7 / AVern7 / Vern9 / AVern9: same keywords  src/verner/gpuvern7.jl:55-61,300-306,
                                                             src/verner/gpuvern9.jl:55-61,411-417
    solve(EnsembleProblem(prob; prob_func), alg; trajectories, kw...)   [SciMLBase ensemble driver]

The whole ensemble crosses the C ABI (include/simplediffeq_cuda.h) in ONE call.  `prob.f` must be
a built-in system (`systems.lorenz`, ...) or a `CudaRHS` (CUDA-C source compiled by NVRTC): an
arbitrary host callable cannot run on the GPU, and there is no CPU fallback.
"""
import ctypes

import numpy as np

from . import _lib
from .jlrange import JuliaRange


# --------------------------------------------------------------------------------------------
# right-hand sides
# --------------------------------------------------------------------------------------------
class System:
    """A right-hand side f(u, p, t) available on the device."""

    def __init__(self, handle, name, n_state, n_param, owned):
        self._handle, self.name, self.n_state, self.n_param, self._owned = handle, name, n_state, n_param, owned

    def __repr__(self):
        return "System(%s, n_state=%d, n_param=%d)" % (self.name, self.n_state, self.n_param)

    def __del__(self):
        try:
            if self._owned and self._handle:
                _lib.lib().sde_system_free(self._handle)
        except Exception:
            pass


def builtin_system(name):
    h = ctypes.c_void_p()
    _lib.check(_lib.lib().sde_system_builtin(name.encode(), ctypes.byref(h)))
    ns, npar = ctypes.c_int(), ctypes.c_int()
    _lib.check(_lib.lib().sde_system_dims(h, ctypes.byref(ns), ctypes.byref(npar)))
    return System(h, name, ns.value, npar.value, False)


class _Systems:
    """Built-in registry: systems.lorenz, .vanderpol, .robertson, .nbody, .lineardecay,
    .scalargrowth, .nonautonomous (resolved lazily so importing the package needs no library)."""
    _names = ("lorenz", "vanderpol", "robertson", "nbody", "lineardecay", "scalargrowth", "nonautonomous")

    def __init__(self):
        self._cache = {}

    def __getattr__(self, name):
        if name.startswith("_") or name not in self._names:
            raise AttributeError(name)
        if name not in self._cache:
            self._cache[name] = builtin_system(name)
        return self._cache[name]

    def names(self):
        return self._names


systems = _Systems()


def CudaRHS(src, n_state, n_param):
    """User right-hand side as CUDA C++ source defining
    `__device__ void rhs(real* du, const real* u, const real* p, real t)`; JIT-compiled with NVRTC
    for sm_100a, --fmad=false (the reference does not fuse the user's f)."""
    h = ctypes.c_void_p()
    log = ctypes.create_string_buffer(16384)
    rc = _lib.lib().sde_system_nvrtc(src.encode(), n_state, n_param, ctypes.byref(h), log, len(log))
    if rc != _lib.SDE_OK:
        raise _lib.SdeError(rc, _lib.lib().sde_last_error().decode("utf-8", "replace"))
    return System(h, "user", n_state, n_param, True)


# --------------------------------------------------------------------------------------------
# algorithms (the reference's singletons)
# --------------------------------------------------------------------------------------------
class _Alg:
    adaptive = False

    def __repr__(self):
        return type(self).__name__ + "()"

    @property
    def alg_id(self):
        return _lib.ALG_IDS[type(self).__name__]


class GPUSimpleTsit5(_Alg):
    """Fixed-step Tsitouras 5(4) (src/tsit5/gpuatsit5.jl:52)."""


class GPUSimpleATsit5(_Alg):
    """Adaptive Tsitouras 5(4) with PI step control (src/tsit5/gpuatsit5.jl:200)."""
    adaptive = True


class GPUSimpleRK4(_Alg):
    """Classic fixed-step RK4; always saves every step (src/rk4/gpurk4.jl:50)."""


class GPUSimpleEuler(_Alg):
    """Forward Euler; always saves every step (src/euler/gpueuler.jl:50)."""


class GPUSimpleVern7(_Alg):
    """Fixed-step Verner 7(6) (src/verner/gpuvern7.jl:52)."""


class GPUSimpleAVern7(_Alg):
    """Adaptive Verner 7(6) (src/verner/gpuvern7.jl:295)."""
    adaptive = True


class GPUSimpleVern9(_Alg):
    """Fixed-step Verner 9(8) (src/verner/gpuvern9.jl:52)."""


class GPUSimpleAVern9(_Alg):
    """Adaptive Verner 9(8) (src/verner/gpuvern9.jl:406)."""
    adaptive = True


# --------------------------------------------------------------------------------------------
# problems
# --------------------------------------------------------------------------------------------
class ODEProblem:
    """ODEProblem{false}(f, u0, tspan, p): out-of-place problem.  eltype(u0) selects Float64/Float32."""

    def __init__(self, f, u0, tspan, p=None, analytic=None):
        """analytic: optional host callable (u0, p, t) -> u, the counterpart of ODEFunction(f; analytic = ...): when
        present every solution carries `u_analytic` and `errors` like the reference's epilogue
        `has_analytic(prob.f) && calculate_solution_errors!(sol; timeseries_errors = true, dense_errors = false)`
        (src/tsit5/gpuatsit5.jl:141-145, :330-334, src/rk4/gpurk4.jl:92-96, ...)."""
        self.analytic = analytic
        if not isinstance(f, System):
            raise TypeError("f must be a built-in system or a CudaRHS (no host callables on the GPU path)")
        u0 = np.atleast_1d(np.asarray(u0))
        self.dtype = np.dtype(np.float32) if u0.dtype == np.float32 else np.dtype(np.float64)
        self.f = f
        self.u0 = u0.astype(self.dtype)
        self.tspan = (tspan[0], tspan[1])
        self.p = np.zeros(0, self.dtype) if p is None else np.atleast_1d(np.asarray(p, dtype=self.dtype))
        if self.u0.shape != (f.n_state,):
            raise ValueError("u0 must have %d components" % f.n_state)
        if self.p.shape != (f.n_param,):
            raise ValueError("p must have %d components" % f.n_param)


def remake(prob, u0=None, p=None, tspan=None):
    return ODEProblem(prob.f, prob.u0 if u0 is None else u0, prob.tspan if tspan is None else tspan,
                      prob.p if p is None else p, analytic=prob.analytic)


class EnsembleProblem:
    """EnsembleProblem(prob; prob_func).  prob_func(prob, i, repeat) -> ODEProblem is evaluated on the
    host for i = 1..trajectories (1-based like the reference).  For large ensembles pass the
    batched form instead: u0s [trajectories, n_state] and/or ps [trajectories, n_param]."""

    def __init__(self, prob, prob_func=None, u0s=None, ps=None):
        self.prob, self.prob_func, self.u0s, self.ps = prob, prob_func, u0s, ps


# --------------------------------------------------------------------------------------------
# solutions
# --------------------------------------------------------------------------------------------
RETCODES = {0: "Default", 1: "DtLessThanMin", 2: "MaxIters"}


class ODESolution:
    """One trajectory: .u [n_saved, n_state], .t [n_saved], .retcode, .naccept, .nreject."""

    def __init__(self, ens, i):
        self._ens, self._i = ens, i

    @property
    def u(self):
        return self._ens._u_of(self._i)

    @property
    def t(self):
        return self._ens._t_of(self._i)

    @property
    def retcode(self):
        return RETCODES.get(int(self._ens.retcode[self._i]), "Failure")

    @property
    def naccept(self):
        return int(self._ens.naccept[self._i])

    @property
    def nreject(self):
        return int(self._ens.nreject[self._i])

    def __len__(self):
        return len(self.t)

    # ---- [EXT] SciMLBase.calculate_solution_errors!(sol; timeseries_errors = true, dense_errors = false), evaluated on
    # the host on demand: u_analytic[i] = analytic(u0, p, t[i]); errors = (final, l-infinity, l2) of u - u_analytic
    @property
    def u_analytic(self):
        an = self._ens.prob.prob.analytic
        if an is None:
            return None
        u0 = self._ens.u0_soa[:, self._i]
        p = self._ens.p_soa[:, self._i] if getattr(self._ens, "p_soa", None) is not None else self._ens.prob.prob.p
        return np.stack([np.atleast_1d(np.asarray(an(u0, p, tk), dtype=self._ens.dtype)) for tk in self.t])

    @property
    def errors(self):
        ua = self.u_analytic
        if ua is None:
            return None
        d = np.asarray(self.u, dtype=np.float64) - ua.astype(np.float64)
        return {"final": float(np.mean(np.abs(d[-1]))), "l∞": float(np.max(np.abs(d))), "l2": float(np.sqrt(np.mean(d * d)))}


class EnsembleSolution:
    """Result of an ensemble solve.  Indexing gives per-trajectory ODESolution views; the raw
    arrays are `.u_raw` (layout-dependent), `.t_shared` / `.t_final`, `.naccept`, `.nreject`,
    `.retcode`."""

    def __init__(self, **kw):
        self.__dict__.update(kw)
        self.converged = bool(np.all(self.retcode == 0))

    def __len__(self):
        return self.n_traj

    def __getitem__(self, i):
        if i < 0:
            i += self.n_traj
        if not 0 <= i < self.n_traj:
            raise IndexError(i)
        return ODESolution(self, i)

    def __iter__(self):
        return (ODESolution(self, i) for i in range(self.n_traj))

    def _series(self, i):
        u = self.u_raw[i] if self.layout == _lib.LAYOUT_TRAJ_MAJOR else self.u_raw[:, :, i]
        if self.t_series is not None:      # adaptive save_everystep: naccept + 1 states were saved
            u = u[:min(int(self.naccept[i]) + 1, u.shape[0])]
        return u

    def _u_of(self, i):
        if self.save_mode == _lib.SAVE_ENDPOINT:
            return np.stack([self.u0_soa[:, i], self.u_raw[:, i]])   # us = [u0, u_end]
        return self._series(i)

    def _t_of(self, i):
        if self.save_mode == _lib.SAVE_ENDPOINT and self.t_final is not None:
            return np.array([self.t0, self.t_final[i]], dtype=self.dtype)
        if self.t_series is not None:
            t = self.t_series[i] if self.layout == _lib.LAYOUT_TRAJ_MAJOR else self.t_series[:, i]
            return t[:min(int(self.naccept[i]) + 1, t.shape[0])]
        return self.t_shared


# --------------------------------------------------------------------------------------------
# solve
# --------------------------------------------------------------------------------------------
def _as_T(x, dtype):
    return float(np.asarray(x).astype(dtype))


def make_options(alg, dtype, n_traj, tspan, dt, abstol, reltol, saveat, save_mode, layout, compat,
                 max_attempts, keep, out_capacity=0):
    """Fill an SdeOptions.  `keep` collects the numpy arrays the struct points into."""
    o = _lib.SdeOptions()
    o.alg = alg.alg_id
    o.dtype = _lib.SDE_F64 if dtype == np.float64 else _lib.SDE_F32
    o.save_mode, o.layout, o.compat = save_mode, layout, compat
    o.n_traj = n_traj
    o.t0, o.tf = _as_T(tspan[0], dtype), _as_T(tspan[1], dtype)
    o.dt = _as_T(dt, dtype)
    o.abstol, o.reltol = _as_T(abstol, dtype), _as_T(reltol, dtype)
    o.max_attempts = max_attempts
    o.out_capacity = out_capacity
    if not alg.adaptive:
        grid = JuliaRange(o.t0, o.dt, o.tf, dtype).collect()   # _ts = tspan[1]:dt:tspan[2]
        if len(grid) == 0:
            raise ValueError("empty time range")
        keep.append(grid)
        o.n_steps = len(grid) - 1
        o.tgrid = grid.ctypes.data
    if save_mode == _lib.SAVE_SAVEAT:
        sa = np.ascontiguousarray(np.asarray(saveat).astype(dtype))
        keep.append(sa)
        o.saveat = sa.ctypes.data
        o.n_save = len(sa)
    return o


def _save_mode(alg, saveat, save_everystep):
    if isinstance(alg, (GPUSimpleRK4, GPUSimpleEuler)):
        return _lib.SAVE_EVERYSTEP       # the reference's RK4 / Euler swallow saveat / save_everystep
    if saveat is not None:
        return _lib.SAVE_SAVEAT
    return _lib.SAVE_EVERYSTEP if save_everystep else _lib.SAVE_ENDPOINT


def fixed_times(o, dtype):
    n = max(int(o.n_save), int(o.n_steps) + 1, 2)
    out = np.empty(n, dtype=dtype)
    nw = ctypes.c_int64()
    _lib.check(_lib.lib().sde_fixed_times(ctypes.byref(o), out.ctypes.data, n, ctypes.byref(nw)))
    return out[:nw.value].copy()


# adaptive save_everystep = true on ensembles larger than EVERYSTEP_PILOT: rows are sized from a pilot sample of
# that many trajectories, times EVERYSTEP_SLACK (see solve)
EVERYSTEP_PILOT = 4096
EVERYSTEP_SLACK = 1.25


def solve(prob, alg, ensemblealg=None, *, trajectories=None, dt=None, abstol=None, reltol=None,
          saveat=None, save_everystep=True, devices=None, layout="traj_major", compat=0,
          maxiters=0, **kwargs):
    """solve(prob, alg; kw...) for an ODEProblem (one trajectory) or an EnsembleProblem.

    Unknown keywords are swallowed like the reference's `kwargs...`.  `devices` (list of CUDA
    ordinals) shards the ensemble by contiguous index ranges; `layout` ("traj_major" | "soa")
    selects the memory layout of series outputs; `maxiters` (0 = unlimited like the reference)
    bounds adaptive attempts per trajectory."""
    from . import em as _em
    if isinstance(alg, _em.SimpleEM):     # SDEProblem + SimpleEM: src/euler_maruyama.jl:48-94
        return _em.solve_em(prob, alg, dt=dt, trajectories=trajectories, save_everystep=save_everystep,
                            devices=devices, layout=layout, **kwargs)
    if isinstance(prob, ODEProblem):
        ens = EnsembleProblem(prob)
        trajectories = 1
        single = True
    else:
        ens, single = prob, False
        if trajectories is None:
            raise TypeError("trajectories is required for an EnsembleProblem")
    base = ens.prob
    dtype, sysm, n = base.dtype, base.f, int(trajectories)

    # reference defaults (Float32 literals, converted to eltype(u0) where they meet the state)
    if dt is None:
        if isinstance(alg, (GPUSimpleRK4, GPUSimpleEuler)):
            raise ValueError("dt is required for this algorithm")   # src/rk4/gpurk4.jl:56, src/euler/gpueuler.jl:56
        dt = np.float32(0.1)
    abstol = np.float32(1e-6) if abstol is None else abstol
    reltol = np.float32(1e-3) if reltol is None else reltol
    if isinstance(saveat, JuliaRange):
        saveat = saveat.collect()

    # ---- collect per-trajectory u0 / p (SoA) ------------------------------------------------
    u0_soa = np.empty((sysm.n_state, n), dtype=dtype)
    p_soa = np.empty((sysm.n_param, n), dtype=dtype)
    u0_soa[:] = base.u0[:, None]
    p_soa[:] = base.p[:, None]
    if ens.u0s is not None:
        u0_soa[:] = np.asarray(ens.u0s, dtype=dtype).reshape(n, sysm.n_state).T
    if ens.ps is not None:
        p_soa[:] = np.asarray(ens.ps, dtype=dtype).reshape(n, sysm.n_param).T
    if ens.prob_func is not None:
        for i in range(n):
            pi = ens.prob_func(base, i + 1, 1)
            if pi.f is not base.f or tuple(pi.tspan) != tuple(base.tspan) or pi.dtype != dtype:
                raise ValueError("prob_func may only change u0 and p on the GPU ensemble path")
            u0_soa[:, i] = pi.u0
            p_soa[:, i] = pi.p

    save_mode = _save_mode(alg, saveat, save_everystep)
    lay = _lib.LAYOUT_TRAJ_MAJOR if layout == "traj_major" else _lib.LAYOUT_SOA
    common = dict(dt=dt, abstol=abstol, reltol=reltol, compat=compat, maxiters=maxiters, devices=devices)
    capacity = 0
    if alg.adaptive and save_mode == _lib.SAVE_EVERYSTEP:
        # variable-length output (the reference push!es every accepted step, gpuatsit5.jl:301-303).  The step
        # sequence is deterministic, so an endpoint-only pass gives the sizes: over the whole ensemble when it is
        # small, else over a pilot sample of evenly spaced trajectories (the full solve then reports the
        # trajectories whose row was too short -- SDE_RET_OUTPUT_FULL with the needed count in naccept -- and
        # only in that case the ensemble is solved once more with the exact capacity).
        if n <= EVERYSTEP_PILOT:
            sel = slice(None)
        else:
            sel = np.unique(np.linspace(0, n - 1, EVERYSTEP_PILOT).astype(np.int64))
        first = solve_arrays(sysm, alg, np.ascontiguousarray(u0_soa[:, sel]), np.ascontiguousarray(p_soa[:, sel]),
                             base.tspan, save_mode=_lib.SAVE_ENDPOINT, **common)
        if np.any(first["retcode"] == _lib.RET_DTMIN):
            raise RuntimeError("dt<dtmin")
        capacity = int(first["naccept"].max()) + 1
        if n > EVERYSTEP_PILOT:
            capacity = int(capacity * EVERYSTEP_SLACK) + 8
    raw = solve_arrays(sysm, alg, u0_soa, p_soa, base.tspan, saveat=saveat, save_mode=save_mode, layout=lay,
                       out_capacity=capacity, **common)
    if np.any(raw["retcode"] == _lib.RET_DTMIN):
        raise RuntimeError("dt<dtmin")       # the reference throws (src/tsit5/gpuatsit5.jl:256)
    if capacity and np.any(raw["retcode"] == _lib.RET_OUTPUT_FULL):
        capacity = int(raw["naccept"].max()) + 1
        raw = solve_arrays(sysm, alg, u0_soa, p_soa, base.tspan, saveat=saveat, save_mode=save_mode, layout=lay,
                           out_capacity=capacity, **common)
    sol = EnsembleSolution(n_traj=n, dtype=dtype, save_mode=save_mode, layout=lay, u0_soa=u0_soa, p_soa=p_soa,
                           u_raw=raw["u"], t_shared=raw["t_shared"], t_final=raw["t_final"],
                           t_series=raw["t_series"],
                           t0=_as_T(base.tspan[0], dtype), naccept=raw["naccept"],
                           nreject=raw["nreject"], retcode=raw["retcode"], prob=ens, alg=alg)
    return sol[0] if single else sol


def solve_arrays(system, alg, u0_soa, p_soa, tspan, *, dt, abstol=1e-6, reltol=1e-3, saveat=None,
                 save_mode=_lib.SAVE_ENDPOINT, layout=_lib.LAYOUT_TRAJ_MAJOR, compat=0, maxiters=0,
                 devices=None, out_capacity=0):
    """Array-level entry: SoA host arrays in, raw arrays out, ONE sde_solve call (H2D, kernel(s), D2H).
    Adaptive SAVE_EVERYSTEP needs out_capacity (slots per trajectory)."""
    dtype = u0_soa.dtype
    n = u0_soa.shape[1]
    keep = []
    o = make_options(alg, dtype, n, tspan, dt, abstol, reltol, saveat, save_mode, layout, compat,
                     maxiters, keep, out_capacity)
    N = system.n_state
    t_series = None
    if save_mode == _lib.SAVE_ENDPOINT:
        out_u = np.empty((N, n), dtype=dtype)
    else:
        slots = (int(o.n_save) if save_mode == _lib.SAVE_SAVEAT
                 else int(out_capacity) if alg.adaptive else int(o.n_steps) + 1)
        shape = (n, slots, N) if layout == _lib.LAYOUT_TRAJ_MAJOR else (slots, N, n)
        out_u = np.empty(shape, dtype=dtype)
        if alg.adaptive and save_mode == _lib.SAVE_EVERYSTEP:
            t_series = np.empty((n, slots) if layout == _lib.LAYOUT_TRAJ_MAJOR else (slots, n), dtype=dtype)
    t_final = np.empty(n, dtype=dtype) if (alg.adaptive and t_series is None) else None
    nacc = np.zeros(n, dtype=np.int32)
    nrej = np.zeros(n, dtype=np.int32)
    ret = np.zeros(n, dtype=np.int32)
    u0c, pc = np.ascontiguousarray(u0_soa), np.ascontiguousarray(p_soa)
    dev = None
    ndev = 0
    if devices:
        dev = (ctypes.c_int * len(devices))(*devices)
        ndev = len(devices)
    rc = _lib.lib().sde_solve(system._handle, ctypes.byref(o), u0c.ctypes.data, pc.ctypes.data if pc.size else None,
                              out_u.ctypes.data,
                              t_series.ctypes.data if t_series is not None else (t_final.ctypes.data if t_final is not None else None),
                              nacc.ctypes.data, nrej.ctypes.data, ret.ctypes.data, dev, ndev)
    _lib.check(rc)
    if alg.adaptive:
        t_shared = keep[-1] if save_mode == _lib.SAVE_SAVEAT else None
    else:
        t_shared = fixed_times(o, dtype)
        nacc[:] = o.n_steps
    return dict(u=out_u, t_shared=t_shared, t_final=t_final, t_series=t_series, naccept=nacc, nreject=nrej,
                retcode=ret, n_steps=int(o.n_steps))


def solve_device(system, alg, d_u0, d_p, tspan, *, dt, abstol=1e-6, reltol=1e-3, saveat=None,
                 save_mode=_lib.SAVE_ENDPOINT, layout=_lib.LAYOUT_TRAJ_MAJOR, compat=0, maxiters=0,
                 out=None, stats=True, stream=None, sync=True, out_capacity=0):
    """Device-resident entry: d_u0 [n_state, n] and d_p [n_param, n] are torch CUDA tensors (SoA);
    returns torch tensors on the same device.  torch only provides memory and the stream.
    Adaptive algorithms with save_mode = SAVE_EVERYSTEP need `out_capacity` slots per trajectory (rows that are too
    short come back with retcode OUTPUT_FULL and the needed count in naccept, like the C ABI reports them); the time of
    every slot is returned as `t_series`."""
    import torch
    assert d_u0.is_cuda and d_u0.is_contiguous()
    dtype = np.dtype(np.float64) if d_u0.dtype == torch.float64 else np.dtype(np.float32)
    N, n = d_u0.shape
    keep = []
    every = alg.adaptive and save_mode == _lib.SAVE_EVERYSTEP
    if every and out_capacity < 1:
        raise ValueError("adaptive save_everystep needs out_capacity >= 1 slots per trajectory")
    o = make_options(alg, dtype, n, tspan, dt, abstol, reltol, saveat, save_mode, layout, compat,
                     maxiters, keep, out_capacity=int(out_capacity) if every else 0)
    slots = 1
    if save_mode != _lib.SAVE_ENDPOINT:
        slots = int(o.n_save) if save_mode == _lib.SAVE_SAVEAT else (int(out_capacity) if every else int(o.n_steps) + 1)
    if out is None:
        if save_mode == _lib.SAVE_ENDPOINT:
            shape = (N, n)
        else:
            shape = (n, slots, N) if layout == _lib.LAYOUT_TRAJ_MAJOR else (slots, N, n)
        out = torch.empty(shape, dtype=d_u0.dtype, device=d_u0.device)
    t_final = nacc = nrej = ret = t_series = None
    if every:       # the C ABI's out_t holds the time of every slot in this mode
        t_series = torch.empty((n, slots) if layout == _lib.LAYOUT_TRAJ_MAJOR else (slots, n), dtype=d_u0.dtype, device=d_u0.device)
    elif alg.adaptive:
        t_final = torch.empty(n, dtype=d_u0.dtype, device=d_u0.device)
    if stats and alg.adaptive:
        nacc = torch.zeros(n, dtype=torch.int32, device=d_u0.device)
        nrej = torch.zeros(n, dtype=torch.int32, device=d_u0.device)
    if stats:
        ret = torch.zeros(n, dtype=torch.int32, device=d_u0.device)
    st = torch.cuda.current_stream(d_u0.device).cuda_stream if stream is None else stream
    ptr = lambda x: None if x is None else x.data_ptr()
    with torch.cuda.device(d_u0.device):
        rc = _lib.lib().sde_solve_device(system._handle, ctypes.byref(o), d_u0.data_ptr(),
                                         ptr(d_p) if d_p is not None and d_p.numel() else None, n,
                                         out.data_ptr(), n, ptr(t_series if every else t_final), ptr(nacc), ptr(nrej), ptr(ret),
                                         st, 0 if sync else 1)
    _lib.check(rc)
    return dict(u=out, t_final=t_final, t_series=t_series, naccept=nacc, nreject=nrej, retcode=ret, n_steps=int(o.n_steps),
                options=o, keep=keep)
