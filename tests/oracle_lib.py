"""ctypes wrapper around oracle/liboracle.so (the CPU restatement of the reference).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs.  Never imported by the product package.
"""
import ctypes
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
LIB_PATH = os.path.join(ORACLE_DIR, "liboracle.so")

# ids of oracle.cpp (deliberately NOT imported from the product's headers)
ALG = dict(Tsit5=0, ATsit5=1, RK4=2, Vern7=3, AVern7=4, Vern9=5, AVern9=6, Euler=7)
SYS = dict(lorenz=0, vanderpol=1, robertson=2, nbody=3, lineardecay=4, scalargrowth=5,
           nonautonomous=6, user=100)
SAVE_ENDPOINT, SAVE_SAVEAT, SAVE_EVERYSTEP = 0, 1, 2
ADAPTIVE = {"ATsit5", "AVern7", "AVern9"}

_lib = None


def build(force=False):
    src = [os.path.join(ORACLE_DIR, f) for f in ("oracle.cpp", "oracle_em.cpp", "tableau_named.hpp", "Makefile")]
    if (not force and os.path.exists(LIB_PATH)
            and os.path.getmtime(LIB_PATH) >= max(os.path.getmtime(s) for s in src)):
        return LIB_PATH
    subprocess.check_call(["make", "-C", ORACLE_DIR, "-s"])
    return LIB_PATH


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = ctypes.CDLL(LIB_PATH)
        _lib.oracle_solve.restype = ctypes.c_int
        _lib.oracle_solve.argtypes = [
            ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int64, ctypes.c_void_p, ctypes.c_void_p,
            ctypes.c_double, ctypes.c_double, ctypes.c_double, ctypes.c_double, ctypes.c_double,
            ctypes.c_int64, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int64, ctypes.c_int, ctypes.c_int,
            ctypes.c_int64, ctypes.c_int64, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p,
            ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int,
            ctypes.c_void_p, ctypes.c_int, ctypes.c_int]
        _lib.oracle_system_dims.argtypes = [ctypes.c_int, ctypes.POINTER(ctypes.c_int), ctypes.POINTER(ctypes.c_int)]
    return _lib


def hardware_threads():
    return lib().oracle_hardware_threads()


def _ptr(a):
    return None if a is None else a.ctypes.data_as(ctypes.c_void_p)


class OracleResult:
    pass


def solve(system, alg, u0, p, t0, tf, dt, *, dtype=np.float64, abstol=1e-6, reltol=1e-3,
          tgrid=None, saveat=None, save_mode=SAVE_ENDPOINT, compat=0, max_out=None,
          max_attempts=0, n_threads=1, want_t=False, user_fn=None, user_dims=None):
    """u0: [n_traj, N] (AoS rows) or [N]; p likewise. Returns OracleResult with
    u [n_traj, max_out, N], t [n_traj, max_out] (if want_t), n, naccept, nreject, retcode."""
    L = lib()
    dtype = np.dtype(dtype)
    u0 = np.atleast_2d(np.asarray(u0, dtype=dtype))
    p = np.atleast_2d(np.asarray(p, dtype=dtype))
    n_traj, N = u0.shape
    if p.shape[0] == 1 and n_traj > 1:
        p = np.repeat(p, n_traj, axis=0)
    u0_soa = np.ascontiguousarray(u0.T)
    p_soa = np.ascontiguousarray(p.T)
    fixed = alg not in ADAPTIVE
    if fixed:
        assert tgrid is not None, "fixed-step algorithms need the time grid t0:dt:tf"
        tgrid = np.ascontiguousarray(tgrid, dtype=dtype)
        n_steps = len(tgrid) - 1
    else:
        n_steps = 0
    n_save = 0
    if saveat is not None:
        saveat = np.ascontiguousarray(saveat, dtype=dtype)
        n_save = len(saveat)
        save_mode = SAVE_SAVEAT
    if max_out is None:
        if save_mode == SAVE_ENDPOINT:
            max_out = 1
        elif save_mode == SAVE_SAVEAT:
            max_out = max(n_save, 1)
        else:
            assert fixed, "adaptive save_everystep needs an explicit max_out capacity"
            max_out = n_steps + 1
    out_u = np.empty((n_traj, max_out, N), dtype=dtype)
    out_t = np.full((n_traj, max_out), np.nan, dtype=dtype) if want_t else None
    out_n = np.zeros(n_traj, dtype=np.int64)
    nacc = np.zeros(n_traj, dtype=np.int32)
    nrej = np.zeros(n_traj, dtype=np.int32)
    ret = np.zeros(n_traj, dtype=np.int32)
    sysid = SYS[system] if isinstance(system, str) else system
    un, unp = (0, 0) if user_dims is None else user_dims
    rc = L.oracle_solve(sysid, ALG[alg], 0 if dtype == np.float64 else 1, n_traj, _ptr(u0_soa), _ptr(p_soa),
                        float(t0), float(tf), float(dt), float(abstol), float(reltol), n_steps,
                        _ptr(tgrid), _ptr(saveat), n_save, save_mode, compat, max_out, max_attempts,
                        _ptr(out_u), _ptr(out_t), _ptr(out_n), _ptr(nacc), _ptr(nrej), _ptr(ret),
                        n_threads, user_fn, un, unp)
    if rc != 0:
        raise RuntimeError("oracle_solve failed with %d" % rc)
    r = OracleResult()
    r.u, r.t, r.n, r.naccept, r.nreject, r.retcode = out_u, out_t, out_n, nacc, nrej, ret
    return r


# ---- SimpleEM (oracle_em.cpp) ------------------------------------------------------------------
EM_SYS = dict(gbm=0, linadd1=1, linadd2=2, ou=3, nondiag2x4=4)


def em_dims(system):
    L = lib()
    a, b, c, d = ctypes.c_int(), ctypes.c_int(), ctypes.c_int(), ctypes.c_int()
    rc = L.oracle_em_dims(EM_SYS[system], ctypes.byref(a), ctypes.byref(b), ctypes.byref(c), ctypes.byref(d))
    assert rc == 0
    return a.value, b.value, c.value, bool(d.value)


def em_solve(system, u0_soa, p_soa, t0, dt, n_steps, noise, n_threads=4):
    """u0_soa [N][n], p_soa [NP][n], noise [n_steps][M][n] -> [n][n_steps+1][N] (every state)."""
    L = lib()
    dtype = u0_soa.dtype
    N, n = u0_soa.shape
    u0c, pc, zc = (np.ascontiguousarray(x, dtype=dtype) for x in (u0_soa, p_soa, noise))
    out = np.empty((n, n_steps + 1, N), dtype=dtype)
    L.oracle_em_solve.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.c_int64, ctypes.c_void_p, ctypes.c_void_p,
                                  ctypes.c_double, ctypes.c_double, ctypes.c_int64, ctypes.c_void_p, ctypes.c_void_p,
                                  ctypes.c_int]
    rc = L.oracle_em_solve(EM_SYS[system], 0 if dtype == np.float64 else 1, n, _ptr(u0c), _ptr(pc), float(t0),
                           float(dt), int(n_steps), _ptr(zc), _ptr(out), n_threads)
    assert rc == 0
    return out


def em_normals(dtype, seed, traj_offset, n_traj, n_steps, M):
    """The normals of the CUDA path's noise specification (Philox4x32-10 + Box-Muller): [n_steps][M][n]."""
    L = lib()
    out = np.empty((n_steps, M, n_traj), dtype=dtype)
    L.oracle_em_normals.argtypes = [ctypes.c_int, ctypes.c_uint64, ctypes.c_int64, ctypes.c_int64, ctypes.c_int64,
                                    ctypes.c_int, ctypes.c_void_p]
    L.oracle_em_normals(0 if np.dtype(dtype) == np.float64 else 1, seed, traj_offset, n_traj, n_steps, M, _ptr(out))
    return out


def philox4x32_10(ctr, key):
    L = lib()
    c = (ctypes.c_uint32 * 4)(*ctr)
    k = (ctypes.c_uint32 * 2)(*key)
    o = (ctypes.c_uint32 * 4)()
    L.oracle_philox4x32_10.restype = None
    L.oracle_philox4x32_10(c, k, o)
    return [int(x) for x in o]
