"""Shared helpers for the parity tests: seeded inputs and the CUDA-vs-oracle comparison."""
import numpy as np

ALG_NAMES = {"GPUSimpleTsit5": "Tsit5", "GPUSimpleATsit5": "ATsit5", "GPUSimpleRK4": "RK4",
             "GPUSimpleVern7": "Vern7", "GPUSimpleAVern7": "AVern7", "GPUSimpleVern9": "Vern9",
             "GPUSimpleAVern9": "AVern9", "GPUSimpleEuler": "Euler"}


def lorenz_sweep(n, dtype=np.float64, rho_max=21.0):
    """u0 = (1,0,0), p_i = (10, rho_i, 8/3), rho_i = rho_max*(i-1)/(N-1)."""
    u0 = np.zeros((n, 3), dtype=dtype)
    u0[:, 0] = 1
    p = np.empty((n, 3), dtype=dtype)
    p[:, 0] = 10
    p[:, 1] = (rho_max * np.arange(n, dtype=np.float64) / max(n - 1, 1)).astype(dtype)
    p[:, 2] = dtype(8.0) / dtype(3.0) if dtype is np.float32 else 8.0 / 3.0
    return u0, p


def vdp_sweep(n, dtype=np.float64, shuffled=False):
    """u0 = (2,0), mu_i = 0.1 + 49.9*(i-1)/(N-1)."""
    u0 = np.zeros((n, 2), dtype=dtype)
    u0[:, 0] = 2
    idx = np.arange(n, dtype=np.int64)
    if shuffled:
        idx = (idx * 2654435761) % n
    p = (0.1 + 49.9 * idx.astype(np.float64) / max(n - 1, 1)).astype(dtype).reshape(n, 1)
    return u0, p


def random_problem(system, n, dtype, seed):
    rng = np.random.default_rng(seed)
    if system == "lorenz":
        u0 = rng.uniform(-5, 5, (n, 3))
        p = np.stack([rng.uniform(5, 12, n), rng.uniform(0, 30, n), rng.uniform(1, 3, n)], 1)
    elif system == "vanderpol":
        u0 = rng.uniform(-2, 2, (n, 2))
        p = rng.uniform(0.1, 5, (n, 1))
    elif system == "robertson":
        u0 = np.stack([rng.uniform(0.5, 1, n), rng.uniform(0, 0.2, n), rng.uniform(0, 0.2, n)], 1)
        p = np.stack([rng.uniform(0.01, 0.1, n), rng.uniform(1, 30, n), rng.uniform(1, 10, n)], 1)
    elif system == "nbody":
        pos = rng.uniform(-1, 1, (n, 6)) + np.array([1.5, 0, -1.5, 0.5, 0, -1.5])
        vel = rng.uniform(-0.3, 0.3, (n, 6))
        u0 = np.concatenate([pos, vel], 1)
        p = rng.uniform(0.5, 1.5, (n, 3))
    elif system == "lineardecay":
        u0 = rng.uniform(0.5, 2, (n, 3))
        p = np.tile([10.0, 28.0, 8 / 3], (n, 1))
    elif system == "scalargrowth":
        u0 = rng.uniform(0.1, 1, (n, 1))
        p = rng.uniform(-1.5, 1.01, (n, 1))
    elif system == "nonautonomous":
        u0 = rng.uniform(-1, 1, (n, 2))
        p = rng.uniform(0.5, 2, (n, 2))
    else:
        raise KeyError(system)
    return u0.astype(dtype), p.astype(dtype)


def bits_equal(a, b):
    """bit-for-bit equality (NaN payloads included; -0 != +0)."""
    a = np.ascontiguousarray(a)
    b = np.ascontiguousarray(b)
    return a.shape == b.shape and a.dtype == b.dtype and a.tobytes() == b.tobytes()


def max_ulp_diff(a, b):
    a = np.ascontiguousarray(a)
    b = np.ascontiguousarray(b)
    it = np.int64 if a.dtype == np.float64 else np.int32
    ai, bi = a.view(it).astype(np.int64), b.view(it).astype(np.int64)
    both_nan = np.isnan(a) & np.isnan(b)
    d = np.abs(ai - bi)
    d[both_nan] = 0
    return int(d.max()) if d.size else 0
