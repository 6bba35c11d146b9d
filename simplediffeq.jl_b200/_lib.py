"""ctypes binding of libsimplediffeq_cuda.so (include/simplediffeq_cuda.h).

There is NO CPU fallback: if the shared library is missing this module raises at import of the
first symbol, and every solve fails loudly when no CUDA device is present.
"""
import ctypes
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libsimplediffeq_cuda.so")

# ids of include/simplediffeq_cuda.h
SDE_OK = 0
ALG_IDS = dict(GPUSimpleTsit5=0, GPUSimpleATsit5=1, GPUSimpleRK4=2, GPUSimpleVern7=3,
               GPUSimpleAVern7=4, GPUSimpleVern9=5, GPUSimpleAVern9=6, GPUSimpleEuler=7)
SDE_F64, SDE_F32 = 0, 1
SAVE_ENDPOINT, SAVE_SAVEAT, SAVE_EVERYSTEP = 0, 1, 2
LAYOUT_TRAJ_MAJOR, LAYOUT_SOA = 0, 1
RET_DEFAULT, RET_DTMIN, RET_MAXITERS, RET_OUTPUT_FULL = 0, 1, 2, 3
COMPAT_FIX_VERN9_INTERP = 1
COMPAT_STRICT_CONTROLLER = 2
COMPAT_LOG2_CONTROLLER = 4
COMPAT_FAST_RHS = 8
COMPAT_FAST_STAGES = 16

EXPORTS = ["sde_version", "sde_last_error", "sde_device_count", "sde_system_builtin",
           "sde_system_nvrtc", "sde_system_dims", "sde_system_free", "sde_system_prepare",
           "sde_solve", "sde_solve_device", "sde_fixed_times", "sde_host_alloc", "sde_host_free",
           "sde_launch_count", "sde_probe_fma_peak", "sde_trim",
           "sde_em_system_builtin", "sde_em_system_nvrtc", "sde_em_system_dims", "sde_em_system_free",
           "sde_em_system_prepare", "sde_em_solve", "sde_em_solve_device", "sde_em_noise", "sde_em_noise_device"]


class SdeOptions(ctypes.Structure):
    _fields_ = [("alg", ctypes.c_int32), ("dtype", ctypes.c_int32), ("save_mode", ctypes.c_int32),
                ("layout", ctypes.c_int32), ("compat", ctypes.c_int32), ("reserved", ctypes.c_int32),
                ("n_traj", ctypes.c_int64), ("t0", ctypes.c_double), ("tf", ctypes.c_double),
                ("dt", ctypes.c_double), ("abstol", ctypes.c_double), ("reltol", ctypes.c_double),
                ("n_steps", ctypes.c_int64), ("tgrid", ctypes.c_void_p), ("saveat", ctypes.c_void_p),
                ("n_save", ctypes.c_int64), ("max_attempts", ctypes.c_int64), ("out_capacity", ctypes.c_int64)]


class SdeEmOptions(ctypes.Structure):
    _fields_ = [("dtype", ctypes.c_int32), ("save_mode", ctypes.c_int32), ("layout", ctypes.c_int32),
                ("noise_mode", ctypes.c_int32), ("n_traj", ctypes.c_int64), ("t0", ctypes.c_double),
                ("dt", ctypes.c_double), ("n_steps", ctypes.c_int64), ("seed", ctypes.c_uint64),
                ("traj_offset", ctypes.c_int64)]


NOISE_PHILOX, NOISE_PROVIDED = 0, 1


class SdeError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("libsimplediffeq_cuda error %d: %s" % (code, msg))
        self.code = code


_lib = None


def lib():
    """Load the native library (built in-tree by build.py). Raises if it is missing."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            "%s not found: build it with `python simplediffeq.jl_b200/build.py` "
            "(there is no CPU fallback)" % LIB_PATH)
    L = ctypes.CDLL(LIB_PATH)
    vp, i32p = ctypes.c_void_p, ctypes.POINTER(ctypes.c_int32)
    L.sde_version.restype = ctypes.c_int
    L.sde_last_error.restype = ctypes.c_char_p
    L.sde_device_count.argtypes = [ctypes.POINTER(ctypes.c_int)]
    L.sde_system_builtin.argtypes = [ctypes.c_char_p, ctypes.POINTER(vp)]
    L.sde_system_nvrtc.argtypes = [ctypes.c_char_p, ctypes.c_int, ctypes.c_int, ctypes.POINTER(vp),
                                   ctypes.c_char_p, ctypes.c_size_t]
    L.sde_system_dims.argtypes = [vp, ctypes.POINTER(ctypes.c_int), ctypes.POINTER(ctypes.c_int)]
    L.sde_system_free.argtypes = [vp]
    L.sde_system_free.restype = None
    L.sde_system_prepare.argtypes = [vp, ctypes.POINTER(SdeOptions)]
    L.sde_solve.argtypes = [vp, ctypes.POINTER(SdeOptions), vp, vp, vp, vp, vp, vp, vp,
                            ctypes.POINTER(ctypes.c_int), ctypes.c_int]
    L.sde_solve_device.argtypes = [vp, ctypes.POINTER(SdeOptions), vp, vp, ctypes.c_int64, vp,
                                   ctypes.c_int64, vp, vp, vp, vp, vp, ctypes.c_int]
    L.sde_fixed_times.argtypes = [ctypes.POINTER(SdeOptions), vp, ctypes.c_int64,
                                  ctypes.POINTER(ctypes.c_int64)]
    emo = ctypes.POINTER(SdeEmOptions)
    L.sde_em_system_builtin.argtypes = [ctypes.c_char_p, ctypes.POINTER(vp)]
    L.sde_em_system_nvrtc.argtypes = [ctypes.c_char_p, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                      ctypes.POINTER(vp), ctypes.c_char_p, ctypes.c_size_t]
    L.sde_em_system_dims.argtypes = [vp] + [ctypes.POINTER(ctypes.c_int)] * 4
    L.sde_em_system_free.argtypes = [vp]
    L.sde_em_system_free.restype = None
    L.sde_em_system_prepare.argtypes = [vp, emo]
    L.sde_em_solve.argtypes = [vp, emo, vp, vp, vp, vp, ctypes.POINTER(ctypes.c_int), ctypes.c_int]
    L.sde_em_solve_device.argtypes = [vp, emo, vp, vp, ctypes.c_int64, vp, ctypes.c_int64, vp, ctypes.c_int64,
                                      vp, ctypes.c_int]
    L.sde_em_noise.argtypes = [emo, ctypes.c_int, vp]
    L.sde_em_noise_device.argtypes = [emo, ctypes.c_int, vp, ctypes.c_int64, vp, ctypes.c_int]
    L.sde_host_alloc.argtypes = [ctypes.POINTER(vp), ctypes.c_size_t]
    L.sde_host_free.argtypes = [vp]
    L.sde_launch_count.restype = ctypes.c_int64
    L.sde_probe_fma_peak.argtypes = [ctypes.c_int, ctypes.POINTER(ctypes.c_double), ctypes.POINTER(ctypes.c_double)]
    for name in EXPORTS:
        if name not in ("sde_last_error", "sde_system_free", "sde_em_system_free", "sde_launch_count", "sde_version"):
            getattr(L, name).restype = ctypes.c_int
    _lib = L
    return L


def check(rc):
    if rc != SDE_OK:
        raise SdeError(rc, lib().sde_last_error().decode("utf-8", "replace"))


def device_count():
    n = ctypes.c_int(0)
    rc = lib().sde_device_count(ctypes.byref(n))
    return n.value if rc == SDE_OK else 0


def launch_count():
    return int(lib().sde_launch_count())


def probe_fma_peak(dtype_id=SDE_F64):
    """(TFLOP/s, ms) of the dense FMA microbenchmark on the current device."""
    tf, ms = ctypes.c_double(), ctypes.c_double()
    check(lib().sde_probe_fma_peak(dtype_id, ctypes.byref(tf), ctypes.byref(ms)))
    return tf.value, ms.value
