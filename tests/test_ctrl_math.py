"""CPU checks of the step-size controller's device math (csrc/device/sde_common.cuh), compiled for the
host through tests/ctrl_host_emul.cpp: the table-driven log2 / exp2 against mpmath, the generated table
against its generator, and the shortened NaN-propagating min/max against Julia's Base.min/max on the
domain their invariants allow."""
import ctypes
import math
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def emul(tmp_path_factory):
    out = str(tmp_path_factory.mktemp("emul") / "libctrl_emul.so")
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-mfma", "-ffp-contract=off", "-shared", "-fPIC",
                           os.path.join(ROOT, "tests", "ctrl_host_emul.cpp"), "-o", out])
    L = ctypes.CDLL(out)
    for f in ("emul_max_abs_nan2", "emul_min_abs_nan1", "emul_jl_max", "emul_jl_min"):
        getattr(L, f).restype = ctypes.c_double
        getattr(L, f).argtypes = [ctypes.c_double, ctypes.c_double]
    L.emul_ctrl.restype = ctypes.c_double
    return L


def _apply(fn, x):
    x = np.ascontiguousarray(x, dtype=np.float64)
    out = np.empty_like(x)
    fn(x.ctypes.data_as(ctypes.c_void_p), out.ctypes.data_as(ctypes.c_void_p), ctypes.c_long(len(x)))
    return out


def test_generated_table_is_current():
    """sde_ctrl_tables_gen.cuh is what tools/gen_ctrl_tables.py writes (no hand edits)."""
    path = os.path.join(ROOT, "simplediffeq.jl_b200", "csrc", "device", "sde_ctrl_tables_gen.cuh")
    before = open(path).read()
    subprocess.check_call([sys.executable, os.path.join(ROOT, "tools", "gen_ctrl_tables.py")],
                          stdout=subprocess.DEVNULL)
    assert open(path).read() == before


def test_log2_fast_accuracy(emul):
    import mpmath as mp
    mp.mp.dps = 40
    rng = np.random.default_rng(20261017)
    # squared error norms seen by the controller: 1e-30 .. 1e10, dense around 1, exact powers of two
    x = np.concatenate([10.0 ** rng.uniform(-30, 10, 1500), rng.uniform(0.5, 2.0, 1500),
                        1 + rng.uniform(-1e-3, 1e-3, 500), 2.0 ** np.arange(-40, 40),
                        np.nextafter(1.0, 0.0) * np.ones(1), np.nextafter(1.0, 2.0) * np.ones(1)])
    got = _apply(emul.emul_log2, x)
    ref = [mp.log(mp.mpf(float(v)), 2) for v in x]
    err = np.array([float(abs(mp.mpf(float(g)) - r)) for g, r in zip(got, ref)])
    # absolute error: <= 1.2e-16 + one rounding of the result
    bound = 1.2e-16 + np.spacing(np.abs(got))
    assert np.all(err <= bound), (err.max(), x[np.argmax(err - bound)])
    assert err[(x > 0.5) & (x < 2)].max() < 2.3e-16
    # the controller relies on log2(0) being hugely negative (the reference's `EEst == 0` branch)
    z = _apply(emul.emul_log2, np.array([0.0, np.inf]))
    assert z[0] <= -1000 and z[1] >= 1000


def test_exp2_fast_accuracy(emul):
    import mpmath as mp
    mp.mp.dps = 40
    rng = np.random.default_rng(7)
    y = np.concatenate([rng.uniform(-3.4, 3.4, 4000), np.arange(-3, 4), np.arange(-217, 218) / 64.0,
                        [1 / 128, -1 / 128, 3.3999, -3.3999, 1e-300, -1e-300]])
    got = _apply(emul.emul_exp2, y)
    rel = np.array([float(abs(mp.mpf(float(g)) / (mp.mpf(2) ** mp.mpf(float(v))) - 1)) for g, v in zip(got, y)])
    assert rel.max() < 2.5e-16, rel.max()
    assert _apply(emul.emul_exp2, np.array([0.0, 1.0, -2.0])).tolist() == [1.0, 2.0, 0.25]


def test_sincos_halfpi_accuracy(emul):
    """sin/cos((pi/2) v), v in [0,4): the angle function of the SimpleEM normal generator (sde_em.cuh)."""
    import mpmath as mp
    mp.mp.dps = 40
    rng = np.random.default_rng(9)
    v = np.concatenate([rng.uniform(0, 4, 4000), np.arange(0, 4, 0.5), np.arange(0, 4, 0.5) + 2.0 ** -51,
                        [np.nextafter(4.0, 0.0), 0.5 - 2.0 ** -53, 1.5, 2.5, 3.5]])
    sn, cs = np.empty_like(v), np.empty_like(v)
    emul.emul_sincos_halfpi(v.ctypes.data_as(ctypes.c_void_p), sn.ctypes.data_as(ctypes.c_void_p),
                            cs.ctypes.data_as(ctypes.c_void_p), ctypes.c_long(len(v)))
    es = max(float(abs(mp.sin(mp.pi / 2 * mp.mpf(float(x))) - mp.mpf(float(s)))) for x, s in zip(v, sn))
    ec = max(float(abs(mp.cos(mp.pi / 2 * mp.mpf(float(x))) - mp.mpf(float(c)))) for x, c in zip(v, cs))
    assert es < 2.3e-16 and ec < 2.3e-16, (es, ec)
    assert np.max(np.abs(sn * sn + cs * cs - 1)) < 5e-16


def test_rcp_fast_accuracy(emul):
    rng = np.random.default_rng(3)
    x = 10.0 ** rng.uniform(-12, 6, 5000)
    got = _apply(emul.emul_rcp, x)
    assert np.max(np.abs(got * x - 1.0)) < 3e-16


def test_short_nan_minmax_equal_julia_semantics(emul):
    """max_abs_nan2 / min_abs_nan1 equal Base.max / Base.min of the absolute values whenever their
    stated invariants hold (a NaN => b NaN, resp. b NaN => a NaN)."""
    nan, inf = float("nan"), float("inf")
    vals = [0.0, -0.0, 1.0, -1.0, 2.5, -2.5, 1e-300, -1e300, inf, -inf, nan]

    def same(p, q):
        return (math.isnan(p) and math.isnan(q)) or (p == q and math.copysign(1, p) == math.copysign(1, q))
    for a in vals:
        for b in vals:
            if not (math.isnan(a) and not math.isnan(b)):      # invariant of max_abs_nan2
                assert same(emul.emul_max_abs_nan2(a, b), emul.emul_jl_max(abs(a), abs(b))), (a, b)
            if not (math.isnan(b) and not math.isnan(a)):      # invariant of min_abs_nan1
                assert same(emul.emul_min_abs_nan1(a, b), emul.emul_jl_min(abs(a), abs(b))), (a, b)


def _pow_pair(emul, x, y):
    x = np.ascontiguousarray(x, dtype=np.float64)
    a, b = np.empty_like(x), np.empty_like(x)
    for fn, out in ((emul.emul_pow_glibc, a), (emul.host_libm_pow, b)):
        fn(x.ctypes.data_as(ctypes.c_void_p), ctypes.c_double(y), out.ctypes.data_as(ctypes.c_void_p),
           ctypes.c_long(len(x)))
    return a, b


def _glibc_version():
    try:
        g = ctypes.CDLL(None).gnu_get_libc_version
        g.restype = ctypes.c_char_p
        return tuple(int(v) for v in g().decode().split(".")[:2])
    except Exception:
        return None


needs_glibc_pow = pytest.mark.skipif(
    _glibc_version() is None or _glibc_version() < (2, 28) or "fma" not in open("/proc/cpuinfo").read(),
    reason="sde_pow_glibc restates the FMA variant of the table-driven pow of glibc >= 2.28")


def test_glibc_pow_tables_are_current():
    """sde_glibc_pow_tables_gen.cuh is what tools/gen_glibc_pow_tables.py reads out of this host's libm."""
    if _glibc_version() is None or _glibc_version() < (2, 28):
        pytest.skip("no glibc >= 2.28 here")
    path = os.path.join(ROOT, "simplediffeq.jl_b200", "csrc", "device", "sde_glibc_pow_tables_gen.cuh")
    before = open(path).read()
    try:
        subprocess.check_call([sys.executable, os.path.join(ROOT, "tools", "gen_glibc_pow_tables.py")],
                              stdout=subprocess.DEVNULL)
        assert open(path).read() == before
    finally:
        open(path, "w").write(before)


@needs_glibc_pow
@pytest.mark.parametrize("y", [7.0 / 50.0, 2.0 / 25.0, 0.5, 1.0 / 3.0, 2.5, -0.14, 1.0, 17.0])
def test_strict_controller_pow_is_the_host_libm_pow_bit_for_bit(emul, y):
    """The device header's sde_pow_glibc (compiled for the host: the same IEEE operations the GPU executes
    under -fmad=false) against the C library the oracle is linked with, on the controller's domain: error
    estimates from 1e-300 to 1e+300, dense around 1 (where the accept/reject decision and the |y log x| <
    2^-54 shortcut live), qold's clamp value 1e-4, subnormals, and the special values."""
    rng = np.random.default_rng(int(abs(y) * 1000) + 20261017)
    x = np.concatenate([
        10.0 ** rng.uniform(-300, 300, 150_000), 10.0 ** rng.uniform(-12, 4, 200_000),
        rng.uniform(0.25, 4.0, 100_000), 1 + rng.uniform(-1e-6, 1e-6, 20_000),
        1 + rng.uniform(-1e-15, 1e-15, 5_000), 2.0 ** np.arange(-1074, 1024).astype(np.float64),
        rng.uniform(0, 1, 5_000) * 2.0 ** -1022,
        np.array([1e-4, 1.0, np.nextafter(1.0, 0), np.nextafter(1.0, 2), 0.0, np.inf, np.nan, 5e-324,
                  np.finfo(np.float64).max, np.finfo(np.float64).tiny])])
    got, ref = _pow_pair(emul, x, y)
    nan = np.isnan(ref)
    assert np.array_equal(np.isnan(got), nan)
    bad = np.flatnonzero(got[~nan].view(np.uint64) != ref[~nan].view(np.uint64))
    assert bad.size == 0, (bad.size, x[~nan][bad[:5]], got[~nan][bad[:5]], ref[~nan][bad[:5]])


@needs_glibc_pow
@pytest.mark.parametrize("y", [np.float32(7.0 / 50.0), np.float32(2.0 / 25.0), np.float32(0.5), np.float32(-1.75),
                               np.float32(1.0), np.float32(3.0)])
def test_strict_controller_powf_is_the_host_libm_powf_bit_for_bit(emul, y):
    """Float32 twin: sde_powf_glibc against the C library's powf on every 97th Float32 bit pattern (44 M
    values: all exponents, both signs, subnormals, inf, NaN) and a dense band around 1."""
    bits = np.arange(0, 2 ** 32, 97, dtype=np.uint64).astype(np.uint32)
    one = np.float32(1.0).view(np.uint32)
    band = (np.arange(-200_000, 200_000, dtype=np.int64) + int(one)).astype(np.uint32)
    x = np.ascontiguousarray(np.concatenate([bits, band]).view(np.float32))
    a, b = np.empty_like(x), np.empty_like(x)
    for fn, out in ((emul.emul_powf_glibc, a), (emul.host_libm_powf, b)):
        fn(x.ctypes.data_as(ctypes.c_void_p), ctypes.c_float(float(y)), out.ctypes.data_as(ctypes.c_void_p),
           ctypes.c_long(len(x)))
    nan = np.isnan(b)
    assert np.array_equal(np.isnan(a), nan)
    bad = np.flatnonzero(a[~nan].view(np.uint32) != b[~nan].view(np.uint32))
    assert bad.size == 0, (bad.size, x[~nan][bad[:5]], a[~nan][bad[:5]], b[~nan][bad[:5]])


@needs_glibc_pow
def test_glibc_pow_restatement_for_arbitrary_exponents(emul):
    """Beyond the controller's two exponents: random (x, y) pairs over the whole main-path domain (results from
    subnormal to near overflow, |y| from 2^-60 to 2^40, negative y), FP64 and FP32 -- the restatement is the libm
    function, not a fit to two exponents.  (Arguments that leave the main path fall back to the platform pow in both
    builds and are trivially equal here; the share that stays on the main path is asserted.)"""
    rng = np.random.default_rng(99)
    n = 2_000_000
    x = np.ascontiguousarray(np.exp(rng.uniform(-700, 700, n)))
    y = np.ascontiguousarray(rng.choice([-1.0, 1.0], n) * np.exp(rng.uniform(np.log(2.0 ** -60), np.log(2.0 ** 40), n)))
    a, b = np.empty(n), np.empty(n)
    for fn, out in ((emul.emul_pow_glibc_xy, a), (emul.host_libm_pow_xy, b)):
        fn(x.ctypes.data_as(ctypes.c_void_p), y.ctypes.data_as(ctypes.c_void_p), out.ctypes.data_as(ctypes.c_void_p),
           ctypes.c_long(n))
    main_path = np.abs(y * np.log(x)) < 500
    assert main_path.mean() > 0.5
    assert np.array_equal(a.view(np.uint64), b.view(np.uint64))
    xf = np.ascontiguousarray(np.exp(rng.uniform(-85, 85, n)).astype(np.float32))
    yf = np.ascontiguousarray((rng.choice([-1.0, 1.0], n) * np.exp(rng.uniform(-20, 8, n))).astype(np.float32))
    af, bf = np.empty(n, np.float32), np.empty(n, np.float32)
    for fn, out in ((emul.emul_powf_glibc_xy, af), (emul.host_libm_powf_xy, bf)):
        fn(xf.ctypes.data_as(ctypes.c_void_p), yf.ctypes.data_as(ctypes.c_void_p), out.ctypes.data_as(ctypes.c_void_p),
           ctypes.c_long(n))
    assert (np.abs(yf.astype(np.float64) * np.log2(xf.astype(np.float64))) < 120).mean() > 0.5
    assert np.array_equal(af.view(np.uint32), bf.view(np.uint32))
