"""CPU tests of the host side: Julia range semantics, the C ABI surface (library loads, exports every
symbol the header declares, fails loudly without a GPU), the NVRTC path (compiles for sm_100a with
no device present), the generated-code spec, and the Python mirror of the reference interface."""
import ctypes
import os
import re
import subprocess
import sys
from fractions import Fraction

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


# ---------------------------------------------------------------------------------------------
# Julia ranges (simplediffeq.jl_b200/jlrange.py)
# ---------------------------------------------------------------------------------------------
def test_range_lengths_match_julia(sde):
    R = sde.JuliaRange
    assert len(R(0.0, 0.1, 1.0)) == 11              # naive floor((1-0)/0.1) would also say 11, but:
    assert len(R(0.0, 0.1, 0.3)) == 4               # 3*0.1 > 0.3 in binary; Julia lifts to rationals
    assert len(R(0.0, 0.001, 10.0)) == 10001
    assert len(R(0.0, 0.01, 10.0)) == 1001
    assert len(R(1.0, 0.5, 0.0)) == 0
    assert len(R(0.0, 1.0, 0.0)) == 1
    # reference default dt = 0.1f0 with a Float64 tspan (quirk Q5): 99 steps, last point ~9.9
    r = R(0.0, float(np.float32(0.1)), 10.0)
    assert len(r) == 100 and r.rational is None
    assert abs(r.collect()[-1] - 9.9) < 1e-6
    assert len(R(0, 0.01, 1, np.float32)) == 101


def test_range_elements_are_exactly_rounded_rationals(sde):
    g = sde.jl_range(0.0, 0.001, 10.0)
    assert g[0] == 0.0 and g[-1] == 10.0
    for i in (1, 3, 7, 333, 2501, 9999):
        assert g[i] == float(Fraction(i, 1000))
    assert g[9] == 0.009 and 9 * 0.001 != 0.009      # not an accumulated / multiplied sum
    h = sde.jl_range(0.1, 0.1, 0.5)
    assert list(h) == [0.1, 0.2, 0.3, 0.4, 0.5]
    f = sde.jl_range(0, 0.01, 1, np.float32)
    assert f.dtype == np.float32 and f[-1] == np.float32(1.0)
    assert all(f[i] == np.float32(i / 100) for i in range(101))


def test_range_fallback_is_literal(sde):
    start, step = 0.0, float(np.float32(0.1))
    g = sde.jl_range(start, step, 10.0)
    assert all(g[i] == start + i * step for i in range(len(g)))


# ---------------------------------------------------------------------------------------------
# C ABI
# ---------------------------------------------------------------------------------------------
def _header_functions():
    text = open(os.path.join(ROOT, "include", "simplediffeq_cuda.h")).read()
    return re.findall(r"^SDE_API\s+[\w\s\*]+?\b(sde_\w+)\s*\(", text, re.M)


def test_library_exports_every_declared_symbol(sde):
    from simplediffeq_b200 import _lib
    L = _lib.lib()
    names = _header_functions()
    assert len(names) >= 15
    for n in names:
        assert hasattr(L, n), "missing export " + n
    assert sorted(names) == sorted(_lib.EXPORTS)
    out = subprocess.run(["nm", "-D", "--defined-only", _lib.LIB_PATH], capture_output=True, text=True).stdout
    exported = set(re.findall(r" T (\w+)", out))
    assert set(names) <= exported
    assert all(e.startswith("sde_") for e in exported), exported   # nothing else leaks out
    assert L.sde_version() == 100


def test_header_ids_match_python_ids(sde):
    from simplediffeq_b200 import _lib
    text = open(os.path.join(ROOT, "include", "simplediffeq_cuda.h")).read()
    ids = dict((k, int(v)) for k, v in re.findall(r"\b(SDE_[A-Z0-9_]+)\s*=\s*(-?\d+)", text))
    assert ids["SDE_ALG_TSIT5"] == _lib.ALG_IDS["GPUSimpleTsit5"] == 0
    assert ids["SDE_ALG_ATSIT5"] == _lib.ALG_IDS["GPUSimpleATsit5"]
    assert ids["SDE_ALG_RK4"] == _lib.ALG_IDS["GPUSimpleRK4"]
    assert ids["SDE_ALG_VERN7"] == _lib.ALG_IDS["GPUSimpleVern7"]
    assert ids["SDE_ALG_AVERN7"] == _lib.ALG_IDS["GPUSimpleAVern7"]
    assert ids["SDE_ALG_VERN9"] == _lib.ALG_IDS["GPUSimpleVern9"]
    assert ids["SDE_ALG_AVERN9"] == _lib.ALG_IDS["GPUSimpleAVern9"]
    assert ids["SDE_ALG_EULER"] == _lib.ALG_IDS["GPUSimpleEuler"] == 7
    assert (ids["SDE_SAVE_ENDPOINT"], ids["SDE_SAVE_SAVEAT"], ids["SDE_SAVE_EVERYSTEP"]) == (0, 1, 2)
    assert (ids["SDE_LAYOUT_TRAJ_MAJOR"], ids["SDE_LAYOUT_SOA"]) == (_lib.LAYOUT_TRAJ_MAJOR, _lib.LAYOUT_SOA)
    assert ids["SDE_COMPAT_FIX_VERN9_INTERP"] == _lib.COMPAT_FIX_VERN9_INTERP
    assert ids["SDE_COMPAT_STRICT_CONTROLLER"] == _lib.COMPAT_STRICT_CONTROLLER
    assert ids["SDE_COMPAT_LOG2_CONTROLLER"] == _lib.COMPAT_LOG2_CONTROLLER
    assert ids["SDE_COMPAT_FAST_RHS"] == _lib.COMPAT_FAST_RHS
    assert ids["SDE_COMPAT_FAST_STAGES"] == _lib.COMPAT_FAST_STAGES
    assert ctypes.sizeof(_lib.SdeOptions) == 6 * 4 + 8 + 5 * 8 + 8 + 8 + 8 + 8 + 8 + 8


def test_builtin_registry(sde):
    dims = dict(lorenz=(3, 3), vanderpol=(2, 1), robertson=(3, 3), nbody=(12, 3), lineardecay=(3, 3),
                scalargrowth=(1, 1), nonautonomous=(2, 2))
    for name, (ns, npar) in dims.items():
        s = getattr(sde.systems, name)
        assert (s.n_state, s.n_param) == (ns, npar)
    with pytest.raises(Exception):
        sde.builtin_system("no-such-system")


def test_every_builtin_kernel_exists(sde):
    """sde_system_prepare = 'is this (system, alg, dtype, save mode) in the build'."""
    from simplediffeq_b200 import _lib
    L = _lib.lib()
    algs = [sde.GPUSimpleTsit5(), sde.GPUSimpleATsit5(), sde.GPUSimpleRK4(), sde.GPUSimpleVern7(),
            sde.GPUSimpleAVern7(), sde.GPUSimpleVern9(), sde.GPUSimpleAVern9(), sde.GPUSimpleEuler()]
    n_ok = 0
    for name in sde.systems.names():
        sysm = getattr(sde.systems, name)
        for alg in algs:
            for dtype in (np.float64, np.float32):
                for mode in (0, 1, 2):
                    keep = []
                    o = sde.api.make_options(alg, np.dtype(dtype), 4, (0.0, 1.0), 0.1, 1e-6, 1e-3,
                                             np.array([0.5]) if mode == 1 else None, mode, 0, 0, 0, keep, out_capacity=16)
                    rc = L.sde_system_prepare(sysm._handle, ctypes.byref(o))
                    unsupported = isinstance(alg, (sde.GPUSimpleRK4, sde.GPUSimpleEuler)) and mode == 1
                    assert (rc == -4) if unsupported else (rc == 0), (name, alg, dtype, mode, L.sde_last_error())
                    n_ok += rc == 0
                    if rc == 0:     # every variant the compat flags select is in the build too (layouts / fast flags / Q2 fix)
                        for compat, layout in ((_lib.COMPAT_FAST_RHS | _lib.COMPAT_FAST_STAGES, 0), (_lib.COMPAT_FAST_STAGES, 1),
                                               (_lib.COMPAT_FIX_VERN9_INTERP, 1)):
                            o.compat, o.layout = compat, layout
                            assert L.sde_system_prepare(sysm._handle, ctypes.byref(o)) == 0, (name, alg, dtype, mode, compat, layout)
    assert n_ok == 7 * 2 * (3 + 3 + 2 + 3 + 3 + 3 + 3 + 2)


def test_option_validation_errors(sde):
    from simplediffeq_b200 import _lib
    L = _lib.lib()
    keep = []
    o = sde.api.make_options(sde.GPUSimpleTsit5(), np.dtype(np.float64), 4, (0.0, 1.0), 0.1, 1e-6, 1e-3, None, 0, 0, 0, 0, keep)
    o.alg = 99
    assert L.sde_system_prepare(sde.systems.lorenz._handle, ctypes.byref(o)) == -1
    assert b"algorithm" in L.sde_last_error()
    o.alg, o.dtype = 0, 7
    assert L.sde_system_prepare(sde.systems.lorenz._handle, ctypes.byref(o)) == -1
    assert L.sde_system_prepare(None, ctypes.byref(o)) == -1
    o.dtype = 0
    assert L.sde_system_prepare(sde.systems.lorenz._handle, ctypes.byref(o)) == 0
    for good in (_lib.COMPAT_FAST_RHS, _lib.COMPAT_FAST_STAGES, _lib.COMPAT_FAST_RHS | _lib.COMPAT_FAST_STAGES | _lib.COMPAT_FIX_VERN9_INTERP):
        o.compat = good                   # the opt-in fast flags combine freely with each other and with the others
        assert L.sde_system_prepare(sde.systems.lorenz._handle, ctypes.byref(o)) == 0
    for bad in (32, 0x40000000, 2 | 4):   # unknown compat bits (bit 30 must reach the kernels as 0: sde::late_flag);
                                         # literal and log2-domain controller forced at the same time
        o.compat = bad
        assert L.sde_system_prepare(sde.systems.lorenz._handle, ctypes.byref(o)) == -1
        assert b"compat" in L.sde_last_error()


def test_no_cpu_fallback_fails_loudly_without_gpu(sde):
    from simplediffeq_b200 import _lib
    if _lib.device_count() > 0:
        pytest.skip("a GPU is present")
    u0 = np.zeros((3, 8)); p = np.ones((3, 8))
    with pytest.raises(_lib.SdeError) as e:
        sde.solve_arrays(sde.systems.lorenz, sde.GPUSimpleTsit5(), u0, p, (0.0, 1.0), dt=0.1)
    assert e.value.code == -2          # SDE_ERR_CUDA, not a silently computed answer


# ---------------------------------------------------------------------------------------------
# NVRTC user right-hand sides (compile only; no device needed)
# ---------------------------------------------------------------------------------------------
LORENZ_SRC = """
__device__ void rhs(real* du, const real* u, const real* p, real t) {
  du[0] = p[0] * (u[1] - u[0]);
  du[1] = u[0] * (p[1] - u[2]) - u[1];
  du[2] = u[0] * u[1] - p[2] * u[2];
}"""


def test_nvrtc_user_rhs_compiles_for_sm100a(sde):
    from simplediffeq_b200 import _lib
    L = _lib.lib()
    user = sde.CudaRHS(LORENZ_SRC, 3, 3)
    for alg, mode in ((sde.GPUSimpleTsit5(), 0), (sde.GPUSimpleTsit5(), 1), (sde.GPUSimpleATsit5(), 0),
                      (sde.GPUSimpleRK4(), 2), (sde.GPUSimpleAVern9(), 1)):
        keep = []
        o = sde.api.make_options(alg, np.dtype(np.float64), 4, (0.0, 1.0), 0.1, 1e-6, 1e-3,
                                 np.array([0.5]) if mode == 1 else None, mode, 0, 0, 0, keep)
        assert L.sde_system_prepare(user._handle, ctypes.byref(o)) == 0, L.sde_last_error()
    # the literal controller (its pow / powf are sde_pow_glibc / sde_powf_glibc with their __device__ tables)
    for dt_ in (np.float64, np.float32):
        keep = []
        o = sde.api.make_options(sde.GPUSimpleAVern7(), np.dtype(dt_), 4, (0.0, 1.0), 0.1, 1e-6, 1e-3, None, 0, 0,
                                 _lib.COMPAT_STRICT_CONTROLLER, 0, keep)
        assert L.sde_system_prepare(user._handle, ctypes.byref(o)) == 0, L.sde_last_error()


def test_nvrtc_cubin_disk_cache(sde, tmp_path, monkeypatch, capfd):
    """SDE_CACHE_DIR: a user RHS is compiled once per (source, variant) per machine; a second handle with the
    same source loads the cubin from disk (reported under SDE_TRACE), a different source does not."""
    from simplediffeq_b200 import _lib
    L = _lib.lib()
    monkeypatch.setenv("SDE_CACHE_DIR", str(tmp_path / "cache"))
    monkeypatch.setenv("SDE_TRACE", "1")

    def prepare(src):
        user = sde.CudaRHS(src, 3, 3)
        keep = []
        o = sde.api.make_options(sde.GPUSimpleTsit5(), np.dtype(np.float64), 4, (0.0, 1.0), 0.1, 1e-6, 1e-3, None, 0, 0, 0, 0, keep)
        assert L.sde_system_prepare(user._handle, ctypes.byref(o)) == 0, L.sde_last_error()
        return capfd.readouterr().err

    first = prepare(LORENZ_SRC)
    files = sorted((tmp_path / "cache").glob("*.cubin"))
    assert "compiled" in first and "cache hit" not in first and len(files) == 1 and files[0].stat().st_size > 1000
    second = prepare(LORENZ_SRC)
    assert "cache hit" in second and "compiled" not in second
    third = prepare(LORENZ_SRC.replace("p[2] * u[2]", "p[2] * u[2] * 1"))
    assert "compiled" in third and len(list((tmp_path / "cache").glob("*.cubin"))) == 2
    monkeypatch.setenv("SDE_CACHE_DIR", "off")
    assert "cache hit" not in prepare(LORENZ_SRC)


def test_nvrtc_compile_error_is_reported(sde):
    from simplediffeq_b200 import _lib
    with pytest.raises(_lib.SdeError) as e:
        sde.CudaRHS("__device__ void rhs(real* du, const real* u, const real* p, real t) { du[0] = nope; }", 1, 1)
    assert e.value.code == -3 and "nope" in str(e.value)
    with pytest.raises(_lib.SdeError):
        sde.CudaRHS(LORENZ_SRC, 0, 3)


# ---------------------------------------------------------------------------------------------
# generated code / spec
# ---------------------------------------------------------------------------------------------
def test_generated_headers_are_current_and_spec_matches_reference():
    if not os.path.isdir("/root/reference"):
        pytest.skip("reference tree not present on this machine")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "check_spec_vs_reference.py")],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    # regenerating must reproduce the committed files byte for byte
    import importlib.util
    spec = importlib.util.spec_from_file_location("gen_tableaus", os.path.join(ROOT, "tools", "gen_tableaus.py"))
    G = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(G)
    ts, v = G.parse_tsit5(), G.parse_verner()
    specs = [G.spec_tsit5(), G.spec_vern7(v), G.spec_vern9(v)]
    pkg = os.path.join(ROOT, "simplediffeq.jl_b200", "csrc", "device")
    assert open(os.path.join(ROOT, "oracle", "tableau_named.hpp")).read() == G.emit_oracle(ts, v)
    assert open(os.path.join(pkg, "sde_tableaus_gen.cuh")).read() == G.emit_device_tables(ts, v)
    assert open(os.path.join(pkg, "sde_methods_gen.cuh")).read() == G.emit_methods(specs)


def test_tableau_consistency():
    """Row sums equal the nodes and the weights sum to one (parsed from the oracle's header)."""
    text = open(os.path.join(ROOT, "oracle", "tableau_named.hpp")).read()

    def block(name):
        body = re.search(r"struct %s \{(.*?)\};" % name, text, re.S).group(1)
        return dict((k, float(v)) for k, v in re.findall(r"static constexpr T (\w+) = T\(([-0-9.e+]+)\);", body))
    t5 = block("Tsit5Tab")
    assert abs(t5["a31"] + t5["a32"] - t5["c2"]) < 1e-15
    assert abs(sum(t5["a7%d" % j] for j in range(1, 7)) - 1.0) < 1e-15
    assert abs(sum(t5["btilde%d" % j] for j in range(1, 8))) < 1e-15
    v7 = block("Vern7Tab")
    for s in range(3, 9):
        row = sum(v for k, v in v7.items() if re.fullmatch(r"a0%d\d" % s, k))
        assert abs(row - v7["c%d" % s]) < 1e-14, s
    assert abs(sum(v7[k] for k in ("b1", "b4", "b5", "b6", "b7", "b8", "b9")) - 1.0) < 1e-15
    v9 = block("Vern9Tab")
    for s in range(3, 15):
        row = sum(v for k, v in v9.items() if re.fullmatch(r"a%02d\d\d" % s, k))
        assert abs(row - v9["c%d" % (s - 1)]) < 1e-14, s
    assert abs(sum(v9["b%d" % j] for j in (1, 8, 9, 10, 11, 12, 13, 14, 15)) - 1.0) < 1e-15
    for s in range(17, 27):
        row = sum(v for k, v in v9.items() if re.fullmatch(r"a%02d\d\d" % s, k))
        assert abs(row - v9["c%d" % s]) < 1e-14, s


# ---------------------------------------------------------------------------------------------
# Python mirror of the reference interface (argument handling; no device needed)
# ---------------------------------------------------------------------------------------------
def test_problem_types_and_defaults(sde):
    prob = sde.ODEProblem(sde.systems.lorenz, [1.0, 0.0, 0.0], (0.0, 10.0), [10.0, 28.0, 8 / 3])
    assert prob.dtype == np.float64
    p32 = sde.ODEProblem(sde.systems.lorenz, np.array([1, 0, 0], np.float32), (0.0, 1.0), [10, 28, 8 / 3])
    assert p32.dtype == np.float32 and p32.p.dtype == np.float32
    with pytest.raises(ValueError):
        sde.ODEProblem(sde.systems.lorenz, [1.0, 0.0], (0.0, 1.0), [10.0, 28.0, 8 / 3])
    with pytest.raises(TypeError):
        sde.ODEProblem(lambda u, p, t: u, [1.0], (0.0, 1.0))     # host callables cannot run on the GPU path
    with pytest.raises(ValueError, match="dt is required"):      # src/rk4/gpurk4.jl:56
        sde.solve(prob, sde.GPUSimpleRK4())
    q = sde.remake(prob, p=[10.0, 5.0, 1.0])
    assert q.p[1] == 5.0 and q.f is prob.f
    with pytest.raises(TypeError):
        sde.solve(sde.EnsembleProblem(prob), sde.GPUSimpleTsit5())    # trajectories missing
    # save-mode selection mirrors the keyword logic of src/tsit5/gpuatsit5.jl:71-83,112-134
    from simplediffeq_b200.api import _save_mode
    assert _save_mode(sde.GPUSimpleTsit5(), None, True) == 2
    assert _save_mode(sde.GPUSimpleTsit5(), None, False) == 0
    assert _save_mode(sde.GPUSimpleTsit5(), [0.5], True) == 1
    assert _save_mode(sde.GPUSimpleRK4(), [0.5], False) == 2        # RK4 swallows both keywords
    assert _save_mode(sde.GPUSimpleATsit5(), None, True) == 2        # adaptive save_everystep=true: variable length


def test_fixed_times_match_the_reference_rules(sde):
    """sol.t of fixed-step solves: t = _ts[i-1] + dt (gpuatsit5.jl:98,111,114); RK4: the range itself."""
    from simplediffeq_b200 import _lib
    keep = []
    o = sde.api.make_options(sde.GPUSimpleTsit5(), np.dtype(np.float64), 1, (0.0, 1.0), 0.1, 0, 0, None, 2, 0, 0, 0, keep)
    t = sde.api.fixed_times(o, np.float64)
    g = sde.jl_range(0.0, 0.1, 1.0)
    assert len(t) == 11 and t[0] == 0.0
    assert all(t[k] == g[k - 1] + 0.1 for k in range(1, 11))
    assert t[3] != g[3]                       # 0.2 + 0.1 != 0.3: the reference's ts are NOT the range
    o = sde.api.make_options(sde.GPUSimpleRK4(), np.dtype(np.float64), 1, (0.0, 1.0), 0.1, 0, 0, None, 2, 0, 0, 0, keep)
    assert np.array_equal(sde.api.fixed_times(o, np.float64), g)
    o = sde.api.make_options(sde.GPUSimpleTsit5(), np.dtype(np.float64), 1, (0.0, 1.0), 0.1, 0, 0, None, 0, 0, 0, 0, keep)
    assert list(sde.api.fixed_times(o, np.float64)) == [0.0, g[9] + 0.1]


def test_fixed_times_match_reference_source_execution(sde):
    """`sol.t` of the fixed-step solves (sde_fixed_times + the Julia range restatement, host code only)
    against the times the reference's own source produced (tests/golden/golden_jlmini_v1.json)."""
    import jlmini_cases as J
    from common import bits_equal
    api = sde.api if hasattr(sde, "api") else __import__("simplediffeq_b200.api", fromlist=["api"])
    checked = 0
    for case in J.load_cases():
        if case["alg"] in J.ADAPTIVE or "error" in case:
            continue
        a = J.case_inputs(case)
        exp_t, _ = J.expected(case)
        alg = getattr(sde, case["alg"])()
        mode = {"endpoint": 0, "saveat": 1, "everystep": 2}[a["kind"]]
        keep = []
        o = api.make_options(alg, np.dtype(a["dtype"]), 1, (a["t0"], a["tf"]), a["dt"], a["abstol"], a["reltol"],
                             a["saveat"], mode, 0, 0, 0, keep)
        t = api.fixed_times(o, a["dtype"])
        assert len(t) == len(exp_t), case["name"]
        assert bits_equal(t.astype(exp_t.dtype), exp_t), case["name"]     # ts has eltype(dt) (quirk Q11)
        checked += 1
    assert checked >= 50


def test_bench_reference_arm_prints_one_contract_line():
    """`bench.py --impl reference` (the CPU arm the driver runs next to ours) needs no GPU: exactly one JSON line on
    stdout with the contract's keys; under torchrun only rank 0 prints."""
    import json
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    cmd = [sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1"]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    for key in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better",
                "scaling", "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert key in d, key
    assert d["impl"] == "reference" and d["metric"] == "trajectory-steps/s" and d["value"] > 1e6
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0
    env = dict(os.environ, RANK="1", WORLD_SIZE="2")
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, env=env)
    assert r.returncode == 0 and r.stdout.strip() == ""          # the other ranks exit 0 without work


def test_prob_func_restrictions_are_checked_on_the_host(sde):
    """The ensemble path runs ONE kernel per ensemble: prob_func may change u0 and p only.  A changed tspan, system or
    element type is rejected before anything is launched (no GPU needed to see the error)."""
    prob = sde.ODEProblem(sde.systems.lorenz, [1.0, 0.0, 0.0], (0.0, 1.0), [10.0, 28.0, 8 / 3])
    bad_tspan = sde.EnsembleProblem(prob, prob_func=lambda pr, i, rep: sde.remake(pr, tspan=(0.0, 2.0)))
    with pytest.raises(ValueError, match="only change u0 and p"):
        sde.solve(bad_tspan, sde.GPUSimpleTsit5(), trajectories=3, dt=0.1)
    other = sde.ODEProblem(sde.systems.lineardecay, [1.0, 1.0, 1.0], (0.0, 1.0), [10.0, 28.0, 8 / 3])
    bad_sys = sde.EnsembleProblem(prob, prob_func=lambda pr, i, rep: other)
    with pytest.raises(ValueError, match="only change u0 and p"):
        sde.solve(bad_sys, sde.GPUSimpleTsit5(), trajectories=3, dt=0.1)
    p32 = sde.ODEProblem(sde.systems.lorenz, np.array([1, 0, 0], np.float32), (0.0, 1.0), [10, 28, 8 / 3])
    bad_type = sde.EnsembleProblem(prob, prob_func=lambda pr, i, rep: p32)
    with pytest.raises(ValueError, match="only change u0 and p"):
        sde.solve(bad_type, sde.GPUSimpleTsit5(), trajectories=3, dt=0.1)
    # prob_func is called with 1-based indices and repeat = 1, like SciMLBase's batch_func
    seen = []

    def pf(pr, i, rep):
        seen.append((i, rep))
        if i == 3:
            raise KeyError("stop here")     # before any device work
        return pr
    with pytest.raises(KeyError):
        sde.solve(sde.EnsembleProblem(prob, prob_func=pf), sde.GPUSimpleTsit5(), trajectories=5, dt=0.1)
    assert seen == [(1, 1), (2, 1), (3, 1)]
