"""Dry run of tests/test_zz_gpu_strict_bitexact.py on a machine without a GPU: the test BODIES are executed with the
host emulation of the device kernel source (tests/kernel_host_emul.cpp) standing in for the device, at the tests' own
sizes.  It validates the tests' logic (keys, shapes, dtypes, expectations) and, once more, the statement they make;
it is not a GPU result.  Run `python -m pytest tests/test_kernel_host_emul.py -q` once before (builds the emulation).
    python tools/dryrun_gpu_strict_tests.py"""
import ctypes
import glob
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np  # noqa: E402
import simplediffeq_b200 as sde  # noqa: E402
from simplediffeq_b200 import api  # noqa: E402
import oracle_lib as oracle  # noqa: E402
import test_kernel_host_emul as E  # noqa: E402
import test_zz_gpu_strict_bitexact as Z  # noqa: E402


def main():
    libs = glob.glob(os.path.join(ROOT, "tests", "_build", "libkernel_emul_*.so"))
    if not libs:
        sys.exit("build the emulation library first (see the docstring)")
    L = ctypes.CDLL(libs[0])
    vp, ll, d = ctypes.c_void_p, ctypes.c_longlong, ctypes.c_double
    L.emul_solve.restype = ctypes.c_int
    L.emul_solve.argtypes = [ctypes.c_int] * 6 + [ll, vp, vp, d, d, d, d, d, ll, vp, vp, ll, ll, ll, vp, vp, vp, vp, vp]
    names = {getattr(sde.systems, k): k for k in E.SYS_ID if hasattr(sde.systems, k)}

    def run(system, algname, u0, p, tspan, dt, abstol, reltol, save_mode, layout, compat, saveat, out_capacity):
        n_out = len(saveat) if save_mode == 1 else (out_capacity if save_mode == 2 else 1)
        return E._run(L, system, algname, u0, p, tspan, dt, save=save_mode, layout=layout, compat=compat, abstol=abstol,
                      reltol=reltol, saveat=saveat, n_out=n_out)

    def fake_gpu(sde_, system, algname, u0, p, tspan, dt=None, abstol=1e-6, reltol=1e-3, save_mode=0, layout=0, compat=0,
                 saveat=None, out_capacity=0, **kw):              # stands in for test_gpu_parity._gpu
        g = run(system, algname, u0, p, tspan, dt, abstol, reltol, save_mode, layout, compat, saveat, out_capacity)
        out = dict(u=g["u"], naccept=g["naccept"], nreject=g["nreject"], retcode=g["retcode"])
        out["t_series" if save_mode == 2 else "t_final"] = g["t"]
        return out

    def fake_solve_arrays(sysm, alg, u0_soa, p_soa, tspan, dt=None, abstol=None, reltol=None, saveat=None, save_mode=0,
                          layout=0, compat=0, maxiters=0, devices=None, out_capacity=0):     # stands in for api.solve_arrays
        T = u0_soa.dtype.type
        sa = None if saveat is None else np.asarray(saveat, dtype=T)
        g = run(names[sysm], type(alg).__name__, np.ascontiguousarray(u0_soa.T), np.ascontiguousarray(p_soa.T),
                (float(tspan[0]), float(tspan[1])), float(T(dt)), float(T(abstol)), float(T(reltol)), save_mode, layout,
                compat, sa, out_capacity)
        return dict(u=g["u"], t_shared=None, t_final=None if save_mode == 2 else g["t"],
                    t_series=g["t"] if save_mode == 2 else None, naccept=g["naccept"], nreject=g["nreject"],
                    retcode=g["retcode"])

    Z._gpu = fake_gpu
    api.solve_arrays = fake_solve_arrays
    count = 0
    for args in Z.SWEEPS:
        Z.test_strict_controller_fp64_is_the_oracle_bit_for_bit(sde, oracle, True, *args)
        count += 1
    for args in [("lorenz", "GPUSimpleATsit5", (0.0, 10.0), 1e-4), ("vanderpol", "GPUSimpleATsit5", (0.0, 20.0), 1e-3),
                 ("lorenz", "GPUSimpleAVern7", (0.0, 5.0), 1e-5), ("lorenz", "GPUSimpleAVern9", (0.0, 5.0), 1e-5)]:
        Z.test_strict_controller_fp32_is_the_oracle_bit_for_bit(sde, oracle, True, *args)
        count += 1
    for alg in E.ADAPT:
        for T in (np.float64, np.float32):
            Z.test_strict_controller_series_outputs_are_the_oracle_bit_for_bit(sde, oracle, True, alg, T)
            count += 1
    for case in Z._JADAPT:
        if case["system"] in E.SYS_ID:
            Z.test_strict_controller_vs_reference_source_execution_bit_for_bit(sde, True, case)
            count += 1
    print("%d test bodies of tests/test_zz_gpu_strict_bitexact.py pass with the host emulation standing in for the device"
          % count)


if __name__ == "__main__":
    main()
