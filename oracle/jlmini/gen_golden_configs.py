#!/usr/bin/env python
"""tests/golden/golden_jlmini_configs_v1.json: the adaptive BASELINE configurations executed by the reference's OWN
source text (oracle/jlmini over src/tsit5/gpuatsit5.jl, src/verner/gpuvern7.jl, src/verner/gpuvern9.jl), with
save_everystep = true so that the WHOLE accepted-step sequence (every time, every state) is on record, not only the
end point: config 1 (Lorenz rho-sweep, ATsit5, 1e-8), config 3 (Van der Pol mu-sweep, ATsit5, 1e-6), AVern7 at 1e-10
and config 4 (AVern9 at 1e-12 -- the configuration whose step sequence depends on the last bit of `EEst^beta1`).
`@fastmath ^` is the C library's pow here (jlmini.py), i.e. the function csrc/device/sde_common.cuh's sde_pow_glibc
restates.  Test infrastructure; needs /root/reference, so the fixture is committed.
    python oracle/jlmini/gen_golden_configs.py"""
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import gen_golden as G  # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.dirname(HERE)), "tests", "golden", "golden_jlmini_configs_v1.json")


def main():
    if not G.R.available():
        print("reference tree not present at %s; nothing generated" % G.R.REF)
        return 1
    del G.CASES[:]
    F01 = G.F01
    n = 10000                                   # config 1's sweep: rho_i = 21 (i - 1) / (N - 1)
    for i in (1, 1234, 3333, 5000, 7777, 10000):
        rho = float(21 * (i - 1)) / float(n - 1)
        G.case("config1_atsit5_i%d" % i, "GPUSimpleATsit5", "lorenz", [1.0, 0.0, 0.0], [10.0, rho, 8.0 / 3.0], (0.0, 10.0),
               dt=F01, abstol=1e-8, reltol=1e-8)
    m = 1 << 20                                 # config 3: mu_i = 0.1 + 49.9 (i - 1) / (N - 1)
    for i in (1, 100000, 400000, 700000, 1000000, m):
        mu = 0.1 + 49.9 * float(i - 1) / float(m - 1)
        G.case("config3_atsit5_vdp_i%d" % i, "GPUSimpleATsit5", "vanderpol", [2.0, 0.0], [mu], (0.0, 20.0),
               dt=F01, abstol=1e-6, reltol=1e-6)
    k = 1000000                                 # config 4: Lorenz sweep with N = 10^6
    for i in (1, 54321, 250000, 400000, 600000, 800000, 987654, k):
        rho = float(21 * (i - 1)) / float(k - 1)
        G.case("config4_avern9_i%d" % i, "GPUSimpleAVern9", "lorenz", [1.0, 0.0, 0.0], [10.0, rho, 8.0 / 3.0], (0.0, 10.0),
               dt=F01, abstol=1e-12, reltol=1e-12)
    for i in (1, 333333, 666667, k):
        rho = float(21 * (i - 1)) / float(k - 1)
        G.case("avern7_1e-10_i%d" % i, "GPUSimpleAVern7", "lorenz", [1.0, 0.0, 0.0], [10.0, rho, 8.0 / 3.0], (0.0, 10.0),
               dt=F01, abstol=1e-10, reltol=1e-10)
    res = [G.run(c) for c in G.CASES]
    doc = {"generator": "oracle/jlmini/gen_golden_configs.py (jlmini interpreter over the reference's own source files)",
           "reference": "SciML/SimpleDiffEq.jl v1.16.3 at /root/reference", "cases": res}
    with open(OUT, "w") as fh:
        json.dump(doc, fh, indent=0, separators=(",", ":"))
    print("wrote %s: %d cases, %.1f KB; accepted steps + 1: %s" % (OUT, len(res), os.path.getsize(OUT) / 1e3,
                                                               [c["n_out"] for c in res]))
    return 0


if __name__ == "__main__":
    sys.exit(main())
