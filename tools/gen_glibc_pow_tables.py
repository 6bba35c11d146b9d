#!/usr/bin/env python
"""Emit csrc/device/sde_glibc_pow_tables_gen.cuh: the lookup tables and polynomial coefficients of
the `pow` of the C library the ORACLE is linked against (glibc >= 2.28, the table-driven
log/exp algorithm: 128-entry log table with a double-double tail, 128-entry 2^(k/128) table).

Why: the strict controller (SDE_COMPAT_STRICT_CONTROLLER) evaluates `EEst^beta1` / `qold^beta2`
literally like gpuatsit5.jl:279-283.  At tolerances where the error estimate is rounding noise
(BASELINE config 4) the accepted-step count depends on the last bit of that pow, so agreement with
the oracle needs the oracle's pow, bit for bit.  The tables are read out of the
libm this interpreter process would load (they are numerical constants, not code); the operation
sequence in sde_common.cuh (sde_pow_glibc) was transcribed from what that libm executes and is
pinned against it bit for bit by tests/test_ctrl_math.py.

The tables are found by signature (ln2hi, ln2lo / InvLn2N, Shift), not by address.
Usage: python tools/gen_glibc_pow_tables.py [path/to/libm.so.6]
"""
import os
import struct
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "simplediffeq.jl_b200", "csrc", "device", "sde_glibc_pow_tables_gen.cuh")
N_LOG = 128
N_EXP = 128


def find_once(blob, sig, what):
    i = blob.find(sig)
    if i < 0 or blob.find(sig, i + 1) >= 0:
        raise SystemExit("cannot locate %s uniquely in libm" % what)
    return i


def extract(path):
    blob = open(path, "rb").read()
    # struct pow_log_data { double ln2hi, ln2lo, poly[7]; struct { double invc, pad, logc, logctail; } tab[128]; }
    # (__log_data starts with the same two constants: take the occurrence whose table starts with
    # pow's first entry {invc = 0x1.6ap+0, pad = 0})
    sig = struct.pack("<QQ", 0x3FE62E42FEFA3800, 0x3D2EF35793C76730)
    first = struct.pack("<dd", float.fromhex("0x1.6ap+0"), 0.0)
    hits, i = [], blob.find(sig)
    while i >= 0:
        if blob[i + 72:i + 88] == first:
            hits.append(i)
        i = blob.find(sig, i + 1)
    if len(hits) != 1:
        raise SystemExit("cannot locate pow_log_data uniquely in libm (%d candidates)" % len(hits))
    i = hits[0]
    head = struct.unpack_from("<9d", blob, i)
    tab = struct.unpack_from("<%dd" % (4 * N_LOG), blob, i + 72)
    assert head[2] == -0.5 and all(tab[4 * k + 1] == 0.0 for k in range(N_LOG))
    assert tab[0] == float.fromhex("0x1.6ap+0") and tab[4 * 127] == float.fromhex("0x1.6cp-1")
    # struct exp_data { double invln2N, shift, negln2hiN, negln2loN, poly[4]; ...; uint64_t tab[256]; }
    # (the members between poly[] and tab[] differ between glibc versions: the table is found by its
    # first entries {0, 0x3ff0000000000000} = tail and bits of 2^0)
    j = find_once(blob, struct.pack("<QQ", 0x40671547652B82FE, 0x4338000000000000), "exp_data")
    ehead = struct.unpack_from("<8d", blob, j)
    k = blob.find(struct.pack("<QQ", 0, 0x3FF0000000000000), j + 64, j + 64 + 512)
    if k < 0 or (k - j) % 8:
        raise SystemExit("cannot locate exp_data.tab")
    etab = struct.unpack_from("<%dQ" % (2 * N_EXP), blob, k)
    import math
    # sanity: entry i holds the bits of 2^(i/128) with the exponent field reduced by i (<< 45) and its tail
    for q in (1, 37, 127):
        sb = (etab[2 * q + 1] + (q << 45)) & 0xFFFFFFFFFFFFFFFF
        v = struct.unpack("<d", struct.pack("<Q", sb))[0]
        assert abs(v - 2.0 ** (q / 128.0)) < 4e-16, (q, v)
    # float: struct powf_log2_data { struct { double invc, logc; } tab[16]; double poly[5]; } and
    # struct exp2f_data { uint64_t tab[32]; double shift_scaled; double poly[3]; ... }
    # (__log2f_data holds the same table followed by a degree-4 polynomial: take the occurrence whose fifth
    # coefficient is powf's 1/ln2)
    fsig = struct.pack("<dd", float.fromhex("0x1.661ec79f8f3bep+0"), float.fromhex("-0x1.efec65b963019p-2"))
    hits, fi = [], blob.find(fsig)
    while fi >= 0:
        if abs(struct.unpack_from("<d", blob, fi + 256 + 32)[0] - 1.0 / math.log(2.0)) < 1e-9:
            hits.append(fi)
        fi = blob.find(fsig, fi + 1)
    if len(hits) != 1:
        raise SystemExit("cannot locate powf_log2_data uniquely in libm (%d candidates)" % len(hits))
    fi = hits[0]
    ftab = struct.unpack_from("<32d", blob, fi)
    fpoly = struct.unpack_from("<5d", blob, fi + 256)
    assert any(ftab[2 * q] == 1.0 and ftab[2 * q + 1] == 0.0 for q in range(16))
    assert abs(fpoly[4] - 1.0 / math.log(2.0)) < 1e-9
    fj = find_once(blob, struct.pack("<QQ", 0x3FF0000000000000, 0x3FEFD9B0D3158574), "exp2f_data")
    f2tab = struct.unpack_from("<32Q", blob, fj)
    f2head = struct.unpack_from("<4d", blob, fj + 256)
    assert f2head[0] == float.fromhex("0x1.8p+47") and abs(f2head[3] - math.log(2.0)) < 1e-9
    return head, tab, ehead, etab, ftab, fpoly, f2tab, f2head


def main():
    path = sys.argv[1] if len(sys.argv) > 1 else None
    if path is None:
        for cand in ("/lib/x86_64-linux-gnu/libm.so.6", "/usr/lib64/libm.so.6", "/lib64/libm.so.6"):
            if os.path.exists(cand):
                path = cand
                break
    head, tab, ehead, etab, ftab, fpoly, f2tab, f2head = extract(path)
    o = []
    o.append("// GENERATED by tools/gen_glibc_pow_tables.py -- do not edit.  Tables and coefficients of glibc's")
    o.append("// table-driven pow (the oracle's libm), used by sde_pow_glibc (sde_common.cuh) on the strict")
    o.append("// controller path only.  Numerical constants read out of libm.so.6; see the generator's docstring.")
    o.append("#pragma once")
    o.append("namespace sde {")
    names = ["kGpLn2Hi", "kGpLn2Lo", "kGpA0", "kGpA1", "kGpA2", "kGpA3", "kGpA4", "kGpA5", "kGpA6"]
    for n, v in zip(names, head):
        o.append("constexpr double %s = %s;   // %s" % (n, repr(v), v.hex()))
    enames = ["kGpInvLn2N", "kGpShift", "kGpNegLn2HiN", "kGpNegLn2LoN", "kGpC2", "kGpC3", "kGpC4", "kGpC5"]
    for n, v in zip(enames, ehead):
        o.append("constexpr double %s = %s;   // %s" % (n, repr(v), v.hex()))
    o.append("// {1/c, log(c) head, log(c) tail} for the 128 sub-intervals of [0x1.69555p-1, 0x1.69555p0)")
    o.append("static __device__ const double k_gpow_log[%d] = {" % (3 * N_LOG))
    for k in range(N_LOG):
        o.append("    %s, %s, %s," % (repr(tab[4 * k]), repr(tab[4 * k + 2]), repr(tab[4 * k + 3])))
    o.append("};")
    o.append("// {tail of 2^(i/128) as double bits, bits of 2^(i/128) - (i << 45)}")
    o.append("static __device__ const unsigned long long k_gpow_exp[%d] = {" % (2 * N_EXP))
    for k in range(N_EXP):
        o.append("    0x%016xULL, 0x%016xULL," % (etab[2 * k], etab[2 * k + 1]))
    o.append("};")
    o.append("// ---- powf: log2 with a 16-entry table, 2^x with a 32-entry table, all in double")
    for n, v in zip(["kGfA0", "kGfA1", "kGfA2", "kGfA3", "kGfA4"], fpoly):
        o.append("constexpr double %s = %s;   // %s" % (n, repr(v), v.hex()))
    for n, v in zip(["kGfShift", "kGfC0", "kGfC1", "kGfC2"], f2head):
        o.append("constexpr double %s = %s;   // %s" % (n, repr(v), v.hex()))
    o.append("// {1/c, log2(c)} for the 16 sub-intervals of [0x1.66p-1, 0x1.66p0)")
    o.append("static __device__ const double k_gpowf_log2[32] = {")
    for k in range(16):
        o.append("    %s, %s," % (repr(ftab[2 * k]), repr(ftab[2 * k + 1])))
    o.append("};")
    o.append("// bits of 2^(i/32) - (i << 47)")
    o.append("static __device__ const unsigned long long k_gpowf_exp2[32] = {")
    for k in range(0, 32, 4):
        o.append("    " + " ".join("0x%016xULL," % v for v in f2tab[k:k + 4]))
    o.append("};")
    o.append("}  // namespace sde")
    open(OUT, "w").write("\n".join(o) + "\n")
    print("wrote", OUT)


if __name__ == "__main__":
    main()
