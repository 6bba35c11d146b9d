#!/usr/bin/env python
"""Emit csrc/device/sde_glibc_pow_tables_gen.cuh: lookup tables and polynomial coefficients of the
table-driven `pow` / `powf` that the ORACLE's C library implements, for the literal step-size
controller (SDE_COMPAT_STRICT_CONTROLLER; sde_common.cuh: gpow_log / gpow_exp / gpowf_*).

Algorithm and constants: Szabolcs Nagy's pow / powf from ARM optimized-routines (math/pow.c,
math/pow_log_data.c, math/exp_data.c, math/powf.c, math/powf_log2_data.c, math/exp2f_data.c; MIT OR
Apache-2.0 WITH LLVM-exception), adopted by glibc 2.28 as sysdeps/ieee754/dbl-64/e_pow.c,
e_pow_log_data.c, e_exp_data.c and sysdeps/ieee754/flt-32/e_powf.c, e_powf_log2_data.c,
e_exp2f_data.c (POW_LOG_TABLE_BITS = 7, EXP_TABLE_BITS = 7, POWF_LOG2_TABLE_BITS = 4,
EXP2F_TABLE_BITS = 5, TOINT_INTRINSICS = 0).

Everything that has a mathematical definition is COMPUTED here (mpmath, 400 bits), following the
comments of those data files:
  pow log table, sub-interval i of [0x1.69555p-1, 0x1.69555p0) (128 equal steps of the bit pattern):
      invc     = 1/c rounded to a multiple of 2^-7 (z < 1) or 2^-8 (z >= 1), c = centre of the sub-interval
      logc     = round(2^43 * log(1/invc)) / 2^43        (so that k*ln2hi + logc is exact)
      logctail = RN(log(1/invc) - logc)
  exp table:  H = RN(2^(i/128)),  T = RN((2^(i/128) - H) / H),  stored as {bits(T), bits(H) - (i << 45)}
  powf log2 table: logc = RN(log2(1/invc)) for the 16 published invc;  exp2f table: bits(RN(2^(i/32))) - (i << 47)
Only the minimax polynomial coefficients, the ln2 splits and powf's 16 invc values are published
constants that have to be quoted (below, as hex floats).

When this host's libm is recognisably that implementation the generated numbers are compared with
the ones inside it (found by signature) and ANY difference is an error: the generator and the
oracle's libm must describe the same function.  tests/test_ctrl_math.py pins the operation sequence
of sde_common.cuh against the host libm bit for bit.

Output layout: two arrays of doubles in exactly the order the kernels copy them to shared memory
(k_gpow, k_gpowf; integer table entries are stored as the double with the same bit pattern -- all of
them are normal numbers), with adjacent constants paired for 16-byte loads.
Usage: python tools/gen_glibc_pow_tables.py [--no-libm-check] [path/to/libm.so.6]
"""
import os
import struct
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "simplediffeq.jl_b200", "csrc", "device", "sde_glibc_pow_tables_gen.cuh")
N_LOG = 128
N_EXP = 128
POW_OFF = 0x3FE6955500000000

H = float.fromhex
# ---- published constants (ARM optimized-routines / glibc data files named in the docstring) ------------------
POW_LN2HI, POW_LN2LO = H("0x1.62e42fefa3800p-1"), H("0x1.ef35793c76730p-45")
# pow_log_data.poly (A[0] = -1/2; the others carry the factors of the source's "* -2", "* 4", "* -8")
POW_A = [H("-0x1.0000000000000p-1"), H("-0x1.5555555555560p-1"), H("0x1.0000000000006p-1"), H("0x1.999999959554ep-1"),
         H("-0x1.555555529a47ap-1"), H("-0x1.2495b9b4845e9p+0"), H("0x1.0002b8b263fc3p+0")]
# exp_data: invln2N = N/ln2, shift = 0x1.8p52, negln2hiN, negln2loN, poly C2..C5
EXP_INVLN2N, EXP_SHIFT = H("0x1.71547652b82fep+7"), H("0x1.8p+52")
EXP_NEGLN2HIN, EXP_NEGLN2LON = H("-0x1.62e42fefa0000p-8"), H("-0x1.cf79abc9e3b3ap-47")
EXP_C = [H("0x1.ffffffffffdbdp-2"), H("0x1.555555555543cp-3"), H("0x1.55555cf172b91p-5"), H("0x1.1111167a4d017p-7")]
# powf_log2_data: invc of the 16 sub-intervals of [0x1.66p-1, 0x1.66p0), poly A0..A4
POWF_INVC = [H(s) for s in (
    "0x1.661ec79f8f3bep+0", "0x1.571ed4aaf883dp+0", "0x1.49539f0f010bp+0", "0x1.3c995b0b80385p+0",
    "0x1.30d190c8864a5p+0", "0x1.25e227b0b8eap+0", "0x1.1bb4a4a1a343fp+0", "0x1.12358f08ae5bap+0",
    "0x1.0953f419900a7p+0", "0x1p+0", "0x1.e608cfd9a47acp-1", "0x1.ca4b31f026aap-1",
    "0x1.b2036576afce6p-1", "0x1.9c2d163a1aa2dp-1", "0x1.886e6037841edp-1", "0x1.767dcf5534862p-1")]
POWF_A = [H("0x1.27616c9496e0bp-2"), H("-0x1.71969a075c67ap-2"), H("0x1.ec70a6ca7baddp-2"), H("-0x1.7154748bef6c8p-1"),
          H("0x1.71547652ab82bp+0")]
# exp2f_data: shift_scaled = 0x1.8p52 / 32, poly C0..C2
EXP2F_SHIFT = H("0x1.8p+47")
EXP2F_C = [H("0x1.c6af84b912394p-5"), H("0x1.ebfce50fac4f3p-3"), H("0x1.62e42ff0c52d6p-1")]


def bits(x):
    return struct.unpack("<Q", struct.pack("<d", x))[0]


def frombits(b):
    return struct.unpack("<d", struct.pack("<Q", b & 0xFFFFFFFFFFFFFFFF))[0]


def compute():
    import mpmath as mp
    mp.mp.prec = 400

    def z_of(ix):       # the z that pow's argument reduction forms from bit pattern ix
        tmp = ix - POW_OFF
        return frombits(ix - (tmp & 0xFFF0000000000000))

    log_tab = []
    for i in range(N_LOG):
        zlo, zhi = z_of(POW_OFF + (i << 45)), z_of(POW_OFF + ((i + 1) << 45) - 1)
        c = (mp.mpf(zlo) + mp.mpf(zhi)) / 2
        den = 128 if zlo < 1 else 256
        invc = float(mp.nint(den / c)) / den
        lc = mp.log(1 / mp.mpf(invc))
        logc = float(mp.nint(lc * 2 ** 43) / 2 ** 43)
        log_tab.append((invc, logc, float(lc - mp.mpf(logc))))
    exp_tab = []
    for i in range(N_EXP):
        v = mp.mpf(2) ** (mp.mpf(i) / N_EXP)
        h = float(v)
        exp_tab.append((bits(float((v - mp.mpf(h)) / mp.mpf(h))), bits(h) - (i << 45)))
    f_log = [(c, float(mp.log(1 / mp.mpf(c), 2))) for c in POWF_INVC]
    f_exp = [bits(float(mp.mpf(2) ** (mp.mpf(i) / 32))) - (i << 47) for i in range(32)]
    return log_tab, exp_tab, f_log, f_exp


def find_once(blob, sig, what):
    i = blob.find(sig)
    if i < 0 or blob.find(sig, i + 1) >= 0:
        raise LookupError("cannot locate %s uniquely in libm" % what)
    return i


def check_against_libm(path, log_tab, exp_tab, f_log, f_exp):
    """The same numbers as they sit inside a glibc >= 2.28 libm (located by signature).  Returns a list of
    differences (empty = identical); raises LookupError when the file is not such a libm."""
    blob = open(path, "rb").read()
    sig = struct.pack("<QQ", bits(POW_LN2HI), bits(POW_LN2LO))
    first = struct.pack("<dd", H("0x1.6ap+0"), 0.0)
    hits, i = [], blob.find(sig)
    while i >= 0:
        if blob[i + 72:i + 88] == first:
            hits.append(i)
        i = blob.find(sig, i + 1)
    if len(hits) != 1:
        raise LookupError("cannot locate pow_log_data uniquely in libm (%d candidates)" % len(hits))
    i = hits[0]
    diffs = []
    head = struct.unpack_from("<9d", blob, i)
    if list(head) != [POW_LN2HI, POW_LN2LO] + POW_A:
        diffs.append("pow_log_data head")
    tab = struct.unpack_from("<%dd" % (4 * N_LOG), blob, i + 72)
    for k in range(N_LOG):
        if (tab[4 * k], tab[4 * k + 2], tab[4 * k + 3]) != log_tab[k]:
            diffs.append("pow log table entry %d" % k)
    j = find_once(blob, struct.pack("<QQ", bits(EXP_INVLN2N), bits(EXP_SHIFT)), "exp_data")
    if list(struct.unpack_from("<8d", blob, j)) != [EXP_INVLN2N, EXP_SHIFT, EXP_NEGLN2HIN, EXP_NEGLN2LON] + EXP_C:
        diffs.append("exp_data head")
    k = blob.find(struct.pack("<QQ", 0, 0x3FF0000000000000), j + 64, j + 64 + 512)
    if k < 0 or (k - j) % 8:
        raise LookupError("cannot locate exp_data.tab")
    etab = struct.unpack_from("<%dQ" % (2 * N_EXP), blob, k)
    for q in range(N_EXP):
        if (etab[2 * q], etab[2 * q + 1]) != exp_tab[q]:
            diffs.append("exp table entry %d" % q)
    fsig = struct.pack("<dd", f_log[0][0], f_log[0][1])
    hits, fi = [], blob.find(fsig)
    while fi >= 0:
        if struct.unpack_from("<d", blob, fi + 256 + 32)[0] == POWF_A[4]:
            hits.append(fi)
        fi = blob.find(fsig, fi + 1)
    if len(hits) != 1:
        raise LookupError("cannot locate powf_log2_data uniquely in libm (%d candidates)" % len(hits))
    fi = hits[0]
    ftab = struct.unpack_from("<32d", blob, fi)
    if [(ftab[2 * q], ftab[2 * q + 1]) for q in range(16)] != f_log:
        diffs.append("powf log2 table")
    if list(struct.unpack_from("<5d", blob, fi + 256)) != POWF_A:
        diffs.append("powf log2 poly")
    fj = find_once(blob, struct.pack("<QQ", f_exp[0], f_exp[1]), "exp2f_data")
    if list(struct.unpack_from("<32Q", blob, fj)) != f_exp:
        diffs.append("exp2f table")
    if list(struct.unpack_from("<4d", blob, fj + 256)) != [EXP2F_SHIFT] + EXP2F_C:
        diffs.append("exp2f head")
    return diffs


def layout(log_tab, exp_tab, f_log, f_exp):
    """[(enum name, [doubles])] for k_gpow and k_gpowf; every block starts at an even index."""
    gp = [("kGP_ln2", [POW_LN2HI, POW_LN2LO]),
          ("kGP_A0", [POW_A[0], 1.0]),                       # A0, one
          ("kGP_A12", [POW_A[1], POW_A[2]]), ("kGP_A34", [POW_A[3], POW_A[4]]), ("kGP_A56", [POW_A[5], POW_A[6]]),
          ("kGP_invln2N", [EXP_INVLN2N, EXP_SHIFT]), ("kGP_negln2", [EXP_NEGLN2HIN, EXP_NEGLN2LON]),
          ("kGP_C23", [EXP_C[0], EXP_C[1]]), ("kGP_C45", [EXP_C[2], EXP_C[3]]),
          ("kGP_log", [v for e in log_tab for v in e[:2]]),           # {invc, logc} x 128
          ("kGP_logtail", [e[2] for e in log_tab]),                  # logctail x 128
          ("kGP_exp", [frombits(b) for e in exp_tab for b in e])]    # {tail, sbits} x 128
    gf = [("kGF_A01", [POWF_A[0], POWF_A[1]]), ("kGF_A23", [POWF_A[2], POWF_A[3]]), ("kGF_A4", [POWF_A[4], 1.0]),
          ("kGF_shift", [EXP2F_SHIFT, EXP2F_C[2]]),                   # shift, C2
          ("kGF_C01", [EXP2F_C[0], EXP2F_C[1]]),
          ("kGF_log2", [v for e in f_log for v in e]),               # {invc, log2 c} x 16
          ("kGF_exp2", [frombits(b) for b in f_exp])]                # sbits x 32
    return gp, gf


def emit(name, enum, count_name, blocks, o):
    o.append("enum %s {" % enum)
    idx = 0
    for n, vals in blocks:
        assert idx % 2 == 0
        o.append("  %s = %d,   // %d entries" % (n, idx, len(vals)))
        idx += len(vals) + (len(vals) & 1)
    o.append("  %s = %d" % (count_name, idx))
    o.append("};")
    o.append("static __device__ const double __align__(16) %s[%d] = {" % (name, idx))
    for n, vals in blocks:
        o.append("    // %s" % n)
        vals = list(vals) + ([0.0] if len(vals) & 1 else [])
        for k in range(0, len(vals), 2):
            assert all(v == v and abs(v) != float("inf") and (v == 0.0 or abs(v) >= 2.3e-308) for v in vals[k:k + 2])
            o.append("    %s, %s,   // %s %s" % (repr(vals[k]), repr(vals[k + 1]), vals[k].hex(), vals[k + 1].hex()))
    o.append("};")


def main():
    args = [a for a in sys.argv[1:] if not a.startswith("--")]
    tabs = compute()
    if "--no-libm-check" not in sys.argv:
        path = args[0] if args else None
        if path is None:
            for cand in ("/lib/x86_64-linux-gnu/libm.so.6", "/usr/lib64/libm.so.6", "/lib64/libm.so.6"):
                if os.path.exists(cand):
                    path = cand
                    break
        if path is not None:
            try:
                diffs = check_against_libm(path, *tabs)
            except LookupError as e:
                print("note: %s -- not a glibc >= 2.28 libm, nothing to compare with" % e)
            else:
                if diffs:
                    raise SystemExit("generated tables differ from %s: %s" % (path, ", ".join(diffs[:8])))
                print("generated tables are identical to the ones inside", path)
    gp, gf = layout(*tabs)
    o = ["// GENERATED by tools/gen_glibc_pow_tables.py -- do not edit.  Tables and coefficients of the table-driven",
         "// pow / powf of ARM optimized-routines (math/pow.c, math/powf.c; MIT OR Apache-2.0 WITH LLVM-exception) as",
         "// adopted by glibc >= 2.28 -- the oracle's libm -- for the literal step-size controller (sde_common.cuh).",
         "// Tables are computed from their definitions (mpmath); polynomial coefficients are the published ones.",
         "// Integer table entries are stored as the double with the same bit pattern.",
         "#pragma once", "namespace sde {"]
    emit("k_gpow", "GpIdx", "kGP_count", gp, o)
    emit("k_gpowf", "GfIdx", "kGF_count", gf, o)
    o.append("}  // namespace sde")
    open(OUT, "w").write("\n".join(o) + "\n")
    print("wrote", OUT)


if __name__ == "__main__":
    main()
