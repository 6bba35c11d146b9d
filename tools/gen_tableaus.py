#!/usr/bin/env python3
"""Generate the coefficient headers and the straight-line stage code.

Inputs (read-only, only needed when REGENERATING; the outputs are committed):
  /root/reference/src/tsit5/atsit5_cache.jl      (Tsit5 c/a/btilde/r literals)
  /root/reference/src/verner/verner_tableaus.jl  (Vern7/Vern9 literals)

Outputs:
  oracle/tableau_named.hpp
      named scalar coefficients for the CPU oracle (names = the reference's names,
      so the oracle can cite the reference line by line).
  simplediffeq.jl_b200/csrc/device/sde_tableaus_gen.cuh
      the same numbers as `__constant__` structs for the sm_100a kernels
      (FP64 and FP32; FP32 = one rounding of the FP64 literal, which is what the
      reference's `convert(Float32, <Float64 literal>)` does).
  simplediffeq.jl_b200/csrc/device/sde_methods_gen.cuh
      straight-line device code of each Runge-Kutta method, generated from the
      STRUCTURE SPEC below (which stage uses which coefficient/stage vector, in which
      order).  The spec is ours; `tools/check_spec_vs_reference.py` re-derives it from
      the reference's solve bodies and fails on any difference.

Only numbers are taken from the reference: every literal is parsed to a double and
re-emitted as a 17-significant-digit decimal (round-trip exact).

FMA placement follows the reference's `@muladd` (MuladdMacro) rewriting, SURVEY.md §8a:
  x + a*b + c*d      -> fma(c, d, fma(a, b, x))
  a*b + c*d + e*f    -> fma(e, f, fma(c, d, a*b))
  uprev + dt*(sum)   -> fma(dt, sum, uprev)
  uprev + dt*a21*k1  -> fma(dt*a21, k1, uprev)
  t + c*dt           -> fma(c, dt, t)
  @evalpoly          -> Horner with fma
"""
import os
import re
import sys

REF = os.environ.get("SDE_REFERENCE", "/root/reference")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "simplediffeq.jl_b200")


def lit(x: float) -> str:
    s = "%.17g" % x
    if "e" not in s and "." not in s and "n" not in s and "i" not in s:
        s += ".0"
    return s


# --------------------------------------------------------------------------------------
# 1. parse numbers
# --------------------------------------------------------------------------------------
def parse_tsit5():
    src = open(os.path.join(REF, "src/tsit5/atsit5_cache.jl")).read()
    # cs = SVector{6, T}(0.161, 0.327, ...)
    m = re.search(r"cs = SVector\{6, T\}\(([^)]*)\)", src)
    cs = [float(v) for v in m.group(1).split(",")]
    assert len(cs) == 6

    def block(name, n):
        m = re.search(name + r" = SVector\{%d, T\}\((.*?)\n    \)" % n, src, re.S)
        vals = re.findall(r"convert\(T, ([-0-9.e]+)\)", m.group(1))
        assert len(vals) == n, (name, len(vals))
        return [float(v) for v in vals]

    a = block("as", 21)
    bt = block("btildes", 7)
    r = block("rs", 22)
    names_a = ["a21", "a31", "a32", "a41", "a42", "a43", "a51", "a52", "a53", "a54",
               "a61", "a62", "a63", "a64", "a65", "a71", "a72", "a73", "a74", "a75", "a76"]
    names_r = ["r11", "r12", "r13", "r14", "r22", "r23", "r24", "r32", "r33", "r34", "r42",
               "r43", "r44", "r52", "r53", "r54", "r62", "r63", "r64", "r72", "r73", "r74"]
    out = []
    out += [("c%d" % (i + 1), cs[i]) for i in range(6)]
    out += list(zip(names_a, a))
    out += [("btilde%d" % (i + 1), bt[i]) for i in range(7)]
    out += list(zip(names_r, r))
    return out


def parse_verner():
    src = open(os.path.join(REF, "src/verner/verner_tableaus.jl")).read()
    res = {}
    for fn in ["Vern7ExtraStages", "Vern7InterpolationCoefficients", "Vern7Tableau",
               "Vern9ExtraStages", "Vern9InterpolationCoefficients", "Vern9Tableau"]:
        m = re.search(r"^function " + fn + r"\(.*?^end", src, re.S | re.M)
        body = m.group(0)
        vals = re.findall(r"^\s+(\w+) = convert\(T2?, ([-0-9.e]+)\)\s*$", body, re.M)
        res[fn] = [(n, float(v)) for n, v in vals]
    assert [len(res[k]) for k in res] == [60, 79, 58, 135, 153, 108], [len(res[k]) for k in res]
    return res


# --------------------------------------------------------------------------------------
# 2. structure spec (ours; verified against the reference by tools/check_spec_vs_reference.py)
#    a stage = (stage_number, time, [(coef_name, k_number), ...])
#    time = coefficient name for `t + c*dt`, or "1" for `t + dt`
# --------------------------------------------------------------------------------------
def spec_tsit5():
    st = []
    cnames = {2: "c1", 3: "c2", 4: "c3", 5: "c4", 6: "1"}
    for s in range(2, 7):
        st.append((s, cnames[s], [("a%d%d" % (s, j), j) for j in range(1, s)]))
    upd = [("a7%d" % j, j) for j in range(1, 7)]
    err = [("btilde%d" % j, j) for j in range(1, 8)]
    polys = [(1, [None, "r11", "r12", "r13", "r14"])]
    for j in range(2, 8):
        polys.append((j, [None, None, "r%d2" % j, "r%d3" % j, "r%d4" % j]))
    return dict(name="Tsit5", stages=st, update=upd, err=err, polys=polys, extra=[],
                nk=7, fsal=True)


def _terms_from_names(names, stage, width_j):
    """coefficients a<stage><j> in struct-field order -> [(name, j)]"""
    out = []
    for n in names:
        if not n.startswith("a"):
            continue
        digits = n[1:]
        s, j = int(digits[:2]), int(digits[2:])
        assert len(digits[2:]) == width_j
        if s == stage:
            out.append((n, j))
    return out


def spec_vern7(v):
    main = [n for n, _ in v["Vern7Tableau"]]
    extra = [n for n, _ in v["Vern7ExtraStages"]]
    st = []
    for s in range(2, 11):
        time = "c%d" % s if s <= 8 else "1"
        st.append((s, time, _terms_from_names(main, s, 1)))
    upd = [("b%d" % j, j) for j in (1, 4, 5, 6, 7, 8, 9)]
    err = [("btilde%d" % j, j) for j in (1, 4, 5, 6, 7, 8, 9, 10)]
    ex = []
    for s in range(11, 17):
        ex.append((s, "c%d" % s, _terms_from_names(extra, s, 2)))
    polys = [(1, [None] + ["r01%d" % d for d in range(1, 8)])]
    for j in (4, 5, 6, 7, 8, 9, 11, 12, 13, 14, 15, 16):
        polys.append((j, [None, None] + ["r%02d%d" % (j, d) for d in range(2, 8)]))
    return dict(name="Vern7", stages=st, update=upd, err=err, polys=polys, extra=ex,
                nk=16, fsal=False)


def spec_vern9(v):
    main = [n for n, _ in v["Vern9Tableau"]]
    extra = [n for n, _ in v["Vern9ExtraStages"]]
    st = []
    for s in range(2, 17):
        time = "c%d" % (s - 1) if s <= 14 else "1"
        st.append((s, time, _terms_from_names(main, s, 2)))
    upd = [("b%d" % j, j) for j in (1, 8, 9, 10, 11, 12, 13, 14, 15)]
    err = [("btilde%d" % j, j) for j in (1, 8, 9, 10, 11, 12, 13, 14, 15, 16)]
    ex = []
    for s in range(17, 27):
        ex.append((s, "c%d" % s, _terms_from_names(extra, s, 2)))
    polys = [(1, [None] + ["r01%d" % d for d in range(1, 10)])]
    for j in list(range(8, 16)) + list(range(17, 27)):
        polys.append((j, [None, None] + ["r%02d%d" % (j, d) for d in range(2, 10)]))
    return dict(name="Vern9", stages=st, update=upd, err=err, polys=polys, extra=ex,
                nk=26, fsal=False)


# --------------------------------------------------------------------------------------
# 3. emitters
# --------------------------------------------------------------------------------------
HDR = "// GENERATED by tools/gen_tableaus.py -- do not edit by hand.\n"


def emit_oracle(ts, v):
    o = [HDR]
    o.append("// Named Runge-Kutta coefficients for the CPU oracle. Numbers parsed from\n"
             "//   src/tsit5/atsit5_cache.jl:1-105 and src/verner/verner_tableaus.jl:65-135,\n"
             "//   219-311,376-452,595-751,909-1083,1201-1335 of SciML/SimpleDiffEq.jl v1.16.3;\n"
             "// each is `convert(T, <Float64 literal>)` there, i.e. T(double) here.\n"
             "#pragma once\nnamespace oracle_tab {\n")

    def struct(name, items):
        o.append("template <class T> struct %s {\n" % name)
        for n, x in items:
            o.append("  static constexpr T %s = T(%s);\n" % (n, lit(x)))
        o.append("};\n\n")

    struct("Tsit5Tab", ts)
    struct("Vern7Tab", v["Vern7Tableau"] + v["Vern7ExtraStages"] + v["Vern7InterpolationCoefficients"])
    struct("Vern9Tab", v["Vern9Tableau"] + v["Vern9ExtraStages"] + v["Vern9InterpolationCoefficients"])
    o.append("}  // namespace oracle_tab\n")
    return "".join(o)


def emit_device_tables(ts, v):
    o = [HDR]
    o.append("// Butcher tableaus / dense-output coefficients in __constant__ memory (sm_100a).\n"
             "// Values: see tools/gen_tableaus.py. FP32 = (float) of the FP64 literal.\n"
             "#pragma once\nnamespace sde {\n\n")
    groups = [
        ("Tsit5Coef", "tsit5", ts),
        ("Vern7Coef", "vern7", v["Vern7Tableau"] + v["Vern7ExtraStages"] + v["Vern7InterpolationCoefficients"]),
        ("Vern9Coef", "vern9", v["Vern9Tableau"] + v["Vern9ExtraStages"] + v["Vern9InterpolationCoefficients"]),
    ]
    for sname, low, items in groups:
        o.append("template <class T> struct %s {\n" % sname)
        names = [n for n, _ in items]
        for i in range(0, len(names), 8):
            o.append("  T " + ", ".join(names[i:i + 8]) + ";\n")
        o.append("};\n")
        for ty, suf in (("double", "f64"), ("float", "f32")):
            o.append("static __constant__ %s<%s> k_%s_%s = {\n" % (sname, ty, low, suf))
            for n, x in items:
                if ty == "double":
                    o.append("  /*%s*/ %s,\n" % (n, lit(x)))
                else:
                    o.append("  /*%s*/ (float)%s,\n" % (n, lit(x)))
            o.append("};\n")
        o.append("\n")
    o.append("template <class T> struct Coefs;\n")
    for ty, suf in (("double", "f64"), ("float", "f32")):
        o.append("template <> struct Coefs<%s> {\n" % ty)
        for sname, low, _ in groups:
            o.append("  static __device__ __forceinline__ const %s<%s>& %s() { return k_%s_%s; }\n"
                     % (sname, ty, low, low, suf))
        o.append("};\n")
    o.append("\n}  // namespace sde\n")
    return "".join(o)


def fold(terms, kname=lambda j: "k%d[i]" % j, coef=lambda n: "C." + n):
    """left fold of products: a1*k1 + a2*k2 + ... -> fma(an,kn, ... fma(a2,k2, a1*k1))"""
    n0, j0 = terms[0]
    e = "%s * %s" % (coef(n0), kname(j0))
    for n, j in terms[1:]:
        e = "fma(%s, %s, %s)" % (coef(n), kname(j), e)
    return e


def horner(coefs, theta="th"):
    """@evalpoly(th, c0, c1, ..., cn) -> Horner with fma; None = literal zero."""
    def c(x):
        return "T(0)" if x is None else "C." + x
    e = c(coefs[-1])
    for x in reversed(coefs[:-1]):
        e = "fma(%s, %s, %s)" % (theta, e, c(x))
    return e


def emit_method(sp, refnote):
    name = sp["name"]
    low = name.lower()
    nk = sp["nk"]
    o = []
    o.append("// %s\n" % refnote)
    o.append("template <class Sys, class T>\nstruct %sMethod {\n" % name)
    o.append("  static constexpr int N = Sys::N;\n")
    o.append("  static constexpr bool kFSAL = %s;\n" % ("true" if sp["fsal"] else "false"))
    o.append("  static constexpr bool kHasExtra = %s;\n" % ("true" if sp["extra"] else "false"))
    used = sorted({1} | {s for s, _, _ in sp["stages"]} | {s for s, _, _ in sp["extra"]} |
                  ({7} if sp["fsal"] else set()))
    o.append("  " + " ".join("T k%d[N];" % j for j in used) + "\n\n")
    if sp["fsal"]:
        o.append("  // FSAL seed: k7 = f(u0, p, t0)\n")
        o.append("  __device__ __forceinline__ void seed(const T* u, const T* p, T t) { Sys::rhs(k7, u, p, t); }\n")
        o.append("  // k1 = k7 at the start of every (accepted) step\n")
        o.append("  __device__ __forceinline__ void begin_step() {\n"
                 "#pragma unroll\n    for (int i = 0; i < N; ++i) k1[i] = k7[i];\n  }\n")
    else:
        o.append("  __device__ __forceinline__ void seed(const T*, const T*, T) {}\n")
        o.append("  __device__ __forceinline__ void begin_step() {}\n")
    # ---- stages
    o.append("\n  // one attempt: all stages + the new state u. kNeedErrStage: also evaluate the\n"
             "  // stage that only the embedded error estimate uses (adaptive methods).\n")
    o.append("  template <bool kNeedErrStage>\n")
    o.append("  __device__ __forceinline__ void stages(const T* uprev, T* u, const T* p, T t, T dt) {\n")
    o.append("    const %sCoef<T>& C = Coefs<T>::%s();\n    T tmp[N];\n" % (name, low))
    if not sp["fsal"]:
        o.append("    Sys::rhs(k1, uprev, p, t);\n")
    last_main = sp["stages"][-1][0]
    upd_emitted = False

    def emit_update():
        o.append("#pragma unroll\n    for (int i = 0; i < N; ++i)\n      u[i] = fma(dt, %s, uprev[i]);\n"
                 % fold(sp["update"]))

    for s, time, terms in sp["stages"]:
        tt = "t + dt" if time == "1" else "fma(C.%s, dt, t)" % time
        err_only = (not sp["fsal"]) and s == last_main
        ind = "    "
        if err_only:
            # update first in the reference? no: reference computes g9,g10,k9,k10 then u. Values are
            # independent of the order, we keep the reference's order of evaluation.
            o.append("    if (kNeedErrStage) {\n")
            ind = "      "
        if len(terms) == 1:
            n, j = terms[0]
            o.append("%s{\n%s  const T a = dt * C.%s;\n#pragma unroll\n%s  for (int i = 0; i < N; ++i) tmp[i] = fma(a, k%d[i], uprev[i]);\n%s}\n"
                     % (ind, ind, n, ind, j, ind))
        else:
            o.append("#pragma unroll\n%sfor (int i = 0; i < N; ++i)\n%s  tmp[i] = fma(dt, %s, uprev[i]);\n"
                     % (ind, ind, fold(terms)))
        o.append("%sSys::rhs(k%d, tmp, p, %s);\n" % (ind, s, tt))
        if err_only:
            o.append("    }\n")
    emit_update()
    if sp["fsal"]:
        o.append("    Sys::rhs(k7, u, p, t + dt);\n")
    o.append("  }\n")
    # ---- error
    o.append("\n  // embedded error estimate: e = dt * (btilde . k).  bt = the btilde coefficients in the order of\n"
             "  // load_btilde(); the adaptive kernels keep them in shared memory (uniform-register relief).\n")
    o.append("  static constexpr int kNBT = %d;\n" % len(sp["err"]))
    o.append("  __device__ __forceinline__ static void load_btilde(T* bt) {\n")
    o.append("    const %sCoef<T>& C = Coefs<T>::%s();\n" % (name, low))
    for idx, (n, j) in enumerate(sp["err"]):
        o.append("    bt[%d] = C.%s;\n" % (idx, n))
    o.append("  }\n")
    o.append("  __device__ __forceinline__ void error(T dt, T* e, const T* bt) const {\n")
    bt_terms = [("bt[%d]" % idx, j) for idx, (n, j) in enumerate(sp["err"])]
    o.append("#pragma unroll\n    for (int i = 0; i < N; ++i)\n      e[i] = dt * (%s);\n  }\n" % fold(bt_terms, coef=lambda n: n))
    # ---- extra stages
    o.append("\n  // extra stages needed by the dense output (none for Tsit5). tx = time base the\n"
             "  // reference passes to f (quirk Q3), kQ2: reference-exact fixed-step Vern9 pairs the\n"
             "  // stage-8..15 coefficients with k2..k9 (quirk Q2).\n")
    o.append("  template <bool kQ2>\n")
    o.append("  __device__ __forceinline__ void dense_prepare(const T* uprev, const T* p, T tx, T dt) {\n")
    if sp["extra"]:
        o.append("    const %sCoef<T>& C = Coefs<T>::%s();\n    T tmp[N];\n" % (name, low))
        for s, time, terms in sp["extra"]:
            def kn(j, q2):
                if q2 and name == "Vern9" and 8 <= j <= 15:
                    return "k%d[i]" % (j - 6)
                return "k%d[i]" % j
            if name == "Vern9":
                o.append("#pragma unroll\n    for (int i = 0; i < N; ++i)\n"
                         "      tmp[i] = kQ2 ? fma(dt, %s, uprev[i])\n"
                         "                   : fma(dt, %s, uprev[i]);\n"
                         % (fold(terms, kname=lambda j: kn(j, True)), fold(terms, kname=lambda j: kn(j, False))))
            else:
                o.append("#pragma unroll\n    for (int i = 0; i < N; ++i)\n      tmp[i] = fma(dt, %s, uprev[i]);\n"
                         % fold(terms))
            o.append("    Sys::rhs(k%d, tmp, p, fma(C.%s, dt, tx));\n" % (s, time))
    o.append("  }\n")
    # ---- dense output
    nb = len(sp["polys"])
    o.append("\n  // dense-output weights b_j(theta) (Horner with fma == @evalpoly), in the reference's order\n")
    o.append("  static constexpr int kNB = %d;\n" % nb)
    o.append("  __device__ __forceinline__ static void bthetas(T th, T* b) {\n")
    o.append("    const %sCoef<T>& C = Coefs<T>::%s();\n" % (name, low))
    for idx, (j, coefs) in enumerate(sp["polys"]):
        o.append("    b[%d] = %s;\n" % (idx, horner(coefs)))
    o.append("  }\n")
    o.append("\n  // dense output: out = uprev + dt * sum_j b_j k_j   (b from bthetas(); fixed-step kernels get\n"
             "  // them precomputed by the host because theta is the same for every trajectory)\n")
    o.append("  template <bool kQ2>\n")
    o.append("  __device__ __forceinline__ void dense_combine(const T* b, T dt, const T* uprev, T* out) const {\n")
    terms = [("b[%d]" % idx, j) for idx, (j, _) in enumerate(sp["polys"])]

    def kn2(j, q2):
        if q2 and name == "Vern9" and 8 <= j <= 15:
            return "k%d[i]" % (j - 6)
        return "k%d[i]" % j
    if name == "Vern9":
        o.append("#pragma unroll\n    for (int i = 0; i < N; ++i)\n"
                 "      out[i] = kQ2 ? fma(dt, %s, uprev[i])\n"
                 "                   : fma(dt, %s, uprev[i]);\n"
                 % (fold(terms, kname=lambda j: kn2(j, True), coef=lambda n: n),
                    fold(terms, kname=lambda j: kn2(j, False), coef=lambda n: n)))
    else:
        o.append("#pragma unroll\n    for (int i = 0; i < N; ++i)\n      out[i] = fma(dt, %s, uprev[i]);\n"
                 % fold(terms, coef=lambda n: n))
    o.append("  }\n")
    o.append("  template <bool kQ2>\n")
    o.append("  __device__ __forceinline__ void dense(T th, T dt, const T* uprev, T* out) const {\n"
             "    T b[kNB];\n    bthetas(th, b);\n    dense_combine<kQ2>(b, dt, uprev, out);\n  }\n")
    o.append("};\n\n")
    return "".join(o)


def emit_methods(specs):
    o = [HDR]
    o.append("// Straight-line device code of the Runge-Kutta methods (one trajectory per thread,\n"
             "// stage vectors k* register-resident). FMA placement = the reference's @muladd\n"
             "// rewriting (see tools/gen_tableaus.py docstring).\n"
             "#pragma once\n#include \"sde_tableaus_gen.cuh\"\nnamespace sde {\n\n")
    notes = {
        "Tsit5": "Tsitouras 5(4): src/tsit5/gpuatsit5.jl:96-110 (stages), :272-275 (error), :118-125 + src/tsit5/tsit5.jl:385-399 (dense)",
        "Vern7": "Verner 7(6): src/verner/gpuvern7.jl:104-136 (stages), :397-402 (error), :153-219 (extra stages + dense), verner_tableaus.jl:1337-1360",
        "Vern9": "Verner 9(8): src/verner/gpuvern9.jl:102-183 (stages), :556-560 (error), :216-331 / :631-757 (extra stages + dense), verner_tableaus.jl:1362-1398",
    }
    for sp in specs:
        o.append(emit_method(sp, notes[sp["name"]]))
    o.append("}  // namespace sde\n")
    return "".join(o)


def emit_host_interp(ts, v, specs):
    """Plain C++ (host) copy of the dense-output polynomials: the launcher precomputes b_j(theta) of
    every save point of a fixed-step solve (theta does not depend on the trajectory)."""
    vals = {"Tsit5": dict(ts),
            "Vern7": dict(v["Vern7InterpolationCoefficients"]),
            "Vern9": dict(v["Vern9InterpolationCoefficients"])}
    o = [HDR]
    o.append("// Host copy of the dense-output polynomial coefficients (ascending powers of theta, zero-padded).\n"
             "// b_j(theta) = Horner with fma, exactly like the device code / the reference's @evalpoly.\n"
             "#pragma once\nnamespace sde_host {\n\n")
    for sp in specs:
        name = sp["name"]
        deg = max(len(c) for _, c in sp["polys"])
        o.append("static const int k%sNB = %d, k%sDeg = %d;\n" % (name, len(sp["polys"]), name, deg))
        o.append("static const double k%sPoly[%d][%d] = {\n" % (name, len(sp["polys"]), deg))
        for j, coefs in sp["polys"]:
            row = [0.0 if c is None else vals[name][c] for c in coefs] + [0.0] * (deg - len(coefs))
            o.append("  {" + ", ".join(lit(x) for x in row) + "},\n")
        o.append("};\n")
        o.append("static const int k%sLen[%d] = {%s};\n\n" % (name, len(sp["polys"]), ", ".join(str(len(c)) for _, c in sp["polys"])))
    # SDE_COMPAT_FAST_STAGES (sde_kernels.cuh: Tsit5FastMethod): the launcher folds the step size into these
    names = ["a%d%d" % (s, j) for s in range(2, 8) for j in range(1, s)]
    o.append("// Tsit5 stage coefficients a21, a31, a32, ... a76 (row by row): the launcher hands dt * a_ij to the\n"
             "// SDE_COMPAT_FAST_STAGES kernels through KArgs::hcoef\n")
    o.append("static const int kTsit5NStageCoef = %d;\n" % len(names))
    o.append("static const double kTsit5StageCoef[%d] = {\n" % len(names))
    for nm in names:
        o.append("  /*%s*/ %s,\n" % (nm, lit(vals["Tsit5"][nm])))
    o.append("};\n\n")
    o.append("}  // namespace sde_host\n")
    return "".join(o)


def main():
    ts = parse_tsit5()
    v = parse_verner()
    specs = [spec_tsit5(), spec_vern7(v), spec_vern9(v)]
    outs = {
        os.path.join(ROOT, "oracle/tableau_named.hpp"): emit_oracle(ts, v),
        os.path.join(PKG, "csrc/device/sde_tableaus_gen.cuh"): emit_device_tables(ts, v),
        os.path.join(PKG, "csrc/device/sde_methods_gen.cuh"): emit_methods(specs),
        os.path.join(PKG, "csrc/sde_interp_host_gen.h"): emit_host_interp(ts, v, specs),
    }
    for path, text in outs.items():
        os.makedirs(os.path.dirname(path), exist_ok=True)
        with open(path, "w") as f:
            f.write(text)
        print("wrote", os.path.relpath(path, ROOT), len(text.splitlines()), "lines")


if __name__ == "__main__":
    if not os.path.isdir(REF):
        sys.exit("reference tree not found at %s (outputs are committed; nothing to do)" % REF)
    main()
