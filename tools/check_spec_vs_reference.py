#!/usr/bin/env python3
"""Check the STRUCTURE SPEC of tools/gen_tableaus.py (which stage combines which coefficients and
stage vectors, in which order) against the reference's solve bodies.

Runs only where the reference tree is present (the build container); exits 0 with a notice
otherwise.  Method: strip all whitespace from the reference source, build from our spec the exact
text each sum must have in the reference's naming, and require (a) every expected sum occurs,
(b) every `coef*k + coef*k + ...` sum that occurs in the reference is one we expected.
"""
import os
import re
import sys

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import gen_tableaus as G  # noqa: E402

REF = G.REF
SUM_RE = re.compile(r"(?:\w+\*k\d+|k\d+\*\w+)(?:\+(?:\w+\*k\d+|k\d+\*\w+))+")


def stripped(path):
    return re.sub(r"\s+", "", open(os.path.join(REF, path)).read())


def sums(terms, kname=lambda j: "k%d" % j, swap=False):
    if swap:
        return "+".join("%s*%s" % (kname(j), n) for n, j in terms)
    return "+".join("%s*%s" % (n, kname(j)) for n, j in terms)


def check(path, expected, singles):
    src = stripped(path)
    found = SUM_RE.findall(src)
    ok = True
    for e in expected:
        if e not in src:
            print("MISSING in %s: %s" % (path, e))
            ok = False
    exp = set(expected)
    for f in found:
        if f not in exp:
            print("UNEXPECTED sum in %s: %s" % (path, f))
            ok = False
    for s in singles:
        if s not in src:
            print("MISSING in %s: %s" % (path, s))
            ok = False
    print("%-28s %3d sums in the reference, %3d distinct expected: %s" % (path, len(found), len(exp), "ok" if ok else "MISMATCH"))
    return ok


def main():
    if not os.path.isdir(REF):
        print("reference tree not present; nothing to check")
        return 0
    v = G.parse_verner()
    ok = True
    # ---- Tsit5
    sp = G.spec_tsit5()
    exp = [sums(t) for _, _, t in sp["stages"] if len(t) > 1] + [sums(sp["update"]), sums(sp["err"])]
    exp.append("+".join("b%dθ*k%d" % (j, j) for j in range(1, 8)))
    singles = ["tmp=uprev+dt*a21*k1", "k2=f(tmp,p,t+c1*dt)", "k3=f(tmp,p,t+c2*dt)", "k4=f(tmp,p,t+c3*dt)",
               "k5=f(tmp,p,t+c4*dt)", "k6=f(tmp,p,t+dt)", "k7=f(u,p,t+dt)"]
    ok &= check("src/tsit5/gpuatsit5.jl", exp, singles)
    # ---- RK4
    src = stripped("src/rk4/gpurk4.jl")
    for s in ["t=ts[i]", "k1=f(u,p,t)", "tmp=uprev+dt*half*k1", "k2=f(tmp,p,t+half*dt)", "tmp=uprev+dt*half*k2",
              "k3=f(tmp,p,t+half*dt)", "tmp=uprev+dt*k3", "k4=f(tmp,p,t+dt)", "u=uprev+dt*sixth*(k1+2k2+2k3+k4)"]:
        if s not in src:
            print("MISSING in gpurk4.jl:", s)
            ok = False
    print("%-28s literal statements: %s" % ("src/rk4/gpurk4.jl", "ok" if ok else "MISMATCH"))
    # ---- Euler
    src = stripped("src/euler/gpueuler.jl")
    for s_ in ["t=ts[i]", "k1=f(u,p,t)", "u=uprev+dt*k1", "us[i]=u"]:
        if s_ not in src:
            print("MISSING in gpueuler.jl:", s_)
            ok = False
    print("%-28s literal statements: %s" % ("src/euler/gpueuler.jl", "ok" if ok else "MISMATCH"))
    # ---- Vern7
    sp = G.spec_vern7(v)
    exp = [sums(t) for _, _, t in sp["stages"] if len(t) > 1] + [sums(sp["update"]), sums(sp["err"])]
    exp += [sums(t) for _, _, t in sp["extra"]]
    exp.append("+".join("k%d*b%dΘ" % (j, j) for j, _ in sp["polys"]))
    singles = ["a=dt*a021", "k2=f(uprev+a*k1,p,t+c2*dt)"]
    for s, time, _ in sp["stages"][1:7]:
        singles.append("p,t+%s*dt)" % time)
    for s, time, _ in sp["extra"]:
        singles += ["t+%s*dt" % time, "t+%s*dtold" % time]
    ok &= check("src/verner/gpuvern7.jl", exp, singles)
    # ---- Vern9 (extra stages and dense output are written with the renamed variables k2..k9, k11..k20)
    sp = G.spec_vern9(v)
    exp = [sums(t) for _, _, t in sp["stages"] if len(t) > 1] + [sums(sp["update"]), sums(sp["err"])]

    def ren(j):
        if 8 <= j <= 15:
            return "k%d" % (j - 6)
        if j >= 17:
            return "k%d" % (j - 6)
        return "k%d" % j
    exp += [sums(t, kname=ren) for _, _, t in sp["extra"]]
    exp.append("+".join("%s*b%dΘ" % (ren(j), j) for j, _ in sp["polys"]))
    singles = ["a=dt*a0201", "k2=f(uprev+a*k1,p,t+c1*dt)", "k2=k8", "k3=k9", "k4=k10", "k5=k11", "k6=k12", "k7=k13",
               "k8=k14", "k9=k15", "k10=k16"]
    for s, time, _ in sp["extra"]:
        singles += ["t+%s*dt" % time, "told+%s*dtold" % time]
    ok &= check("src/verner/gpuvern9.jl", exp, singles)
    # ---- interpolation polynomials
    src = stripped("src/verner/verner_tableaus.jl") + stripped("src/tsit5/tsit5.jl")
    for sp in (G.spec_tsit5(), G.spec_vern7(v), G.spec_vern9(v)):
        for j, coefs in sp["polys"]:
            zero = "zero(T)" if sp["name"] == "Tsit5" else "0"
            args = ",".join(zero if c is None else c for c in coefs)
            s = "@evalpoly(θ,%s)" % args
            if s not in src:
                print("MISSING polynomial:", s)
                ok = False
    print("interpolation polynomials: %s" % ("ok" if ok else "MISMATCH"))
    return 0 if ok else 1


if __name__ == "__main__":
    sys.exit(main())
