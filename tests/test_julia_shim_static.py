"""Static ABI check of the Julia shim (julia/SimpleDiffEqCUDA.jl).  Julia is not installed in the build
container, so the shim cannot be executed; the failure class that a never-run binding is most exposed to --
a struct mirror or a `ccall` signature that drifted from include/simplediffeq_cuda.h -- is checked here by
parsing both files: field order / types of the two option structs, and return type, argument count and
argument types of every `ccall`."""
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = open(os.path.join(ROOT, "include", "simplediffeq_cuda.h")).read()
SHIM = open(os.path.join(ROOT, "julia", "SimpleDiffEqCUDA.jl")).read()

_nocomment = re.sub(r"/\*.*?\*/", " ", HEADER, flags=re.S)
_nocomment = re.sub(r"//[^\n]*", " ", _nocomment)

# C type -> Julia types a ccall signature / struct mirror may use for it
C2JL = {
    "int": {"Cint"}, "int32_t": {"Int32", "Cint"}, "int64_t": {"Int64"}, "uint64_t": {"UInt64"},
    "double": {"Float64", "Cdouble"}, "size_t": {"Csize_t"},
    "const char*": {"Cstring"}, "char*": {"Ptr{UInt8}"},
    "void*": {"Ptr{Cvoid}"}, "const void*": {"Ptr{Cvoid}"}, "void**": {"Ref{Ptr{Cvoid}}"},
    "sde_system_t": {"Ptr{Cvoid}"}, "sde_em_system_t": {"Ptr{Cvoid}"},
    "sde_system_t*": {"Ref{Ptr{Cvoid}}"}, "sde_em_system_t*": {"Ref{Ptr{Cvoid}}"},
    "int*": {"Ref{Cint}", "Ptr{Cint}"}, "const int*": {"Ptr{Cint}"},
    "int32_t*": {"Ptr{Int32}"}, "int64_t*": {"Ref{Int64}", "Ptr{Int64}"}, "double*": {"Ref{Float64}", "Ptr{Float64}"},
    "const sde_options_t*": {"Ref{SdeOptions}"}, "const sde_em_options_t*": {"Ref{SdeEmOptions}"},
    "void": {"Cvoid"},
}


def _ctype(decl):
    """'const void* u0' -> 'const void*' (drops the parameter name)."""
    decl = " ".join(decl.replace("*", " * ").split())
    toks = decl.split(" ")
    if toks[-1] != "*" and len(toks) > 1:
        toks = toks[:-1]                      # parameter name
    t = " ".join(toks).replace(" *", "*")
    return t


def header_prototypes():
    protos = {}
    for m in re.finditer(r"SDE_API\s+([\w\s\*]+?)\s*\b(sde_\w+)\s*\(([^)]*)\)\s*;", _nocomment):
        ret, name, args = m.group(1).strip(), m.group(2), m.group(3).strip()
        ret = " ".join(ret.replace("*", " * ").split()).replace(" *", "*")
        params = [] if args in ("void", "") else [_ctype(a) for a in args.split(",")]
        protos[name] = (ret, params)
    return protos


def header_struct(tag):
    m = re.search(r"typedef struct %s \{(.*?)\}" % tag, _nocomment, flags=re.S)
    fields = []
    for stmt in m.group(1).split(";"):
        stmt = " ".join(stmt.split())
        if not stmt:
            continue
        if "*" in stmt:
            ty = " ".join(stmt.rsplit("*", 1)[0].split()) + "*"
            names = [stmt.rsplit("*", 1)[1].strip()]
        else:
            ty, rest = stmt.split(" ", 1)
            names = [n.strip() for n in rest.split(",")]
        fields.extend((n, ty) for n in names)
    return fields


def julia_struct(name):
    m = re.search(r"^struct %s\n(.*?)^end" % name, SHIM, flags=re.S | re.M)
    return [tuple(x.strip() for x in line.split("::")) for line in m.group(1).strip().splitlines()]


def _split_top(s):
    out, depth, cur = [], 0, ""
    for ch in s:
        if ch in "{(":
            depth += 1
        elif ch in "})":
            depth -= 1
        if ch == "," and depth == 0:
            out.append(cur.strip())
            cur = ""
        else:
            cur += ch
    if cur.strip():
        out.append(cur.strip())
    return out


def julia_ccalls():
    calls = []
    for m in re.finditer(r"ccall\(\(:(\w+), libsde\),\s*(\w+),\s*\(", SHIM):
        i, depth = m.end(), 1
        while depth:
            depth += {"(": 1, ")": -1}.get(SHIM[i], 0)
            i += 1
        sig = _split_top(SHIM[m.end():i - 1])
        # actual arguments: up to the ccall's closing parenthesis
        j, depth = i, 1
        while depth:
            depth += {"(": 1, ")": -1, "[": 1, "]": -1}.get(SHIM[j], 0)
            j += 1
        actual = _split_top(SHIM[i:j - 1].lstrip(", \n"))
        calls.append((m.group(1), m.group(2), sig, actual, SHIM[:m.start()].count("\n") + 1))
    return calls


def test_option_structs_mirror_the_header_field_by_field():
    for tag, jl in (("sde_options", "SdeOptions"), ("sde_em_options", "SdeEmOptions")):
        c_fields, j_fields = header_struct(tag), julia_struct(jl)
        assert [n for n, _ in c_fields] == [n for n, _ in j_fields], (tag, "field names / order")
        for (n, cty), (_, jty) in zip(c_fields, j_fields):
            assert jty in C2JL[cty], "%s.%s: C %s vs Julia %s" % (tag, n, cty, jty)


def test_every_ccall_matches_its_prototype():
    protos = header_prototypes()
    assert len(protos) >= 25 and "sde_solve" in protos and "sde_em_solve" in protos
    calls = julia_ccalls()
    assert {c[0] for c in calls} >= {"sde_solve", "sde_fixed_times", "sde_system_builtin", "sde_system_dims",
                                     "sde_system_nvrtc", "sde_last_error", "sde_em_solve", "sde_em_system_builtin",
                                     "sde_em_system_dims", "sde_em_system_nvrtc"}
    for name, ret, sig, actual, line in calls:
        assert name in protos, "shim line %d binds %s, which the header does not declare" % (line, name)
        cret, cparams = protos[name]
        assert ret in C2JL[cret], "%s (line %d): return %s vs C %s" % (name, line, ret, cret)
        assert len(sig) == len(cparams), "%s (line %d): %d argument types for %d parameters" % (name, line, len(sig), len(cparams))
        assert len(actual) == len(cparams), "%s (line %d): %d arguments passed for %d parameters: %r" % (
            name, line, len(actual), len(cparams), actual)
        for k, (jty, cty) in enumerate(zip(sig, cparams)):
            assert jty in C2JL[cty], "%s (line %d) argument %d: Julia %s vs C %s" % (name, line, k + 1, jty, cty)


def test_constants_agree_with_the_header():
    enum = dict((m.group(1), int(m.group(2))) for m in re.finditer(r"\b(SDE_\w+)\s*=\s*(-?\d+)", _nocomment))
    for alg, cname in (("GPUSimpleTsit5", "SDE_ALG_TSIT5"), ("GPUSimpleATsit5", "SDE_ALG_ATSIT5"), ("GPUSimpleRK4", "SDE_ALG_RK4"),
                       ("GPUSimpleVern7", "SDE_ALG_VERN7"), ("GPUSimpleAVern7", "SDE_ALG_AVERN7"),
                       ("GPUSimpleVern9", "SDE_ALG_VERN9"), ("GPUSimpleAVern9", "SDE_ALG_AVERN9"),
                       ("GPUSimpleEuler", "SDE_ALG_EULER")):
        m = re.search(r"alg_id\(::%s\) = Int32\((\d+)\)" % alg, SHIM)
        assert m and int(m.group(1)) == enum[cname], alg
    m = re.search(r"const SDE_SAVE_ENDPOINT, SDE_SAVE_SAVEAT, SDE_SAVE_EVERYSTEP = Int32\((\d)\), Int32\((\d)\), Int32\((\d)\)", SHIM)
    assert [int(x) for x in m.groups()] == [enum["SDE_SAVE_ENDPOINT"], enum["SDE_SAVE_SAVEAT"], enum["SDE_SAVE_EVERYSTEP"]]
    m = re.search(r"const SDE_LAYOUT_TRAJ_MAJOR, SDE_LAYOUT_SOA = Int32\((\d)\), Int32\((\d)\)", SHIM)
    assert [int(x) for x in m.groups()] == [enum["SDE_LAYOUT_TRAJ_MAJOR"], enum["SDE_LAYOUT_SOA"]]
    m = re.search(r"const SDE_COMPAT_FIX_VERN9_INTERP, SDE_COMPAT_STRICT_CONTROLLER, SDE_COMPAT_LOG2_CONTROLLER = Int32\((\d)\), Int32\((\d)\), Int32\((\d)\)", SHIM)
    assert [int(x) for x in m.groups()] == [enum["SDE_COMPAT_FIX_VERN9_INTERP"], enum["SDE_COMPAT_STRICT_CONTROLLER"], enum["SDE_COMPAT_LOG2_CONTROLLER"]]
    m = re.search(r"const SDE_COMPAT_FAST_RHS = Int32\((\d)\)", SHIM)
    assert int(m.group(1)) == enum["SDE_COMPAT_FAST_RHS"]
    m = re.search(r"const SDE_COMPAT_FAST_STAGES = Int32\((\d+)\)", SHIM)
    assert int(m.group(1)) == enum["SDE_COMPAT_FAST_STAGES"]
    assert enum["SDE_RET_DTMIN"] == 1 and "any(==(1), ret) && error(\"dt<dtmin\")" in SHIM
    # the python mirror and the shim construct the option struct with the same number of fields
    n_args = len(_split_top(re.search(r"SdeOptions\((alg_id\(alg\).*?capacity)\)\)", SHIM, flags=re.S).group(1)))
    assert n_args == len(julia_struct("SdeOptions"))
    n_args = len(_split_top(re.search(r"Ref\(SdeEmOptions\((.*?)\)\)\n", SHIM, flags=re.S).group(1)))
    assert n_args == len(julia_struct("SdeEmOptions"))


def test_python_ctypes_mirror_matches_the_header_too(sde):
    """The tested twin of the shim (simplediffeq.jl_b200/_lib.py): struct mirrors field by field and the argument
    count of every declared prototype, against the same parse of the header."""
    import ctypes
    from simplediffeq_b200 import _lib
    c2ct = {"int32_t": ctypes.c_int32, "int64_t": ctypes.c_int64, "uint64_t": ctypes.c_uint64, "double": ctypes.c_double,
            "const void*": ctypes.c_void_p}
    for tag, cls in (("sde_options", _lib.SdeOptions), ("sde_em_options", _lib.SdeEmOptions)):
        c_fields = header_struct(tag)
        assert [n for n, _ in c_fields] == [n for n, _ in cls._fields_], tag
        for (n, cty), (_, ct) in zip(c_fields, cls._fields_):
            assert ct is c2ct[cty], "%s.%s" % (tag, n)
    L = _lib.lib()
    protos = header_prototypes()
    assert set(protos) == set(_lib.EXPORTS)
    for name, (ret, params) in protos.items():
        fn = getattr(L, name)
        if params:
            assert fn.argtypes is not None and len(fn.argtypes) == len(params), name
        if ret == "int":
            assert fn.restype is ctypes.c_int, name


def test_shim_block_structure_is_balanced():
    """No Julia here, so nothing can `include` the shim; this at least keeps it from rotting syntactically: the file is
    tokenised with the Julia tokenizer of oracle/jlmini (strings, docstrings, comments, symbols) and every (, [, { must
    close in order, every block opener (module / function / struct / if / for / while / let / do / try / begin / macro /
    quote) outside brackets must meet its `end`, and the file must end at depth zero.  (`end` and `for` inside brackets
    or parentheses are an index / a generator, not blocks.)"""
    import sys
    sys.path.insert(0, os.path.join(ROOT, "oracle", "jlmini"))
    import jlmini
    toks = jlmini.tokenize(SHIM)
    openers = {"module", "function", "struct", "if", "for", "while", "let", "do", "try", "begin", "macro", "quote"}
    pairs = {")": "(", "]": "[", "}": "{"}
    stack, blocks = [], []
    prev = None
    for t in toks:
        if t.kind == "op" and t.val in "([{" and len(t.val) == 1:
            stack.append((t.val, t.line))
        elif t.kind == "op" and t.val in pairs:
            assert stack and stack[-1][0] == pairs[t.val], "unbalanced %r at line %d" % (t.val, t.line)
            stack.pop()
        elif t.kind == "kw" and not stack:
            if t.val in openers and not (t.val == "struct" and prev is not None and prev.kind == "kw" and prev.val == "mutable"
                                         and blocks and blocks[-1][0] == "mutable"):
                blocks.append((t.val, t.line))
            elif t.val == "end":
                assert blocks, "`end` without an open block at line %d" % t.line
                blocks.pop()
        if t.kind != "nl":
            prev = t
    assert not stack, "unclosed bracket opened at line %d" % stack[-1][1]
    assert not blocks, "unclosed `%s` opened at line %d" % blocks[-1]
    # every exported entry point the shim binds exists in the header (names only; signatures are checked above)
    for sym in set(re.findall(r"ccall\(\(:(\w+),", SHIM)):
        assert re.search(r"\b%s\s*\(" % sym, HEADER), sym
