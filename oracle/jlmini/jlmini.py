"""jlmini -- a small interpreter for the subset of Julia that SimpleDiffEq.jl's GPU-style `solve`
methods are written in.  TEST INFRASTRUCTURE ONLY (like everything under oracle/).

Why: Julia is not installed in the build container, so the reference cannot run here, and the
reference ships no golden vectors for this path.  Instead of trusting only a hand-written
restatement (oracle/oracle.cpp), this module EXECUTES THE REFERENCE'S OWN SOURCE TEXT: it parses
`src/tsit5/gpuatsit5.jl`, `src/rk4/gpurk4.jl`, `src/euler/gpueuler.jl`, `src/verner/gpuvern7.jl`,
`src/verner/gpuvern9.jl`, the tableau constructors (`src/tsit5/atsit5_cache.jl`,
`src/verner/verner_tableaus.jl`), `bθs` (`src/tsit5/tsit5.jl:385-399`),
`build_adaptive_controller_cache` (`src/SimpleDiffEq.jl:67-77`) and the right-hand sides of the
reference's tests (`test/gpusimpleatsit5_tests.jl:3-13`, `test/gpu_ode_regression.jl:2-4`) and
evaluates them with IEEE Float64 / Float32 scalars.  `oracle/jlmini/gen_golden.py` turns the
results into fixtures under tests/golden/ which pin the C++ oracle (tests/test_oracle_jlmini.py).
The out-of-place SimpleEM method of `src/euler_maruyama.jl:46-94` is executed the same way
(`gen_golden_em.py`, with `randn` supplied by the caller) and pins oracle/oracle_em.cpp.

What is NOT the reference's text and therefore restated here ([EXT], same assumptions A1-A9 as
SURVEY.md section 8c; each is one small function below so that it can be flipped):
  * `@muladd` (MuladdMacro.jl): `to_muladd` below applies the macro's rewriting to the parsed AST;
  * `muladd` on scalars = fused multiply-add; on SVectors element-wise (StaticArrays);
  * `@evalpoly` = Horner with muladd (Base.Math);
  * `@fastmath` `^` = libm pow / powf, `max`/`min` = ifelse(y > x, ...) forms, `/` = IEEE division;
  * `DiffEqBase.ODE_DEFAULT_NORM` = sqrt(sum(abs2, u) / length(u)) as a left fold, `abs` for scalars;
  * `a:s:b` float ranges = Base's TwicePrecision ranges (simplediffeq.jl_b200/jlrange.py);
  * `Base.min/max` propagate NaN; `build_solution` just carries (ts, us).

Not a general Julia implementation: it supports exactly the constructs those files use and raises
`JlSyntaxError` / `JlRuntimeError` on anything else.
"""
import ctypes
import ctypes.util
from fractions import Fraction
import math
import os
import re
import sys

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_ROOT = os.path.dirname(os.path.dirname(_HERE))


class JlSyntaxError(Exception):
    pass


class JlRuntimeError(Exception):
    pass


class JlError(Exception):
    """Julia `error("...")` raised by interpreted code."""


# =============================================================================================
# tokenizer
# =============================================================================================
_NUM_RE = re.compile(r"\d+\.\d+(?:[ef][+-]?\d+)?|\d+\.(?![\w.(])(?:[ef][+-]?\d+)?|\d+[ef][+-]?\d+|\d+")
_ID_RE = re.compile(r"[^\W\d]\w*", re.UNICODE)
_OPS = ["...", "===", "!==", ".+", ".-", ".*", "./", ".^", "==", "!=", "<=", ">=", "&&", "||", "+=", "-=",
        "*=", "/=", "//", "->", "::", "<:", "=", "<", ">", "+", "-", "*", "/", "^", "!", ":", ",", ";", "(", ")",
        "[", "]", "{", "}", ".", "?", "'"]
_KEYWORDS = {"function", "end", "if", "elseif", "else", "while", "for", "in", "return", "struct", "mutable",
             "begin", "where", "export", "using", "import", "const", "break", "continue", "module", "let",
             "do", "macro", "abstract", "primitive", "type"}


class Tok:
    __slots__ = ("kind", "val", "line", "sp")

    def __init__(self, kind, val, line, sp):
        self.kind, self.val, self.line, self.sp = kind, val, line, sp   # sp: whitespace before the token

    def __repr__(self):
        return "Tok(%s,%r,l%d)" % (self.kind, self.val, self.line)


def tokenize(src):
    toks = []
    i, n, line = 0, len(src), 1
    depth = 0
    sp = False
    while i < n:
        c = src[i]
        if c == "\n":
            if depth == 0 and toks and toks[-1].kind != "nl":
                toks.append(Tok("nl", "\n", line, sp))
            line += 1
            i += 1
            sp = True
            continue
        if c in " \t\r":
            i += 1
            sp = True
            continue
        if c == "#":
            if src.startswith("#=", i):
                lvl, j = 1, i + 2
                while j < n and lvl:
                    if src.startswith("#=", j):
                        lvl += 1
                        j += 2
                    elif src.startswith("=#", j):
                        lvl -= 1
                        j += 2
                    else:
                        if src[j] == "\n":
                            line += 1
                        j += 1
                i = j
            else:
                while i < n and src[i] != "\n":
                    i += 1
            sp = True
            continue
        if c == '"':
            if src.startswith('"""', i):
                j = src.find('"""', i + 3)
                if j < 0:
                    raise JlSyntaxError("unterminated triple-quoted string at line %d" % line)
                s = src[i + 3:j]
                line += s.count("\n")
                toks.append(Tok("str", s, line, sp))
                i = j + 3
            else:
                j = i + 1
                while j < n and src[j] != '"':
                    j += 2 if src[j] == "\\" else 1
                toks.append(Tok("str", src[i + 1:j], line, sp))
                i = j + 1
            sp = False
            continue
        if c.isdigit():
            m = _NUM_RE.match(src, i)
            toks.append(Tok("num", m.group(0), line, sp))
            i = m.end()
            sp = False
            continue
        if c == "@":
            m = _ID_RE.match(src, i + 1)
            if not m:       # @. / @.. : only ever skipped, never evaluated
                j = i + 1
                while j < n and not src[j].isspace():
                    j += 1
                toks.append(Tok("macro", src[i + 1:j], line, sp))
                i = j
                sp = False
                continue
            toks.append(Tok("macro", m.group(0), line, sp))
            i = m.end()
            sp = False
            continue
        m = _ID_RE.match(src, i)
        if m:
            name = m.group(0)
            j = m.end()
            if j < n and src[j] == "!" and not src.startswith("!=", j):
                name += "!"
                j += 1
            toks.append(Tok("kw" if name in _KEYWORDS else "id", name, line, sp))
            i = j
            sp = False
            continue
        for op in _OPS:
            if src.startswith(op, i):
                # `1 .+ x` style dotted operators only; a '.' directly followed by '(' or an identifier is
                # field access / broadcast call
                toks.append(Tok("op", op, line, sp))
                if op in "([{":
                    depth += 1
                elif op in ")]}":
                    depth -= 1
                i += len(op)
                break
        else:
            toks.append(Tok("op", c, line, sp))     # characters only skipped code uses ($, %, &, ...)
            i += 1
        sp = False
    toks.append(Tok("nl", "\n", line, True))
    toks.append(Tok("eof", None, line, True))
    return toks


# =============================================================================================
# parser  (AST = nested tuples, first element = node kind)
# =============================================================================================
_CMP_OPS = {"==", "!=", "===", "!==", "<", "<=", ">", ">=", "<:"}
_ASSIGN_OPS = {"=", "+=", "-=", "*=", "/="}


class Parser:
    def __init__(self, toks, fname="<src>"):
        self.t = toks
        self.i = 0
        self.fname = fname

    # ---- token helpers
    def peek(self, k=0):
        return self.t[self.i + k]

    def next(self):
        tok = self.t[self.i]
        self.i += 1
        return tok

    def err(self, msg):
        tok = self.peek()
        raise JlSyntaxError("%s:%d: %s (at %r)" % (self.fname, tok.line, msg, tok.val))

    def is_op(self, *ops):
        tok = self.peek()
        return tok.kind == "op" and tok.val in ops

    def is_kw(self, *kws):
        tok = self.peek()
        return tok.kind == "kw" and tok.val in kws

    def expect_op(self, op):
        if not self.is_op(op):
            self.err("expected %r" % op)
        return self.next()

    def expect_kw(self, kw):
        if not self.is_kw(kw):
            self.err("expected %r" % kw)
        return self.next()

    def skip_nl(self):
        while self.peek().kind == "nl" or self.is_op(";"):
            self.next()

    def skip_only_nl(self):
        while self.peek().kind == "nl":
            self.next()

    # ---- blocks and statements
    def parse_block(self, terminators=("end",)):
        stmts = []
        while True:
            self.skip_nl()
            if self.peek().kind == "eof":
                self.err("unexpected end of file in block")
            if self.peek().kind == "kw" and self.peek().val in terminators:
                return ("block", stmts)
            stmts.append(self.parse_statement())

    def parse_statement(self):
        tok = self.peek()
        if tok.kind == "kw":
            kw = tok.val
            if kw == "function":
                return self.parse_function()
            if kw in ("struct", "mutable"):
                return self.parse_struct()
            if kw == "if":
                return self.parse_if()
            if kw == "while":
                self.next()
                cond = self.parse_expr()
                body = self.parse_block()
                self.expect_kw("end")
                return ("while", cond, body)
            if kw == "for":
                self.next()
                var = self.next()
                if var.kind != "id":
                    self.err("for: loop variable expected")
                if self.is_kw("in") or self.is_op("="):
                    self.next()
                else:
                    self.err("for: `in` expected")
                it = self.parse_expr()
                body = self.parse_block()
                self.expect_kw("end")
                return ("for", var.val, it, body)
            if kw == "return":
                self.next()
                if self.peek().kind == "nl" or self.is_kw("end"):
                    return ("return", ("id", "nothing"))
                return ("return", self.parse_comma_expr())
            if kw == "begin":
                self.next()
                body = self.parse_block()
                self.expect_kw("end")
                return body
            if kw in ("break", "continue"):
                self.next()
                return (kw,)
            self.err("unsupported keyword %r" % kw)
        if tok.kind == "macro":
            return self.parse_macro(statement=True)
        lhs = self.parse_comma_expr()
        if self.peek().kind == "op" and self.peek().val in _ASSIGN_OPS:
            op = self.next().val
            self.skip_only_nl()
            rhs = self.parse_assignment_rhs()
            if op == "=":
                return ("assign", lhs, rhs)
            return ("opassign", op[0], lhs, rhs)
        return lhs

    def parse_assignment_rhs(self):
        rhs = self.parse_comma_expr()
        if self.is_op("="):     # a = b = c
            self.next()
            self.skip_only_nl()
            return ("assign", rhs, self.parse_assignment_rhs())
        return rhs

    def parse_comma_expr(self):
        """expr [, expr ...] at statement level (tuple without parentheses)."""
        first = self.parse_expr()
        if not self.is_op(","):
            return first
        elems = [first]
        while self.is_op(","):
            self.next()
            self.skip_only_nl()
            elems.append(self.parse_expr())
        return ("tuple", elems)

    def parse_macro(self, statement):
        tok = self.next()
        name = tok.val
        if self.is_op("(") and not self.peek().sp:
            self.next()
            args, kwargs = self.parse_call_args(")")
            if kwargs:
                self.err("keyword arguments in a macro call")
            return ("macro", name, args)
        if self.peek().kind == "nl":
            return ("macro", name, [])
        if statement:
            return ("macro", name, [self.parse_statement()])
        return ("macro", name, [self.parse_expr()])

    def parse_if(self):
        self.expect_kw("if")
        branches = []
        cond = self.parse_expr()
        body = self.parse_block(("elseif", "else", "end"))
        branches.append((cond, body))
        other = None
        while True:
            if self.is_kw("elseif"):
                self.next()
                cond = self.parse_expr()
                body = self.parse_block(("elseif", "else", "end"))
                branches.append((cond, body))
            elif self.is_kw("else"):
                self.next()
                other = self.parse_block(("end",))
            else:
                self.expect_kw("end")
                return ("if", branches, other)

    def parse_struct(self):
        if self.is_kw("mutable"):
            self.next()
        self.expect_kw("struct")
        name = self.next().val
        tparams = []
        if self.is_op("{"):
            self.next()
            while not self.is_op("}"):
                tparams.append(self.next().val)
                if self.is_op(","):
                    self.next()
            self.next()
        if self.is_op("<:"):
            self.next()
            self.parse_postfix()
        fields = []
        while True:
            self.skip_nl()
            if self.is_kw("end"):
                self.next()
                break
            f = self.next()
            if f.kind != "id":
                self.err("struct field expected")
            if self.is_op("::"):
                self.next()
                self.parse_postfix()
            fields.append(f.val)
        return ("struct", name, tparams, fields)

    def parse_function(self):
        self.expect_kw("function")
        name = self.next().val
        while self.is_op("."):
            self.next()
            name = self.next().val          # DiffEqBase.solve -> solve
        self.expect_op("(")
        params, kwparams = [], []
        in_kw = False
        while True:
            self.skip_only_nl()
            if self.is_op(")"):
                self.next()
                break
            if self.is_op(";"):
                self.next()
                in_kw = True
                continue
            pname, ptype, default, splat = None, None, None, False
            if self.is_op("::"):
                self.next()
                ptype = self.parse_postfix()
            else:
                pname = self.next().val
                if self.is_op("::"):
                    self.next()
                    ptype = self.parse_postfix()
                if self.is_op("..."):
                    self.next()
                    splat = True
                if self.is_op("="):
                    self.next()
                    default = self.parse_expr()
            (kwparams if in_kw else params).append((pname, ptype, default, splat))
            if self.is_op(","):
                self.next()
        where = {}
        if self.is_kw("where"):
            self.next()
            braces = self.is_op("{")
            if braces:
                self.next()
            while True:
                if braces:
                    self.skip_only_nl()
                    if self.is_op("}"):      # trailing comma: `where {\n uType,\n tType,\n }`
                        break
                tv = self.next().val
                bound = None
                if self.is_op("<:"):
                    self.next()
                    bound = self.parse_postfix()
                where[tv] = bound
                if braces:
                    self.skip_only_nl()
                if self.is_op(","):
                    self.next()
                    continue
                break
            if braces:
                self.expect_op("}")
        body = self.parse_block()
        self.expect_kw("end")
        return ("function", name, params, kwparams, where, body)

    # ---- expressions
    def parse_expr(self):
        return self.parse_ternary()

    def parse_ternary(self):
        c = self.parse_or()
        if self.is_op("?"):
            self.next()
            a = self.parse_ternary()
            self.expect_op(":")
            b = self.parse_ternary()
            return ("ternary", c, a, b)
        return c

    def parse_or(self):
        a = self.parse_and()
        while self.is_op("||"):
            self.next()
            self.skip_only_nl()
            a = ("or", a, self.parse_and())
        return a

    def parse_and(self):
        a = self.parse_cmp()
        while self.is_op("&&"):
            self.next()
            self.skip_only_nl()
            a = ("and", a, self.parse_cmp())
        return a

    def parse_cmp(self):
        a = self.parse_range()
        if self.peek().kind == "id" and self.peek().val == "isa":     # `x isa T`
            self.next()
            return ("isa", a, self.parse_range())
        if self.peek().kind == "op" and self.peek().val in _CMP_OPS:
            operands, ops = [a], []
            while self.peek().kind == "op" and self.peek().val in _CMP_OPS:
                ops.append(self.next().val)
                self.skip_only_nl()
                operands.append(self.parse_range())
            return ("cmp", operands, ops)
        return a

    def parse_range(self):
        a = self.parse_add()
        if self.is_op(":") and not self._in_ternary_else():
            self.next()
            b = self.parse_add()
            if self.is_op(":"):
                self.next()
                c = self.parse_add()
                return ("range", a, b, c)
            return ("range", a, None, b)
        return a

    def _in_ternary_else(self):
        return False    # the supported sources have no `?:`; kept for clarity

    def parse_add(self):
        """Julia flattens chains of `+` into one n-ary call (`a + b + c` is +(a, b, c)); `-` and the
        dotted operators are binary and left-associative.  This matters for @muladd."""
        a = self.parse_mul()
        chain = False    # True while `a` is a flattenable +-call produced by this loop
        while self.peek().kind == "op" and self.peek().val in ("+", "-", ".+", ".-"):
            op = self.next().val
            self.skip_only_nl()
            b = self.parse_mul()
            if op == "+" and chain:
                a[2].append(b)
            else:
                a = ("op", op, [a, b])
                chain = (op == "+")
        return a

    def parse_mul(self):
        a = self.parse_unary()
        chain = False
        while self.peek().kind == "op" and self.peek().val in ("*", "/", ".*", "./", "//"):
            op = self.next().val
            self.skip_only_nl()
            b = self.parse_unary()
            if op == "*" and chain:
                a[2].append(b)
            else:
                a = ("op", op, [a, b])
                chain = (op == "*")
        return a

    def parse_unary(self):
        self.skip_only_nl()
        if self.is_op("-"):
            self.next()
            if self.peek().kind == "num" and not self.peek().sp:
                # `-4` is a literal in Julia's parser (unless followed by ^)
                operand = self.parse_power()
                if operand[0] == "num":
                    return ("num", "-" + operand[1])
                return ("neg", operand)
            return ("neg", self.parse_unary())
        if self.is_op("+"):
            self.next()
            return self.parse_unary()
        if self.is_op("!"):
            self.next()
            return ("not", self.parse_unary())
        return self.parse_power()

    def parse_power(self):
        base = self.parse_juxt()
        if self.is_op("^", ".^"):
            op = self.next().val
            expo = self.parse_unary()
            return ("op", op, [base, expo])
        return base

    def parse_juxt(self):
        tok = self.peek()
        a = self.parse_postfix()
        if tok.kind == "num" and a[0] == "num":
            nxt = self.peek()
            if nxt.kind == "id" and not nxt.sp:      # 2k2  ->  2 * k2
                b = self.parse_postfix()
                return ("op", "*", [a, b])
        return a

    def parse_postfix(self):
        a = self.parse_primary()
        while True:
            tok = self.peek()
            if tok.kind != "op" or tok.sp and tok.val in ("(", "[", "{"):
                return a
            if tok.val == "(":
                self.next()
                args, kwargs = self.parse_call_args(")")
                a = ("call", a, args, kwargs, False)
            elif tok.val == "[":
                self.next()
                args, kwargs = self.parse_call_args("]")
                a = ("index", a, args)
            elif tok.val == "{":
                self.next()
                args, kwargs = self.parse_call_args("}")
                a = ("curly", a, args)
            elif tok.val == "." and not tok.sp:
                nxt = self.peek(1)
                if nxt.kind == "op" and nxt.val == "(" and not nxt.sp:
                    self.next()
                    self.next()
                    args, kwargs = self.parse_call_args(")")
                    a = ("call", a, args, kwargs, True)
                elif nxt.kind in ("id", "kw") and not nxt.sp:
                    self.next()
                    a = ("field", a, self.next().val)
                else:
                    return a
            elif tok.val == "::":
                self.next()
                ty = self.parse_primary_with_curly()
                a = ("typed", a, ty)
            else:
                return a

    def parse_primary_with_curly(self):
        a = self.parse_primary()
        while self.is_op("{") and not self.peek().sp:
            self.next()
            args, _ = self.parse_call_args("}")
            a = ("curly", a, args)
        return a

    def parse_call_args(self, closer):
        args, kwargs = [], []
        in_kw = False
        while True:
            self.skip_only_nl()
            if self.is_op(closer):
                self.next()
                return args, kwargs
            if self.is_op(";"):
                self.next()
                in_kw = closer == ")"
                continue
            if self.is_op(","):
                self.next()
                continue
            e = self.parse_expr()
            if self.is_op("=") and e[0] == "id" and closer == ")":
                self.next()
                self.skip_only_nl()
                kwargs.append((e[1], self.parse_expr()))
            elif self.is_op("..."):
                self.next()
                args.append(("splat", e))
            elif in_kw and e[0] == "id":
                kwargs.append((e[1], e))
            else:
                args.append(e)

    def parse_primary(self):
        self.skip_only_nl()
        tok = self.next()
        if tok.kind == "num":
            return ("num", tok.val)
        if tok.kind == "id":
            return ("id", tok.val)
        if tok.kind == "str":
            return ("str", tok.val)
        if tok.kind == "macro":
            self.i -= 1
            return self.parse_macro(statement=False)
        if tok.kind == "op":
            if tok.val == "(":
                self.skip_only_nl()
                if self.is_op(")"):
                    self.next()
                    return ("tuple", [])
                e = self.parse_expr()
                self.skip_only_nl()
                if self.is_op(","):
                    elems = [e]
                    while self.is_op(","):
                        self.next()
                        self.skip_only_nl()
                        if self.is_op(")"):
                            break
                        elems.append(self.parse_expr())
                        self.skip_only_nl()
                    self.expect_op(")")
                    return ("tuple", elems)
                self.expect_op(")")
                return ("paren", e)
            if tok.val == "[":
                save = self.i
                self.skip_only_nl()
                if not self.is_op("]"):
                    first = self.parse_expr()
                    if self.is_kw("for"):     # comprehension [expr for v in iter]
                        self.next()
                        var = self.next()
                        if not (self.is_kw("in") or self.is_op("=")):
                            self.err("comprehension: `in` expected")
                        self.next()
                        it = self.parse_expr()
                        self.skip_only_nl()
                        self.expect_op("]")
                        return ("comprehension", first, var.val, it)
                self.i = save
                args, _ = self.parse_call_args("]")
                return ("vect", args)
            if tok.val == ":":   # symbol literal
                return ("str", self.next().val)
        self.i -= 1
        self.err("expression expected")


def parse_definitions(src, fname, wanted=None, first_only=False):
    """Parse the top-level `function` / `struct` definitions of a file (optionally preceded by macros
    such as @muladd / @inline); everything else at top level is skipped line by line."""
    toks = tokenize(src)
    p = Parser(toks, fname)
    defs = []
    while p.peek().kind != "eof":
        tok = p.peek()
        start = p.i
        macros = []
        while p.peek().kind == "macro" and p.peek().val in ("muladd", "inline", "inbounds", "noinline"):
            macros.append(p.next().val)
        if p.is_kw("function") or p.is_kw("struct") or (p.is_kw("mutable") and p.peek(1).val == "struct"):
            name_tok = p.peek(1)
            if wanted is not None and p.is_kw("function"):
                # resolve dotted name
                j = p.i + 1
                nm = p.t[j].val
                while p.t[j + 1].kind == "op" and p.t[j + 1].val == ".":
                    j += 2
                    nm = p.t[j].val
                if nm not in wanted or (first_only and any(d[0][0] == "function" and d[0][1] == nm for d in defs)):
                    p.i = start
                    _skip_definition(p)
                    continue
            node = p.parse_statement()
            defs.append((node, macros, tok.line))
            continue
        p.i = start
        _skip_line(p)
    return defs


_BLOCK_OPENERS = ("function", "struct", "if", "while", "for", "begin", "module", "let", "do", "macro", "quote",
                  "try")


def _skip_line(p):
    """Skip to the end of the current top-level line; block openers inside it are balanced (`end` inside
    brackets, as in `us[end]`, is an index, not a block closer)."""
    depth = brackets = 0
    while p.peek().kind != "eof":
        tok = p.next()
        if tok.kind == "op" and tok.val == "[":
            brackets += 1
        elif tok.kind == "op" and tok.val == "]":
            brackets -= 1
        elif tok.kind == "kw" and tok.val in _BLOCK_OPENERS and brackets == 0:
            depth += 1
        elif tok.kind == "kw" and tok.val == "end" and brackets == 0:
            depth -= 1
        elif tok.kind == "nl" and depth <= 0:
            return


def _skip_definition(p):
    depth = brackets = 0
    while p.peek().kind != "eof":
        tok = p.next()
        if tok.kind == "op" and tok.val == "[":
            brackets += 1
        elif tok.kind == "op" and tok.val == "]":
            brackets -= 1
        elif tok.kind == "kw" and tok.val in _BLOCK_OPENERS and brackets == 0:
            depth += 1
        elif tok.kind == "kw" and tok.val == "end" and brackets == 0:
            depth -= 1
            if depth == 0:
                return


# =============================================================================================
# @muladd  ([EXT] MuladdMacro.jl >= 0.2.4, assumptions A1-A3 of SURVEY.md section 8c)
# =============================================================================================
MULADD_NOTES = []   # (description, line) of constructs the macro may treat differently between versions


def _is_mul(node, dotted):
    return node[0] == "op" and node[1] == (".*" if dotted else "*")


def to_muladd(node):
    """x + a*b + c*d  -> muladd(c, d, muladd(a, b, x));  a*b + c*d + e*f -> muladd(e, f, muladd(c, d, a*b));
    a product of more than two factors splits as (all-but-last) * last; a dotted `.+` fuses only
    dotted `.*` operands.  Everything else is rebuilt with its children transformed."""
    if not isinstance(node, tuple):
        if isinstance(node, list):
            return [to_muladd(x) for x in node]
        return node
    kind = node[0]
    if kind == "op" and node[1] in ("+", ".+"):
        dotted = node[1] == ".+"
        operands = [to_muladd(x) for x in node[2]]
        muls = [x for x in operands if _is_mul(x, dotted)]
        if not muls:
            return ("op", node[1], operands)
        odd = [x for x in operands if not _is_mul(x, dotted)]
        if not odd:
            acc = muls[0]
            rest = muls[1:]
        else:
            acc = odd[0] if len(odd) == 1 else ("op", node[1], odd)
            rest = muls
        for m in rest:
            factors = m[2]
            head = factors[0] if len(factors) == 2 else ("op", m[1], factors[:-1])
            acc = ("muladd", head, factors[-1], acc, dotted)
        return acc
    if kind == "op" and node[1] in ("-", ".-"):
        operands = [to_muladd(x) for x in node[2]]
        if any(_is_mul(x, node[1] == ".-") for x in operands):
            MULADD_NOTES.append("subtraction with a product operand (newer MuladdMacro versions fuse it)")
        return ("op", node[1], operands)
    if kind == "opassign":
        rhs = to_muladd(node[3])
        if node[1] == "+" and _is_mul(rhs, False):
            MULADD_NOTES.append("`+=` with a product right-hand side")
        return ("opassign", node[1], node[2], rhs)
    if kind == "if":
        return ("if", [(to_muladd(c), to_muladd(b)) for c, b in node[1]], to_muladd(node[2]) if node[2] else None)
    if kind == "call":
        return ("call", to_muladd(node[1]), [to_muladd(a) for a in node[2]],
                [(k, to_muladd(v)) for k, v in node[3]], node[4])
    if kind == "function":
        return node[:5] + (to_muladd(node[5]),)
    if kind in ("num", "id", "str", "struct", "break", "continue"):
        return node
    return tuple(to_muladd(x) if isinstance(x, (tuple, list)) else x for x in node)


# =============================================================================================
# values
# =============================================================================================
_libm = ctypes.CDLL(ctypes.util.find_library("m") or "libm.so.6")
_libm.fma.restype = ctypes.c_double
_libm.fma.argtypes = [ctypes.c_double] * 3
_libm.fmaf.restype = ctypes.c_float
_libm.fmaf.argtypes = [ctypes.c_float] * 3
_libm.pow.restype = ctypes.c_double
_libm.pow.argtypes = [ctypes.c_double] * 2
_libm.powf.restype = ctypes.c_float
_libm.powf.argtypes = [ctypes.c_float] * 2

F64, F32 = np.float64, np.float32


class JlType:
    def __init__(self, name, np_type=None):
        self.name, self.np = name, np_type

    def __call__(self, x):
        return convert(self, x)

    def __repr__(self):
        return self.name


T_F64 = JlType("Float64", F64)
T_F32 = JlType("Float32", F32)
T_INT = JlType("Int64", int)
T_BOOL = JlType("Bool", bool)


def typeof_scalar(x):
    if isinstance(x, (bool, np.bool_)):
        return T_BOOL
    if isinstance(x, F32):
        return T_F32
    if isinstance(x, F64):
        return T_F64
    if isinstance(x, (int, np.integer)):
        return T_INT
    raise JlRuntimeError("typeof: unsupported value %r" % (x,))


def convert(T, x):
    if isinstance(x, SVec):
        return SVec([convert(T, v) for v in x.v])
    if isinstance(x, Fraction):     # convert(T, n//d) = T(n) / T(d)
        return T.np(x.numerator) / T.np(x.denominator)
    if T is T_F64:
        return F64(x)
    if T is T_F32:
        return F32(x)      # one rounding from the source value
    if T is T_INT:
        if isinstance(x, (float, np.floating)) and float(x) != int(x):
            raise JlError("InexactError: Int64(%r)" % (x,))
        return int(x)
    raise JlRuntimeError("convert: unsupported target %r" % (T,))


def is_float(x):
    return isinstance(x, (F64, F32))


def promote2(a, b):
    """Julia promotion of two real scalars to a common floating type (ints stay ints together)."""
    fa, fb = is_float(a), is_float(b)
    if fa and fb:
        if type(a) is type(b):
            return a, b
        return F64(a), F64(b)
    if fa:
        return a, type(a)(b)
    if fb:
        return type(b)(a), b
    return a, b


class SVec:
    """StaticArrays.SVector restricted to what the solvers use."""
    __slots__ = ("v",)

    def __init__(self, v):
        self.v = tuple(v)

    def __len__(self):
        return len(self.v)

    def __repr__(self):
        return "SVec(%s)" % (", ".join(repr(float(x)) for x in self.v),)


class JlVector:
    """Vector / MVector (1-based, growable); entries start out as None (`undef`)."""

    def __init__(self, items, eltype=None, fixed=False):
        self.items = list(items)
        self.eltype = eltype
        self.fixed = fixed

    def __len__(self):
        return len(self.items)


class UnitRange:
    def __init__(self, a, b):
        self.a, self.b = int(a), int(b)

    def __len__(self):
        return max(0, self.b - self.a + 1)


class FloatRange:
    """`a:s:b` on floats -- Base's TwicePrecision range via the host layer's restatement."""

    def __init__(self, a, s, b):
        sys.path.insert(0, os.path.join(_ROOT, "simplediffeq.jl_b200"))
        import jlrange
        a, s = promote2(a, s)
        a, b = promote2(a, b)
        a, s = promote2(a, s)
        if not is_float(a):
            a, s, b = F64(a), F64(s), F64(b)
        self.T = type(a)
        self.r = jlrange.JuliaRange(a, s, b, dtype=self.T)
        self.vals = self.r.collect()

    def __len__(self):
        return len(self.r)


class Struct:
    def __init__(self, tname, fields, values):
        self.tname = tname
        self.fields = dict(zip(fields, values))

    def __repr__(self):
        return "%s(...)" % self.tname


class StructType:
    def __init__(self, name, tparams, fields):
        self.name, self.tparams, self.fields = name, tparams, fields


class TypeApp:
    def __init__(self, base, params):
        self.base, self.params = base, params


class Problem:
    """SciMLBase.ODEProblem{false}: f, u0, tspan, p."""

    def __init__(self, f, u0, tspan, p):
        self.f, self.u0, self.tspan, self.p = f, u0, tuple(tspan), p


class SDEProblem:
    """SciMLBase.SDEProblem{uType, tType, false}: f, g, u0, tspan, p (diagonal noise: no noise_rate_prototype)."""

    def __init__(self, f, g, u0, tspan, p=None, noise_rate_prototype=None):
        self.f, self.g, self.u0, self.tspan, self.p = f, g, u0, tuple(tspan), p
        self.noise_rate_prototype = noise_rate_prototype
        self.iip = False


class Solution:
    def __init__(self, t, u):
        self.t, self.u = t, u
        self.retcode = "Default"


class Function:
    def __init__(self, name):
        self.name = name
        self.methods = []   # (params, kwparams, where, body)


class _Return(Exception):
    def __init__(self, v):
        self.v = v


class _Break(Exception):
    pass


class _Continue(Exception):
    pass


# ---- arithmetic -----------------------------------------------------------------------------
def fma_scalar(a, b, c):
    """muladd(a, b, c) on real scalars: fused for floats ([EXT]: LLVM fmuladd -> hardware FMA)."""
    a, b = promote2(a, b)
    a, c = promote2(a, c)
    a, b = promote2(a, b)
    b, c = promote2(b, c)
    if isinstance(a, F64):
        return F64(_libm.fma(float(a), float(b), float(c)))
    if isinstance(a, F32):
        return F32(_libm.fmaf(float(a), float(b), float(c)))
    return a * b + c


def jl_muladd(a, b, c):
    sa, sb, sc = isinstance(a, SVec), isinstance(b, SVec), isinstance(c, SVec)
    if not (sa or sb or sc):
        return fma_scalar(a, b, c)
    # StaticArrays: muladd(scalar, SA, SA) and muladd(SA, scalar, SA) map the scalar muladd (assumption A4)
    if sc and sb and not sa:
        return SVec(fma_scalar(a, x, y) for x, y in zip(b.v, c.v))
    if sc and sa and not sb:
        return SVec(fma_scalar(x, b, y) for x, y in zip(a.v, c.v))
    raise JlRuntimeError("muladd: unsupported operand shapes")


def _arith(op, a, b):
    if op == "//":
        return Fraction(int(a), int(b))
    a, b = promote2(a, b)
    with np.errstate(all="ignore"):
        if op == "+":
            return a + b
        if op == "-":
            return a - b
        if op == "*":
            return a * b
        if op == "/":
            if not is_float(a):
                a, b = F64(a), F64(b)
            return a / b
    raise JlRuntimeError("bad operator " + op)


def jl_binop(op, a, b, broadcast=False):
    sa, sb = isinstance(a, SVec), isinstance(b, SVec)
    if not sa and not sb:
        return _arith(op, a, b)
    if sa and sb:
        if len(a) != len(b):
            raise JlRuntimeError("DimensionMismatch")
        if op in ("+", "-") or broadcast:
            return SVec(_arith(op, x, y) for x, y in zip(a.v, b.v))
        raise JlRuntimeError("SVector %s SVector is not defined without a dot" % op)
    if sa:
        if op in ("*", "/") or broadcast:
            return SVec(_arith(op, x, b) for x in a.v)
        raise JlRuntimeError("SVector %s scalar needs a dot" % op)
    if op == "*" or broadcast:
        return SVec(_arith(op, a, y) for y in b.v)
    raise JlRuntimeError("scalar %s SVector needs a dot" % op)


def jl_pow(x, y, fast):
    x, y = promote2(x, y)
    if isinstance(x, F64):
        return F64(_libm.pow(float(x), float(y)))
    if isinstance(x, F32):
        return F32(_libm.powf(float(x), float(y)))
    return x ** y


def jl_min(a, b):
    a, b = promote2(a, b)
    if is_float(a) and (a != a or b != b):
        return a + b
    return b if b < a else a


def jl_max(a, b):
    a, b = promote2(a, b)
    if is_float(a) and (a != a or b != b):
        return a + b
    return b if b > a else a


def fast_min(x, y):      # Base.FastMath.min_fast(x, y) = ifelse(y > x, x, y)
    x, y = promote2(x, y)
    return x if y > x else y


def fast_max(x, y):      # Base.FastMath.max_fast(x, y) = ifelse(y > x, y, x)
    x, y = promote2(x, y)
    return y if y > x else x


def jl_abs(x):
    return type(x)(abs(x)) if is_float(x) else abs(x)


def jl_sqrt(x):
    with np.errstate(all="ignore"):
        return np.sqrt(x) if is_float(x) else np.sqrt(F64(x))


def ode_default_norm(u, t):
    """[EXT] DiffEqBase.ODE_DEFAULT_NORM: sqrt(sum(abs2, u) / length(u)), scalars: abs (A5, A6)."""
    if isinstance(u, SVec):
        acc = None
        for x in u.v:
            sq = _arith("*", x, x)
            acc = sq if acc is None else _arith("+", acc, sq)
        return jl_sqrt(_arith("/", acc, len(u)))
    return jl_abs(u)


def evalpoly(x, coefs):
    """[EXT] Base.Math.@evalpoly: Horner with muladd."""
    acc = coefs[-1]
    for c in reversed(coefs[:-1]):
        acc = fma_scalar(x, acc, c)
    return acc


# =============================================================================================
# interpreter
# =============================================================================================
class Env:
    def __init__(self, parent=None):
        self.vars = {}
        self.parent = parent

    def lookup(self, name):
        e = self
        while e is not None:
            if name in e.vars:
                return e.vars[name]
            e = e.parent
        raise JlRuntimeError("UndefVarError: %s" % name)

    def set(self, name, val):
        # Julia: assignment inside a function updates an existing local of an enclosing scope of the same
        # function, else creates a local
        e = self
        while e is not None and e.parent is not None:
            if name in e.vars:
                e.vars[name] = val
                return
            e = e.parent
        self.vars[name] = val


class Module:
    def __init__(self, interp):
        self.interp = interp


class Interp:
    def __init__(self):
        self.globals = Env()
        self.fastmath = 0
        self.stats = {"f_calls": 0}
        g = self.globals.vars
        g.update({
            "nothing": None, "true": True, "false": False, "Inf": F64(np.inf), "NaN": F64(np.nan),
            "undef": "undef",
            "Float64": T_F64, "Float32": T_F32, "Int": T_INT, "Int64": T_INT, "Bool": T_BOOL,
            "DiffEqBase": Module(self), "SciMLBase": Module(self), "Base": Module(self),
            "isinplace": lambda prob: False,
            "recursivecopy": lambda x: x,
            "copy": lambda x: x,
            "length": self._length,
            "eltype": self._eltype,
            "typeof": self._typeof,
            "convert": convert,
            "zero": lambda T: convert(T, 0) if isinstance(T, JlType) else convert(self._typeof(T), 0),
            "one": lambda T: convert(T, 1),
            "inv": lambda x: _arith("/", type(x)(1) if is_float(x) else 1, x),
            "iszero": lambda x: bool(x == 0),
            "isnan": lambda x: bool(x != x),
            "abs": jl_abs,
            "abs2": lambda x: _arith("*", x, x),
            "sqrt": jl_sqrt,
            "min": jl_min, "max": jl_max,
            "muladd": jl_muladd,
            "push!": self._push,
            "error": self._error,
            "ODE_DEFAULT_NORM": ode_default_norm,
            "build_solution": lambda prob, alg, ts, us, **kw: Solution(ts, us),
            "has_analytic": lambda f: False,
            "calculate_solution_errors!": lambda *a, **k: None,
            "SVector": "SVector", "MVector": "MVector", "Vector": "Vector", "SArray": self._sarray,
            "ODEProblem": "ODEProblem", "Type": "Type", "Number": "Number", "Real": "Real",
            # SimpleEM (src/euler_maruyama.jl): Julia's task-local `randn` is not reproducible outside
            # its process, so the normals are handed in by the caller (self.randn_hook)
            "is_diagonal_noise": lambda prob: prob.noise_rate_prototype is None,
            "randn": lambda *a: self.randn_hook(*a),
            "size": lambda x, d: x.shape[d - 1],
        })
        self.randn_hook = self._no_randn

    @staticmethod
    def _no_randn(*a):
        raise JlRuntimeError("randn called without a noise source (set Interp.randn_hook)")

    # ---- builtins
    @staticmethod
    def _length(x):
        return len(x)

    def _eltype(self, x):
        if isinstance(x, SVec):
            return typeof_scalar(x.v[0])
        if isinstance(x, JlVector):
            return x.eltype
        if isinstance(x, FloatRange):
            return T_F64 if x.T is F64 else T_F32
        return typeof_scalar(x)

    def _typeof(self, x):
        if isinstance(x, SVec):
            return TypeApp("SVector", [len(x), typeof_scalar(x.v[0])])
        if isinstance(x, Struct):
            return x.tname
        return typeof_scalar(x)

    @staticmethod
    def _push(vec, x):
        if vec.fixed:
            raise JlRuntimeError("push! on a fixed-size vector")
        if isinstance(vec.eltype, JlType) and not isinstance(x, SVec):
            x = convert(vec.eltype, x)       # push!(::Vector{T}, x) converts
        vec.items.append(x)
        return vec

    @staticmethod
    def _error(msg):
        raise JlError(msg)

    @staticmethod
    def _sarray(x):
        return x

    # ---- loading definitions
    def load(self, path, wanted=None, force_muladd=(), first_only=False):
        """first_only: take only the first method of each wanted name (euler_maruyama.jl: the out-of-place
        method; the in-place one that follows uses broadcast-assignment macros outside the subset)."""
        src = open(path, encoding="utf-8").read()
        for node, macros, line in parse_definitions(src, os.path.basename(path), wanted, first_only):
            if node[0] == "struct":
                _, name, tparams, fields = node
                self.globals.vars[name] = StructType(name, tparams, fields)
                continue
            _, name, params, kwparams, where, body = node
            if "muladd" in macros:
                body = to_muladd(body)
            fn = self.globals.vars.get(name)
            if isinstance(fn, StructType):
                key = "#ctor#" + name
                fn = self.globals.vars.get(key)
                if fn is None:
                    fn = self.globals.vars[key] = Function(name)
            elif not isinstance(fn, Function):
                fn = self.globals.vars[name] = Function(name)
            fn.methods.append((params, kwparams, where, body, line, os.path.basename(path)))

    # ---- dispatch
    def _matches(self, ann, val, where):
        if ann is None:
            return True
        kind = ann[0]
        if kind == "id":
            name = ann[1]
            if name in where:
                bound = where[name]
                return True if bound is None else self._matches(bound, val, {})
            if name == "ODEProblem":
                return isinstance(val, Problem)
            if name == "Any":
                return True
            if isinstance(val, Struct):
                return val.tname == name
            if name in ("Float64", "Float32"):
                return isinstance(val, F64 if name == "Float64" else F32)
            if name == "Number" or name == "Real":
                return is_float(val) or isinstance(val, int)
            return False
        if kind == "curly":
            base = ann[1][1]
            if base == "Type":
                return isinstance(val, JlType)
            if base == "SDEProblem":            # SDEProblem{uType, tType, isinplace}
                if not isinstance(val, SDEProblem):
                    return False
                iip = ann[2][2] if len(ann[2]) > 2 else None
                return iip is None or iip[0] != "id" or iip[1] not in ("true", "false") or (iip[1] == "true") == val.iip
            if base == "SVector":
                n = ann[2][0]
                return isinstance(val, SVec) and (n[0] != "num" or len(val) == int(n[1]))
            if isinstance(val, Struct):
                return val.tname == base
            return False
        return False

    def call_function(self, fn, args, kwargs):
        cands = []
        for m in fn.methods:
            params = m[0]
            if params and params[-1][3]:          # trailing `args...` absorbs the remaining positionals
                if len(args) < len(params) - 1:
                    continue
            elif len(params) != len(args):
                continue
            if all(self._matches(pt, a, m[2]) for (pn, pt, pd, ps), a in zip(params, args) if not ps):
                cands.append(m)
        if not cands:
            raise JlRuntimeError("MethodError: no method of %s matches %r" % (fn.name, [type(a).__name__ for a in args]))
        if len(cands) > 1:
            # most specific = the one with the most annotated parameters
            cands.sort(key=lambda m: -sum(1 for p in m[0] if p[1] is not None))
        params, kwparams, where, body, line, fname = cands[0]
        env = Env(self.globals)
        if params and params[-1][3]:
            env.vars[params[-1][0]] = tuple(args[len(params) - 1:])
            params = params[:-1]
        # bind `where` type variables used as `::Type{T}` / `x::T`
        for (pn, pt, pd, ps), a in zip(params, args):
            if pn is not None:
                env.vars[pn] = a
            if pt is not None:
                if pt[0] == "curly" and pt[1][1] == "Type" and pt[2] and pt[2][0][0] == "id" and pt[2][0][1] in where:
                    env.vars[pt[2][0][1]] = a
                elif pt[0] == "id" and pt[1] in where and is_float(a):
                    env.vars[pt[1]] = typeof_scalar(a)
                elif pt[0] == "curly" and pt[1][1] == "SVector" and isinstance(a, SVec):
                    tv = pt[2][1] if len(pt[2]) > 1 else None
                    if tv is not None and tv[0] == "id" and tv[1] in where:
                        env.vars[tv[1]] = typeof_scalar(a.v[0])
        for pn, pt, pd, ps in kwparams:
            if ps:
                continue
            if pn in kwargs:
                env.vars[pn] = kwargs[pn]
            elif pd is not None:
                env.vars[pn] = self.eval(pd, env)
            else:
                raise JlRuntimeError("UndefKeywordError: %s" % pn)
        try:
            val = self.exec_block(body, env)
        except _Return as r:
            return r.v
        return val

    def call(self, f, args, kwargs=None):
        kwargs = kwargs or {}
        if isinstance(f, Function):
            return self.call_function(f, args, kwargs)
        if isinstance(f, StructType):
            ctor = self.globals.vars.get("#ctor#" + f.name)
            if ctor is not None and not (len(args) == len(f.fields) and len(args) > 4):
                return self.call_function(ctor, args, kwargs)
            if len(args) != len(f.fields):
                raise JlRuntimeError("constructor %s: %d arguments for %d fields" % (f.name, len(args), len(f.fields)))
            return Struct(f.name, f.fields, args)
        if isinstance(f, JlType):
            return convert(f, *args)
        if isinstance(f, TypeApp):
            return self._construct(f, args)
        if f == "SVector":
            return self._construct(TypeApp("SVector", []), args)
        if callable(f):
            return f(*args, **kwargs)
        raise JlRuntimeError("not callable: %r" % (f,))

    def _construct(self, ta, args):
        base = ta.base
        if base in ("Vector",):
            if args and args[0] == "undef":
                n = args[1] if len(args) > 1 else 0
                return JlVector([None] * n, eltype=ta.params[0] if ta.params else None)
            raise JlRuntimeError("Vector constructor form not supported")
        if base == "MVector":
            if args and args[0] == "undef":
                return JlVector([None] * int(ta.params[0]), eltype=ta.params[1], fixed=True)
            raise JlRuntimeError("MVector constructor form not supported")
        if base == "SVector":
            vals = list(args[0].items) if len(args) == 1 and isinstance(args[0], JlVector) else list(args)
            if len(vals) == 1 and isinstance(vals[0], tuple):
                vals = list(vals[0])
            if len(ta.params) >= 1 and int(ta.params[0]) != len(vals):
                raise JlRuntimeError("SVector{%d} from %d values" % (ta.params[0], len(vals)))
            if len(ta.params) >= 2:
                return SVec(convert(ta.params[1], v) for v in vals)
            return SVec(_promote_all(vals))
        raise JlRuntimeError("constructor of %s not supported" % base)

    # ---- statements
    def exec_block(self, block, env):
        val = None
        for st in block[1]:
            val = self.exec(st, env)
        return val

    def exec(self, node, env):
        kind = node[0]
        if kind == "assign":
            val = self.eval(node[2], env) if node[2][0] != "assign" else self.exec(node[2], env)
            self.assign(node[1], val, env)
            return val
        if kind == "opassign":
            cur = self.eval(node[2], env)
            val = jl_binop(node[1], cur, self.eval(node[3], env))
            self.assign(node[2], val, env)
            return val
        if kind == "if":
            for cond, body in node[1]:
                if self.truth(self.eval(cond, env)):
                    return self.exec_block(body, env)
            if node[2] is not None:
                return self.exec_block(node[2], env)
            return None
        if kind == "while":
            while self.truth(self.eval(node[1], env)):
                try:
                    self.exec_block(node[2], Env(env))
                except _Break:
                    break
                except _Continue:
                    continue
            return None
        if kind == "for":
            it = self.eval(node[2], env)
            if isinstance(it, UnitRange):
                seq = range(it.a, it.b + 1)
            elif isinstance(it, FloatRange):
                seq = list(it.vals)
            elif isinstance(it, JlVector):
                seq = it.items
            else:
                raise JlRuntimeError("for: cannot iterate %r" % (it,))
            for x in seq:
                inner = Env(env)
                inner.vars[node[1]] = x
                try:
                    self.exec_block(node[3], inner)
                except _Break:
                    break
                except _Continue:
                    continue
            return None
        if kind == "return":
            raise _Return(self.eval(node[1], env))
        if kind == "block":
            return self.exec_block(node, env)
        if kind == "break":
            raise _Break()
        if kind == "continue":
            raise _Continue()
        if kind == "macro":
            return self.exec_macro(node, env)
        if kind == "function":
            raise JlRuntimeError("nested function definitions are not supported")
        return self.eval(node, env)

    def exec_macro(self, node, env):
        name, args = node[1], node[2]
        if name == "assert":
            if not self.truth(self.eval(args[0], env)):
                raise JlError("AssertionError")
            return None
        if name in ("inbounds", "inline", "noinline", "simd"):
            return self.exec(args[0], env)
        if name == "fastmath":
            self.fastmath += 1
            try:
                return self.exec(args[0], env)
            finally:
                self.fastmath -= 1
        if name == "unpack":      # Parameters.@unpack a, b = obj
            st = args[0]
            if st[0] != "assign":
                raise JlRuntimeError("@unpack: assignment expected")
            obj = self.eval(st[2], env)
            names = st[1][1] if st[1][0] == "tuple" else [st[1]]
            for nm in names:
                env.set(nm[1], obj.fields[nm[1]])
            return None
        if name == "muladd":
            return self.exec(to_muladd(args[0]), env)
        return self.eval(node, env)

    def assign(self, lhs, val, env):
        kind = lhs[0]
        if kind == "id":
            env.set(lhs[1], val)
        elif kind == "typed":
            T = self.eval(lhs[2], env)
            self.assign(lhs[1], convert(T, val) if isinstance(T, JlType) else val, env)
        elif kind == "tuple":
            seq = self._destructure(val)
            if len(seq) < len(lhs[1]):
                raise JlRuntimeError("BoundsError in destructuring: %d targets, %d values" % (len(lhs[1]), len(seq)))
            for tgt, v in zip(lhs[1], seq):
                self.assign(tgt, v, env)
        elif kind == "index":
            obj = self.eval(lhs[1], env)
            idx = self.eval(lhs[2][0], env)
            if not isinstance(obj, JlVector):
                raise JlRuntimeError("setindex! on %r" % (obj,))
            if not 1 <= idx <= len(obj.items):
                raise JlError("BoundsError")
            if isinstance(obj.eltype, JlType) and not isinstance(val, SVec):
                val = convert(obj.eltype, val)
            obj.items[idx - 1] = val
        else:
            raise JlRuntimeError("cannot assign to %s" % kind)

    @staticmethod
    def _destructure(val):
        if isinstance(val, tuple):
            return list(val)
        if isinstance(val, SVec):
            return list(val.v)
        if isinstance(val, JlVector):
            return list(val.items)
        raise JlRuntimeError("cannot destructure %r" % (val,))

    @staticmethod
    def truth(v):
        if isinstance(v, (bool, np.bool_)):
            return bool(v)
        raise JlRuntimeError("TypeError: non-boolean (%r) used in boolean context" % (v,))

    # ---- expressions
    def eval(self, node, env):
        kind = node[0]
        if kind == "num":
            return parse_number(node[1])
        if kind == "id":
            return env.lookup(node[1])
        if kind == "str":
            return node[1]
        if kind == "paren":
            return self.eval(node[1], env)
        if kind == "muladd":
            a, b, c = self.eval(node[1], env), self.eval(node[2], env), self.eval(node[3], env)
            if node[4]:     # dotted: broadcast muladd
                return _broadcast(fma_scalar, [a, b, c])
            return jl_muladd(a, b, c)
        if kind == "op":
            op = node[1]
            vals = [self.eval(x, env) for x in node[2]]
            if op in ("^", ".^"):
                return jl_pow(vals[0], vals[1], self.fastmath > 0)
            dotted = op.startswith(".")
            acc = vals[0]
            for v in vals[1:]:
                acc = jl_binop(op[-1], acc, v, broadcast=dotted)
            return acc
        if kind == "neg":
            v = self.eval(node[1], env)
            if isinstance(v, SVec):
                return SVec(-x for x in v.v)
            return -v
        if kind == "not":
            return not self.truth(self.eval(node[1], env))
        if kind == "and":
            a = self.eval(node[1], env)
            if not self.truth(a):
                return False
            return self.eval(node[2], env)      # value of the last operand (may be non-boolean: `c && error()`)
        if kind == "or":
            a = self.eval(node[1], env)
            if self.truth(a):
                return True
            return self.eval(node[2], env)
        if kind == "cmp":
            vals = [self.eval(node[1][0], env)]
            for k, op in enumerate(node[2]):
                vals.append(self.eval(node[1][k + 1], env))
                if not self._compare(op, vals[-2], vals[-1]):
                    return False
            return True
        if kind == "ternary":
            return self.eval(node[2] if self.truth(self.eval(node[1], env)) else node[3], env)
        if kind == "range":
            a = self.eval(node[1], env)
            b = self.eval(node[3], env)
            if node[2] is None:
                return UnitRange(a, b)
            return FloatRange(a, self.eval(node[2], env), b)
        if kind == "tuple":
            return tuple(self.eval(x, env) for x in node[1])
        if kind == "vect":
            vals = _promote_all([self.eval(x, env) for x in node[1]])
            return JlVector(vals, eltype=typeof_scalar(vals[0]) if vals and not isinstance(vals[0], SVec) else None)
        if kind == "index":
            obj = self.eval(node[1], env)
            idx = self.eval(node[2][0], env)
            return self._getindex(obj, idx)
        if kind == "field":
            obj = self.eval(node[1], env)
            if isinstance(obj, Module):
                return self.globals.lookup(node[2])
            if isinstance(obj, Struct):
                return obj.fields[node[2]]
            return getattr(obj, node[2])
        if kind == "curly":
            base = self.eval(node[1], env)
            params = [self.eval(x, env) for x in node[2]]
            if isinstance(base, StructType):
                return base
            return TypeApp(base, params)
        if kind == "typed":
            return self.eval(node[1], env)
        if kind == "isa":
            val, T = self.eval(node[1], env), self.eval(node[2], env)
            if T in ("Number", "Real"):
                return is_float(val) or isinstance(val, (int, np.integer)) and not isinstance(val, (bool, np.bool_))
            if isinstance(T, JlType):
                return not isinstance(val, (SVec, JlVector)) and typeof_scalar(val) is T
            raise JlRuntimeError("isa: unsupported type %r" % (T,))
        if kind == "comprehension":
            it = self.eval(node[3], env)
            if not isinstance(it, UnitRange):
                raise JlRuntimeError("comprehension over %r" % (it,))
            items = []
            for x in range(it.a, it.b + 1):
                inner = Env(env)
                inner.vars[node[2]] = x
                items.append(self.eval(node[1], inner))
            scalar = items and not isinstance(items[0], (SVec, JlVector))
            return JlVector(items, eltype=typeof_scalar(items[0]) if scalar else None)
        if kind == "call":
            return self.eval_call(node, env)
        if kind == "macro":
            name, args = node[1], node[2]
            if name == "evalpoly":
                vals = [self.eval(a, env) for a in args]
                return evalpoly(vals[0], vals[1:])
            if name == "SVector":
                v = self.eval(args[0], env)
                return SVec(v.items)
            if name in ("fastmath", "inbounds", "muladd"):
                return self.exec_macro(node, env)
            raise JlRuntimeError("macro @%s not supported in expressions" % name)
        if kind in ("assign", "opassign", "if", "while", "for", "block"):
            return self.exec(node, env)
        raise JlRuntimeError("cannot evaluate node %s" % kind)

    @staticmethod
    def _compare(op, a, b):
        if op == "===":
            return (a is b) or (a is not None and b is not None and not isinstance(a, (SVec, JlVector)) and type(a) is type(b) and a == b)
        if op == "!==":
            return not Interp._compare("===", a, b)
        if isinstance(a, SVec) or isinstance(b, SVec):
            eq = isinstance(a, SVec) and isinstance(b, SVec) and all(x == y for x, y in zip(a.v, b.v))
            if op == "==":
                return eq
            if op == "!=":
                return not eq
            raise JlRuntimeError("ordering comparison of SVectors")
        a, b = promote2(a, b)
        return bool({"==": a == b, "!=": a != b, "<": a < b, "<=": a <= b, ">": a > b, ">=": a >= b}[op])

    @staticmethod
    def _getindex(obj, idx):
        if isinstance(obj, SVec):
            return obj.v[idx - 1]
        if isinstance(obj, tuple):
            return obj[idx - 1]
        if isinstance(obj, JlVector):
            if not 1 <= idx <= len(obj.items):
                raise JlError("BoundsError")
            v = obj.items[idx - 1]
            if v is None:
                raise JlError("UndefRefError")
            return v
        if isinstance(obj, FloatRange):
            if not 1 <= idx <= len(obj):
                raise JlError("BoundsError")
            return obj.T(obj.vals[idx - 1])
        if isinstance(obj, UnitRange):
            return obj.a + idx - 1
        raise JlRuntimeError("getindex on %r" % (obj,))

    def eval_call(self, node, env):
        _, fexpr, argx, kwx, dotted = node
        args = []
        for a in argx:
            if a[0] == "splat":
                args.extend(self._destructure(self.eval(a[1], env)))
            else:
                args.append(self.eval(a, env))
        kwargs = {k: self.eval(v, env) for k, v in kwx}
        if self.fastmath and fexpr[0] == "id" and fexpr[1] in ("max", "min") and not dotted:
            return (fast_max if fexpr[1] == "max" else fast_min)(*args)
        f = self.eval(fexpr, env)
        if dotted:
            return _broadcast(lambda *xs: self.call(f, list(xs)), args)
        return self.call(f, args, kwargs)


def _promote_all(vals):
    if any(isinstance(v, SVec) for v in vals):
        return vals
    if any(isinstance(v, F64) for v in vals) or (any(is_float(v) for v in vals) is False and False):
        return [F64(v) for v in vals]
    if any(isinstance(v, F32) for v in vals):
        return [F32(v) for v in vals]
    return vals


def _broadcast(fn, args):
    n = None
    for a in args:
        if isinstance(a, SVec):
            n = len(a)
    if n is None:
        return fn(*args)
    cols = [(a.v if isinstance(a, SVec) else (a,) * n) for a in args]
    return SVec(fn(*xs) for xs in zip(*cols))


def parse_number(text):
    neg = text.startswith("-")
    body = text[1:] if neg else text
    if "f" in body:
        mant, expo = body.split("f")
        # Julia parses Float32 literals by correctly rounding the decimal string once
        v = F32(np.float32(mant + "e" + expo))
    elif "." in body or "e" in body:
        v = F64(float(body))
    else:
        v = int(body)
    return -v if neg else v
