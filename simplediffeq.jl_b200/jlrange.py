"""Julia float-range semantics for the host side (`t0:dt:tf`).

The reference computes its step count and step times from a Julia range:
`_ts = tspan[1]:dt:tspan[2]`, `for i in 2:length(_ts)`, `t = _ts[i-1]`
(src/tsit5/gpuatsit5.jl:92-98, src/rk4/gpurk4.jl:65-74, src/verner/gpuvern7.jl:102-106,
src/verner/gpuvern9.jl:100-104), and `saveat` is usually a range too.  In a Julia deployment the
shim passes `collect(t0:dt:tf)` straight to the C ABI (julia/SimpleDiffEqCUDA.jl); this module is
the same thing for the Python host layer, restating Julia Base's published algorithm
(base/twiceprecision.jl: `(:)(start::T, step::T, stop::T) where T<:IEEEFloat`, `rat`,
`floatrange`, `steprangelen_hp`; [EXT] assumption A7 of SURVEY.md):

* if start/step/stop lift to small rationals, the length is computed in integers and element i is
  the exactly rounded `(start_n + (i-1)*step_n)/den` (Float64: TwicePrecision evaluation, which
  agrees with correct rounding except within ~2^-100 of a tie; Float32: evaluated in Float64
  and rounded once);
* otherwise `len = round((stop-start)/step)+1` (minus one on overshoot) and element i is
  `start + (i-1)*step` with the product and the sum each rounded (Float32: in Float64, then
  rounded to Float32).
"""
from fractions import Fraction
import math

import numpy as np


def _maxintfloat(T):
    return {np.float64: 9007199254740992, np.float32: 16777216, np.float16: 2048}[T]


def _narrow(T):
    return {np.float64: np.float32, np.float32: np.float16, np.float16: np.float16}[T]


def _rat(x, T):
    """Base.rat: continued-fraction rational approximation in T arithmetic."""
    y = T(x)
    a = d = 1
    b = c = 0
    m = _maxintfloat(_narrow(T))
    while abs(y) <= m:
        f = int(math.trunc(float(y)))
        y = T(y - T(f))
        a, c = f * a + c, a
        b, d = f * b + d, b
        if not max(abs(a), abs(b)) <= m:
            return c, d
        if b != 0 and T(T(a) / T(b)) == T(x):
            break
        if y == 0:
            # inv(0) = Inf -> loop ends with abs(y) > m
            break
        y = T(T(1) / y)
    return a, b


def _isbetween(a, x, b):
    return (a <= x <= b) or (b <= x <= a)


def _div_trunc(a, b):
    q = abs(a) // abs(b)
    return q if (a >= 0) == (b >= 0) else -q


class JuliaRange:
    """start:step:stop for Float64 or Float32, with Julia's length and element values."""

    def __init__(self, start, step, stop, dtype=np.float64):
        T = np.dtype(dtype).type
        assert T in (np.float64, np.float32)
        self.T = T
        start, step, stop = T(start), T(step), T(stop)
        if step == 0:
            raise ValueError("range step cannot be zero")
        self.start, self.step, self.stop = start, step, stop
        self.rational = None
        step_n, step_d = _rat(step, T)
        if step_d != 0 and T(step_n / step_d) == step:
            start_n, start_d = _rat(start, T)
            stop_n, stop_d = _rat(stop, T)
            if (start_d != 0 and stop_d != 0 and T(start_n / start_d) == start
                    and T(stop_n / stop_d) == stop):
                den = start_d * step_d // math.gcd(start_d, step_d)
                m = _maxintfloat(T)
                if (den != 0 and abs(float(T(start * T(den)))) <= m and abs(float(T(step * T(den)))) <= m
                        and den % start_d == 0 and den % step_d == 0):
                    start_n = int(round(float(T(start * T(den)))))
                    step_n = int(round(float(T(step * T(den)))))
                    ln = max(0, _div_trunc(den * stop_n - stop_d * start_n, step_n * stop_d) + 1)
                    last = T(start + T(T(ln - 1) * step))
                    nxt = T(start + T(T(ln) * step))
                    if _isbetween(start, last, T(stop + T(step / T(2)))) and not _isbetween(start, nxt, stop):
                        self.rational = (start_n, step_n, den)
                        self.len = ln
                        return
        lf = T(T(stop - start) / step)
        if lf < 0:
            ln = 0
        elif lf == 0:
            ln = 1
        else:
            ln = int(round(float(lf))) + 1
            stopp = T(start + T(T(ln - 1) * step))
            ln -= int(start < stop < stopp) + int(start > stop > stopp)
        self.len = ln

    def __len__(self):
        return self.len

    def collect(self):
        """All elements as a numpy array of the range's dtype."""
        n = self.len
        T = self.T
        if n == 0:
            return np.empty(0, dtype=T)
        if self.rational is not None:
            start_n, step_n, den = self.rational
            if T is np.float64:
                out = np.empty(n, dtype=np.float64)
                # exactly rounded rationals; vectorised fast path when everything is exactly
                # representable (|num| < 2^53 and den < 2^53): fl(num/den) is correctly rounded
                num = start_n + step_n * np.arange(n, dtype=object)
                if max(abs(start_n), abs(start_n + step_n * (n - 1))) < 2 ** 53 and den < 2 ** 53:
                    out[:] = np.asarray(num, dtype=np.float64) / float(den)
                else:
                    for i in range(n):
                        out[i] = float(Fraction(int(num[i]), den))
                return out
            # Float32: ref and step kept in Float64 (StepRangeLen{Float32,Float64,Float64}),
            # offset = index of the smallest-magnitude element
            imin = 1
            if n >= 2 and step_n != 0:
                imin = min(max(int(round(-start_n / step_n + 1)), 1), n)
            ref = (start_n + (imin - 1) * step_n) / den
            stp = step_n / den
            idx = np.arange(1, n + 1, dtype=np.float64) - imin
            return (ref + idx * stp).astype(np.float32)
        if T is np.float64:
            idx = np.arange(n, dtype=np.float64)
            return np.float64(self.start) + idx * np.float64(self.step)
        idx = np.arange(n, dtype=np.float64)
        return (np.float64(self.start) + idx * np.float64(self.step)).astype(np.float32)


def jl_range(start, step, stop, dtype=np.float64):
    """collect(start:step:stop) with Julia semantics."""
    return JuliaRange(start, step, stop, dtype).collect()
