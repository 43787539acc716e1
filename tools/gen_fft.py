#!/usr/bin/env python3
"""Generate straight-line register FFT bodies for the fused log-mel kernel.

Writes ``adt_str_b200/csrc/fft_gen.cuh`` with

* ``rdft64<V>(const V (&x)[64], V (&re)[33], V (&im)[33])`` - real-input
  64-point DFT, outputs bins 0..32 (im[0] = im[32] = 0), and
* ``cdft32<V>(V (&re)[32], V (&im)[32])`` - complex 32-point DFT, in place,
  natural order in and out.

``V`` is ``float`` or ``float2``: with ``float2`` every operation is one packed
sm_100a instruction (FADD2 / FMUL2 / FFMA2) working on two independent transforms
at once - the log-mel kernel runs two frames per warp that way.  Negations,
constant broadcasts and half swaps are operand modifiers in SASS, so the packed
body has exactly the instruction count of the scalar one.

Both are radix-2 decimation-in-time recursions executed *symbolically*: every
value is a signed reference to a C variable (or a known zero), so negations,
multiplications by 1 / -1 / -i and arithmetic on zeros cost nothing, twiddles
are literal constants, and a butterfly with a general twiddle is 6 FMAs:

    X1 = E + w*O      (4 FMA: two chained per component)
    X2 = 2E - X1      (2 FMA)

For real input the recursion keeps only bins 0..N/2 of every sub-transform and
uses  X[N/2-k] = conj(E[k] - w_k O[k]),  so it costs half a complex transform.

Run from the repo root:  python tools/gen_fft.py
"""
from __future__ import annotations

import math
import os
import sys


class Emitter:
    def __init__(self):
        self.lines = []
        self.n = 0
        self.ops = 0

    def new(self, expr: str) -> "Val":
        name = f"t{self.n}"
        self.n += 1
        self.ops += 1
        self.lines.append(f"    const V {name} = {expr};")
        return Val(name, 1)


class Val:
    """sign * variable, or zero (name None)."""
    __slots__ = ("name", "sign")

    def __init__(self, name, sign=1):
        self.name, self.sign = name, sign

    @property
    def zero(self):
        return self.name is None

    def neg(self):
        return Val(self.name, -self.sign)


ZERO = Val(None)


def lit(c: float) -> str:
    return f"{c:.9e}f"


def add(em: Emitter, a: Val, b: Val) -> Val:
    if a.zero:
        return b
    if b.zero:
        return a
    if a.sign > 0 and b.sign > 0:
        return em.new(f"vadd({a.name}, {b.name})")
    if a.sign > 0:
        return em.new(f"vsub({a.name}, {b.name})")
    if b.sign > 0:
        return em.new(f"vsub({b.name}, {a.name})")
    return em.new(f"vadd({a.name}, {b.name})").neg()


def sub(em, a, b):
    return add(em, a, b.neg())


def fma(em: Emitter, c: float, x: Val, y: Val) -> Val:
    """c*x + y with a literal c."""
    if x.zero or c == 0.0:
        return y
    c = c * x.sign
    if y.zero:
        return em.new(f"vmul({lit(c)}, {x.name})")
    if y.sign > 0:
        return em.new(f"vfma({lit(c)}, {x.name}, {y.name})")
    return em.new(f"vfma({lit(-c)}, {x.name}, {y.name})").neg()


def butterfly(em, E, O, wr, wi, need2=True):
    """(E + w O, E - w O) for complex symbolic E, O and literal w."""
    Er, Ei = E
    Or, Oi = O
    eps = 1e-12
    if abs(wr - 1) < eps and abs(wi) < eps:          # w = 1
        return (add(em, Er, Or), add(em, Ei, Oi)), (sub(em, Er, Or), sub(em, Ei, Oi))
    if abs(wr) < eps and abs(wi + 1) < eps:          # w = -i : wO = (Oi, -Or)
        return (add(em, Er, Oi), sub(em, Ei, Or)), (sub(em, Er, Oi), add(em, Ei, Or))
    x1r = fma(em, -wi, Oi, fma(em, wr, Or, Er))
    x1i = fma(em, wi, Or, fma(em, wr, Oi, Ei))
    if not need2:
        return (x1r, x1i), None
    # X2 = 2E - X1  (falls back to E - wO when E is a known zero)
    if Er.zero:
        x2r = x1r.neg()
    else:
        x2r = fma(em, 2.0, Er, x1r.neg())
    if Ei.zero:
        x2i = x1i.neg()
    else:
        x2i = fma(em, 2.0, Ei, x1i.neg())
    return (x1r, x1i), (x2r, x2i)


def cfft(em, xs):
    n = len(xs)
    if n == 1:
        return xs
    E, O = cfft(em, xs[0::2]), cfft(em, xs[1::2])
    out = [None] * n
    for k in range(n // 2):
        a = -2.0 * math.pi * k / n
        out[k], out[k + n // 2] = butterfly(em, E[k], O[k], math.cos(a), math.sin(a))
    return out


def rfft(em, xs):
    """real symbolic inputs -> bins 0..n/2."""
    n = len(xs)
    if n == 1:
        return [(xs[0], ZERO)]
    E, O = rfft(em, xs[0::2]), rfft(em, xs[1::2])
    out = [None] * (n // 2 + 1)
    for k in range(n // 4 + 1):
        a = -2.0 * math.pi * k / n
        mirror = n // 2 - k != k
        x1, x2 = butterfly(em, E[k], O[k], math.cos(a), math.sin(a), need2=mirror)
        out[k] = x1
        if mirror:
            out[n // 2 - k] = (x2[0], x2[1].neg())   # conj(E - wO)
    return out


def store(v: Val) -> str:
    if v.zero:
        return "vzero<V>()"
    return v.name if v.sign > 0 else f"vneg({v.name})"


def gen_rdft64() -> tuple[str, int]:
    em = Emitter()
    xs = [Val(f"x[{i}]") for i in range(64)]
    X = rfft(em, xs)
    body = list(em.lines)
    for k in range(33):
        body.append(f"    re[{k}] = {store(X[k][0])};")
        body.append(f"    im[{k}] = {store(X[k][1])};")
    src = ("template <typename V>\n__device__ __forceinline__ void rdft64(const V (&x)[64], V (&re)[33], V (&im)[33]) {\n"
           + "\n".join(body) + "\n}\n")
    return src, em.ops


def gen_cdft32() -> tuple[str, int]:
    em = Emitter()
    head = []
    xs = []
    for i in range(32):
        head.append(f"    const V xr{i} = re[{i}], xi{i} = im[{i}];")
        xs.append((Val(f"xr{i}"), Val(f"xi{i}")))
    X = cfft(em, xs)
    body = head + list(em.lines)
    for k in range(32):
        body.append(f"    re[{k}] = {store(X[k][0])};")
        body.append(f"    im[{k}] = {store(X[k][1])};")
    src = ("template <typename V>\n__device__ __forceinline__ void cdft32(V (&re)[32], V (&im)[32]) {\n"
           + "\n".join(body) + "\n}\n")
    return src, em.ops


PRELUDE = r"""
// Element operations: float, or float2 = two independent transforms per thread in packed
// FADD2 / FMUL2 / FFMA2 instructions (sm_100a).  Negation and constant broadcast fold into
// SASS operand modifiers.
template <typename V> __device__ __forceinline__ V vzero();
template <> __device__ __forceinline__ float vzero<float>() { return 0.0f; }
template <> __device__ __forceinline__ float2 vzero<float2>() { return make_float2(0.0f, 0.0f); }
__device__ __forceinline__ float vneg(float a) { return -a; }
__device__ __forceinline__ float vadd(float a, float b) { return a + b; }
__device__ __forceinline__ float vsub(float a, float b) { return a - b; }
__device__ __forceinline__ float vmul(float c, float a) { return c * a; }
__device__ __forceinline__ float vfma(float c, float a, float b) { return fmaf(c, a, b); }
__device__ __forceinline__ float2 vneg(float2 a) { return make_float2(-a.x, -a.y); }
__device__ __forceinline__ float2 vadd(float2 a, float2 b) { return __fadd2_rn(a, b); }
__device__ __forceinline__ float2 vsub(float2 a, float2 b) { return __fadd2_rn(a, make_float2(-b.x, -b.y)); }
__device__ __forceinline__ float2 vmul(float c, float2 a) { return __fmul2_rn(a, make_float2(c, c)); }
__device__ __forceinline__ float2 vfma(float c, float2 a, float2 b) { return __ffma2_rn(a, make_float2(c, c), b); }
"""


def main():
    out = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "adt_str_b200", "csrc", "fft_gen.cuh")
    r_src, r_ops = gen_rdft64()
    c_src, c_ops = gen_cdft32()
    with open(out, "w") as f:
        f.write("// GENERATED by tools/gen_fft.py - do not edit.\n"
                f"// rdft64: {r_ops} float ops per thread, cdft32: {c_ops} float ops per thread.\n"
                "#pragma once\n\nnamespace adtfe {\n" + PRELUDE + "\n" + r_src + "\n" + c_src + "\n}  // namespace adtfe\n")
    print(f"wrote {os.path.normpath(out)}: rdft64 {r_ops} ops, cdft32 {c_ops} ops", file=sys.stderr)


if __name__ == "__main__":
    main()
