"""Helpers for fft128 tests: exact double-double <-> rational conversion, an exact schoolbook
negacyclic product, and the pointwise product the reference's tests use between fwd and inv.
Test infrastructure only."""
from fractions import Fraction

import numpy as np


def dd_to_fraction(hi, lo):
    return Fraction(float(hi)) + Fraction(float(lo))


def _two_sum(a, b):
    s = a + b
    bb = s - a
    return s, (a - (s - bb)) + (b - bb)


def _two_diff(a, b):
    s = a - b
    bb = s - a
    return s, (a - (s - bb)) - (b + bb)


def _quick_two_sum(a, b):
    s = a + b
    return s, b - (s - a)


def _split(a):
    t = a * 134217729.0
    hi = t - (t - a)
    return hi, a - hi


def _two_prod(a, b):
    # Dekker/Veltkamp product: exact like fma(a, b, -p) for values away from over/underflow
    p = a * b
    ah, al = _split(a)
    bh, bl = _split(b)
    return p, ((ah * bh - p) + ah * bl + al * bh) + al * bl


def dd_add(a, b):
    s, e = _two_sum(a[0], b[0])
    e = e + (a[1] + b[1])
    return _quick_two_sum(s, e)


def dd_sub(a, b):
    s, e = _two_diff(a[0], b[0])
    e = e + a[1]
    e = e - b[1]
    return _quick_two_sum(s, e)


def dd_mul(a, b):
    p, e = _two_prod(a[0], b[0])
    e = e + (a[0] * b[1] + a[1] * b[0])
    return _quick_two_sum(p, e)


def dd_mul_pointwise(L, R, factor):
    """(L * R) * factor on planar (re0, re1, im0, im1) numpy arrays; factor a power of two.
    Follows the loop at src/fft128/mod.rs:2033-2047 (vectorised over the arrays)."""
    lre, lim = (L[0], L[1]), (L[2], L[3])
    rre, rim = (R[0], R[1]), (R[2], R[3])
    rr, ri = dd_mul(lre, rre), dd_mul(lre, rim)
    ir, ii = dd_mul(lim, rre), dd_mul(lim, rim)
    pre = dd_sub(rr, ii)
    pim = dd_add(ir, ri)
    return [pre[0] * factor, pre[1] * factor, pim[0] * factor, pim[1] * factor]


def negacyclic_schoolbook_exact(lhs, rhs):
    """Exact negacyclic product of two float64 coefficient vectors (values in [0, 1) that are
    multiples of 2^-53), as Fractions.  Integer limbs keep every partial sum inside int64."""
    n = len(lhs)
    a = [int(Fraction(float(v)) * (1 << 53)) for v in lhs]
    b = [int(Fraction(float(v)) * (1 << 53)) for v in rhs]
    LB = 18
    mask = (1 << LB) - 1
    la = [np.array([(v >> (LB * k)) & mask for v in a], dtype=np.int64) for k in range(3)]
    lb = [np.array([(v >> (LB * k)) & mask for v in b], dtype=np.int64) for k in range(3)]
    full = [0] * (2 * n - 1)
    for i in range(3):
        for j in range(3):
            c = np.convolve(la[i], lb[j])
            sh = LB * (i + j)
            for t, v in enumerate(c.tolist()):
                full[t] += v << sh
    full.append(0)
    den = 1 << 106
    return [Fraction(full[i] - full[i + n], den) for i in range(n)]
