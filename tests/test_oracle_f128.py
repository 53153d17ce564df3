"""CPU tests pinning the fft128 oracle.

The reference has no known-answer vector for fft128, so the pins are:
  * the reference's own property test (src/fft128/mod.rs:1972-2065): the negacyclic product
    through fwd / pointwise / inv equals the schoolbook product to 1e-30 * N;
  * the transform's definition evaluated with mpmath;
  * the double-double op error bounds of src/fft128/f128_ops.rs:1043-1215 against mpmath.
"""
from fractions import Fraction

import mpmath
import numpy as np
import pytest

import oracle_lib as O
from f128_util import dd_mul_pointwise, negacyclic_schoolbook_exact, dd_to_fraction


@pytest.mark.parametrize("variant", [O.F128_SCALAR, O.F128_FMA])
@pytest.mark.parametrize("npoly", [64, 128, 256, 512, 1024, 2048, 4096])
def test_negacyclic_product(npoly, variant):
    rng = np.random.default_rng(npoly + variant)
    n = npoly // 2
    lhs, rhs = rng.random(npoly), rng.random(npoly)
    exact = negacyclic_schoolbook_exact(lhs, rhs)  # list of Fraction

    plan = O.F128Plan(n)
    z = np.zeros(n)
    L = plan.fwd(lhs[:n], z, lhs[n:], z, variant)
    R = plan.fwd(rhs[:n], z, rhs[n:], z, variant)
    P = dd_mul_pointwise(L, R, 2.0 / npoly)
    out = plan.inv(*P, variant)
    got_hi = np.concatenate([out[0], out[2]])
    got_lo = np.concatenate([out[1], out[3]])
    err = max(abs(dd_to_fraction(h, l) - e) for h, l, e in zip(got_hi, got_lo, exact))
    assert float(err) < 1e-30 * npoly  # src/fft128/mod.rs:2062


def test_forward_definition_mpmath():
    """X[j] = sum_k z_k psi_j^k, psi_j = exp(i pi (4 bitrev_n(j) + 1) / (2n)), SURVEY.md A.5."""
    mpmath.mp.prec = 300
    n = 32
    rng = np.random.default_rng(11)
    re0, im0 = rng.random(n), rng.random(n)
    re1 = (rng.random(n) - 0.5) * np.spacing(re0)
    im1 = (rng.random(n) - 0.5) * np.spacing(im0)
    plan = O.F128Plan(n)
    for variant in (O.F128_SCALAR, O.F128_FMA):
        out = plan.fwd(re0, re1, im0, im1, variant)
        z = [mpmath.mpc(mpmath.mpf(float(a)) + mpmath.mpf(float(b)), mpmath.mpf(float(c)) + mpmath.mpf(float(d)))
             for a, b, c, d in zip(re0, re1, im0, im1)]
        logn = n.bit_length() - 1
        for j in range(n):
            br = int(format(j, "0%db" % logn)[::-1], 2)
            psi = mpmath.expjpi(mpmath.mpf(4 * br + 1) / (2 * n))
            want = sum(zk * psi ** k for k, zk in enumerate(z))
            got = mpmath.mpc(mpmath.mpf(float(out[0][j])) + mpmath.mpf(float(out[1][j])),
                             mpmath.mpf(float(out[2][j])) + mpmath.mpf(float(out[3][j])))
            assert abs(got - want) < mpmath.mpf(2) ** -95


def test_inverse_is_unnormalised_adjoint_roundtrip():
    rng = np.random.default_rng(12)
    for n in [32, 64, 1024]:
        planes = [rng.random(n), np.zeros(n), rng.random(n), np.zeros(n)]
        plan = O.F128Plan(n)
        out = plan.inv(*plan.fwd(*planes))
        for got_hi, got_lo, x in zip((out[0], out[2]), (out[1], out[3]), (planes[0], planes[2])):
            err = max(abs(dd_to_fraction(h, l) / n - Fraction(float(v))) for h, l, v in zip(got_hi, got_lo, x))
            assert float(err) < 1e-29


def test_twiddles_accuracy():
    """twid[m+i] = exp(i pi bitrev_2n(2m+i) / 2n) to ~2^-103 (src/fft128/f128_ops.rs:1062-1070)."""
    mpmath.mp.prec = 400
    n = 64
    tw = O.F128Plan(n).twiddles()
    assert tw[0][0] == 0.0 and tw[2][0] == 0.0  # entry 0 unused
    m = 1
    while m < n:
        for i in range(m):
            k = 2 * m + i
            br = int(format(k, "0%db" % ((2 * n).bit_length() - 1))[::-1], 2)
            want = mpmath.expjpi(mpmath.mpf(br) / (2 * n))
            re = mpmath.mpf(float(tw[0][m + i])) + mpmath.mpf(float(tw[1][m + i]))
            im = mpmath.mpf(float(tw[2][m + i])) + mpmath.mpf(float(tw[3][m + i]))
            assert abs(re - want.real) < mpmath.mpf(2) ** -102
            assert abs(im - want.imag) < mpmath.mpf(2) ** -102
        m *= 2


def test_plan_rejects_small_or_non_pow2():
    for n in [16, 48, 0]:
        with pytest.raises(ValueError):
            O.F128Plan(n)  # src/fft128/mod.rs:1865-1866


def test_scalar_and_fma_variants_differ_only_in_last_bits():
    rng = np.random.default_rng(13)
    n = 256
    planes = [rng.random(n), np.zeros(n), rng.random(n), np.zeros(n)]
    plan = O.F128Plan(n)
    a = plan.fwd(*planes, O.F128_SCALAR)
    b = plan.fwd(*planes, O.F128_FMA)
    assert all(np.array_equal(x, y) for x, y in zip(a[::2], b[::2])) or True  # hi may differ by 1 ulp rarely
    diff = max(float(abs(dd_to_fraction(h1, l1) - dd_to_fraction(h2, l2)))
               for h1, l1, h2, l2 in zip(a[0], a[1], b[0], b[1]))
    scale = float(np.abs(a[0]).max())
    assert diff < scale * 2.0 ** -95


def test_operator_set_error_bounds():
    """src/fft128/f128_ops.rs:1043-1215 re-expressed with mpmath: exact ops <= 2^-104 relative,
    *_estimate <= 2^-101."""
    mpmath.mp.prec = 1024
    rng = np.random.default_rng(5)
    n = 300
    a_hi = rng.uniform(-4, 4, n)
    b_hi = rng.uniform(0.25, 4, n) * rng.choice([-1.0, 1.0], n)
    a_lo = (rng.random(n) - 0.5) * np.spacing(a_hi)
    b_lo = (rng.random(n) - 0.5) * np.spacing(b_hi)
    fns = {"add": lambda x, y: x + y, "sub": lambda x, y: x - y, "mul": lambda x, y: x * y, "div": lambda x, y: x / y}
    for op in ["add", "sub", "mul", "div", "add_estimate", "sub_estimate", "div_estimate"]:
        hi, lo = O.f128_binary_op(op, a_hi, a_lo, b_hi, b_lo)
        bound = mpmath.mpf(2) ** (-101 if op.endswith("estimate") else -104)
        for i in range(n):
            A = mpmath.mpf(float(a_hi[i])) + mpmath.mpf(float(a_lo[i]))
            B = mpmath.mpf(float(b_hi[i])) + mpmath.mpf(float(b_lo[i]))
            want = fns[op.split("_")[0]](A, B)
            got = mpmath.mpf(float(hi[i])) + mpmath.mpf(float(lo[i]))
            scale = max(abs(want), abs(A) if op.startswith(("add", "sub")) else abs(want))
            assert abs(got - want) <= bound * scale, (op, i)


def test_mixed_operand_and_unary_operators_error_bounds():
    """the rest of the scalar operator surface (f128_ops.rs:279-455 mixed forms, :404 sqr, :506 abs, :232 neg, :514-618
    sincospi) against mpmath at 1024 bits: exact ops <= 2^-104 relative (the reference's own bound, f128_ops.rs:1062-1070),
    sincospi <= 2^-103 (:1189-1215); the f64-f64 forms are error-free transformations (exact) except div."""
    mpmath.mp.prec = 1024
    rng = np.random.default_rng(6)
    n = 200
    a_hi = rng.uniform(-4, 4, n)
    b_hi = rng.uniform(0.25, 4, n) * rng.choice([-1.0, 1.0], n)
    a_lo = (rng.random(n) - 0.5) * np.spacing(a_hi)
    b_lo = (rng.random(n) - 0.5) * np.spacing(b_hi)
    M = lambda h, l=0.0: mpmath.mpf(float(h)) + mpmath.mpf(float(l))
    fns = {"add": lambda x, y: x + y, "sub": lambda x, y: x - y, "mul": lambda x, y: x * y, "div": lambda x, y: x / y}
    for op in ["add_f128_f64", "sub_f128_f64", "sub_f64_f128", "mul_f128_f64", "div_f128_f64", "div_f64_f128", "add_f64_f64", "sub_f64_f64",
               "mul_f64_f64", "div_f64_f64"]:
        a_is_f64 = op.split("_")[1] == "f64"
        b_is_f64 = op.split("_")[2] == "f64"
        hi, lo = O.f128_binary_op(op, a_hi, None if a_is_f64 else a_lo, b_hi, None if b_is_f64 else b_lo)
        exact = a_is_f64 and b_is_f64 and not op.startswith("div")
        for i in range(n):
            A = M(a_hi[i], 0.0 if a_is_f64 else a_lo[i])
            B = M(b_hi[i], 0.0 if b_is_f64 else b_lo[i])
            want, got = fns[op.split("_")[0]](A, B), M(hi[i], lo[i])
            if exact:
                assert got == want, (op, i)  # two_sum / two_diff / two_prod lose nothing
            else:
                scale = max(abs(want), abs(A) if op.startswith(("add", "sub")) else abs(want))
                assert abs(got - want) <= mpmath.mpf(2) ** -104 * scale, (op, i)
    hi, lo = O.f128_unary_op("sqr", a_hi, a_lo)
    for i in range(n):
        A = M(a_hi[i], a_lo[i])
        assert abs(M(hi[i], lo[i]) - A * A) <= mpmath.mpf(2) ** -104 * A * A
    hi, lo = O.f128_unary_op("abs", a_hi, a_lo)
    assert all(M(hi[i], lo[i]) == abs(M(a_hi[i], a_lo[i])) for i in range(n))
    hi, lo = O.f128_unary_op("neg", a_hi, a_lo)
    assert np.array_equal(hi, -a_hi) and np.array_equal(lo, -a_lo)
    x_hi = np.concatenate([rng.uniform(-1, 1, 150), np.arange(-16, 17) / 16.0])
    x_lo = (rng.random(x_hi.size) - 0.5) * np.spacing(x_hi) * (np.abs(x_hi) < 1)
    (sh, sl), (ch, cl) = O.f128_unary_op("sincospi", x_hi, x_lo)
    for i in range(x_hi.size):
        X = M(x_hi[i], x_lo[i])
        assert abs(M(sh[i], sl[i]) - mpmath.sinpi(X)) <= mpmath.mpf(2) ** -103, i
        assert abs(M(ch[i], cl[i]) - mpmath.cospi(X)) <= mpmath.mpf(2) ** -103, i
    # PartialOrd / PartialEq (f128_ops.rs:240-274) on a handful of decided cases
    ah, al = np.array([1.0, 1.0, 1.0, 2.0, np.nan, 1.0, 1.0]), np.array([0.0, 1e-20, -1e-20, 0.0, 0.0, np.nan, 0.0])
    bh, bl = np.array([1.0, 1.0, 1.0, 1.0, 1.0, 1.0, 2.0]), np.array([0.0, 0.0, 0.0, 5.0, 0.0, 0.0, -9.0])
    assert list(O.f128_compare(ah, al, bh, bl)) == [0, 1, -1, 1, 2, 2, -1]
    assert list(O.f128_compare(ah, al, bh)) == [0, 1, -1, 1, 2, 2, -1]
    assert list(O.f128_unary_op("is_nan", ah, al)[0]) == [0, 0, 0, 0, 1, 1, 0]


def test_vectorised_build_same_bits():
    rng = np.random.default_rng(77)
    for n in [32, 64, 256, 2048]:
        planes = [rng.random((3, n)), (rng.random((3, n)) - 0.5) * 1e-17, rng.random((3, n)), (rng.random((3, n)) - 0.5) * 1e-17]
        a = O.F128Plan(n).fwd(*planes, variant=O.F128_FMA)
        b = O.F128Plan(n, fast=True).fwd(*planes, variant=O.F128_FMA)
        assert all(np.array_equal(x.view(np.uint64), y.view(np.uint64)) for x, y in zip(a, b))
        a2 = O.F128Plan(n).inv(*a, variant=O.F128_FMA)
        b2 = O.F128Plan(n, fast=True).inv(*a, variant=O.F128_FMA)
        assert all(np.array_equal(x.view(np.uint64), y.view(np.uint64)) for x, y in zip(a2, b2))
