"""CPU checks of the oracle's statement of the caller-side steps around the c64 transform (oracle/poly_oracle.c: fold,
integer / torus conversion, negacyclic twist, untwist, rounding).  The reference does not contain these steps (they live
in its caller), so they are pinned here by exact integer arithmetic: the pipeline fold -> twist -> reference fwd ->
element-wise product -> reference inv -> untwist -> round must return the schoolbook negacyclic product exactly."""
import numpy as np
import pytest

import oracle_lib as O


def negacyclic_schoolbook(a, b):
    """exact product of integer polynomials modulo X^N + 1 (python integers)"""
    npoly = len(a)
    full = np.convolve(a.astype(object), b.astype(object))
    out = full[:npoly].copy()
    out[: npoly - 1] -= full[npoly:]
    return out


@pytest.mark.parametrize("n,base_n,k", [(32, 32, 1), (64, 32, 2), (512, 256, 3), (2048, 256, 2), (2048, 32, 1)])
def test_integer_product_is_exact(n, base_n, k):
    rng = np.random.default_rng(n + k)
    batch = 3
    a = rng.integers(-(1 << 18), 1 << 18, size=(batch, k, 2 * n))
    bp = rng.integers(-(1 << 9), 1 << 9, size=(k, 2 * n))
    plan = O.UnorderedPlan(n, O.DIF16 if base_n >= 32 and base_n != 32 else O.DIF4, base_n)
    fb = plan.fwd(O.poly_fold_twist(bp))
    got = O.poly_mul(plan, a, fb, threads=2)
    for r in range(batch):
        want = sum(negacyclic_schoolbook(a[r, j], bp[j]) for j in range(k))
        assert np.array_equal(got[r], want.astype(np.int64))
    # accumulate adds modulo 2^64
    acc0 = rng.integers(-(1 << 62), 1 << 62, size=(batch, 2 * n))
    got2 = O.poly_mul(plan, a, fb, acc=acc0)
    assert np.array_equal(got2.view(np.uint64), (acc0.view(np.uint64) + got.view(np.uint64)))


def test_torus_product_matches_integer_arithmetic_modulo_2_64():
    """torus mode: a uniformly random u64 polynomial (a ciphertext mask) times a small integer polynomial (a key / a
    decomposed digit), product taken modulo 2^64: the f64 pipeline keeps the top ~53 - log2(N * |b|) bits."""
    n, npoly = 512, 1024
    rng = np.random.default_rng(5)
    a = rng.integers(-(1 << 63), 1 << 63, size=(2, 1, npoly), dtype=np.int64)
    bp = rng.integers(-4, 5, size=(1, npoly))
    plan = O.UnorderedPlan(n, O.DIF16, 256)
    fb = plan.fwd(O.poly_fold_twist(bp))  # the small polynomial goes in as plain integers
    got = O.poly_mul(plan, a, fb, torus=True)
    for r in range(2):
        want = negacyclic_schoolbook(a[r, 0], bp[0])  # exact, python integers
        diff = np.array([(int(g) - int(w)) % (1 << 64) for g, w in zip(got[r], want)], dtype=object)
        diff = np.array([d - (1 << 64) if d >= (1 << 63) else d for d in diff], dtype=object)
        assert max(abs(int(d)) for d in diff) < (1 << 27)  # ~ 2^64 * N * |b| * 2^-53


def test_twist_tables():
    for n in (32, 256, 4096):
        tw, un = O.poly_twist_tables(n)
        j = np.arange(n)
        assert np.abs(tw - np.exp(1j * np.pi * j / (2 * n))).max() < 3e-16
        assert tw[0] == 1.0 and un[0] == 1.0 / n
        assert np.array_equal(un, np.conj(tw) / n)  # exact: n is a power of two
        # j = n / 2 is e^{i pi / 4}: the reference's sincospi64 is within one ulp there (cos ...476, sin ...475)
        assert abs(tw[n // 2].real - tw[n // 2].imag) <= 2.0 ** -52


def test_rounding_semantics():
    """integer mode: f64::round (half away from zero) then a saturating cast; torus mode: fractional part x 2^64 modulo 2^64."""
    n = 32
    un = np.ones(n, np.complex128)  # identity untwist so the conversions see the values below unchanged
    vals = np.array([0.5, -0.5, 1.5, 2.5, -2.5, 0.49999999999999994, 1e30, -1e30, np.nan, 123456789.0, -7.0, 2.0 ** 62, -(2.0 ** 63)] + [0.0] * (n - 13))
    z = vals + 1j * vals[::-1]
    out = np.zeros(2 * n, np.int64)
    O.lib().orc_poly_untwist_round(n, 0, 0, z.ctypes.data, un.ctypes.data, out.ctypes.data)
    assert list(out[:13]) == [1, -1, 2, 3, -3, 0, 2 ** 63 - 1, -(2 ** 63), 0, 123456789, -7, 2 ** 62, -(2 ** 63)]
    tv = np.array([0.25, -0.25, 0.5, -0.5, 1.25, 1e30, np.nan, 2.0 ** -64, 3.0] + [0.0] * (n - 9))
    z = tv + 0j
    O.lib().orc_poly_untwist_round(n, 1, 0, z.ctypes.data, un.ctypes.data, out.ctypes.data)
    u = out.view(np.uint64)
    # 0.5 - round(0.5) = -0.5 -> -2^63 = 2^63 (mod 2^64); -0.5 -> +0.5 -> 2^63 as well
    assert [int(x) for x in u[:9]] == [1 << 62, (1 << 64) - (1 << 62), 1 << 63, 1 << 63, 1 << 62, 0, 0, 1, 0]


def test_fold_and_conversion():
    n = 32
    rng = np.random.default_rng(1)
    poly = rng.integers(-(1 << 40), 1 << 40, size=2 * n)
    tw, _ = O.poly_twist_tables(n)
    z = O.poly_fold_twist(poly)
    want = O.c64_pointwise((poly[:n] + 1j * poly[n:]).astype(np.complex128), tw)  # num_complex product, no FMA
    assert np.array_equal(z.view(np.uint64), want.view(np.uint64))
    zt = O.poly_fold_twist(poly, torus=True)
    want_t = O.c64_pointwise(((poly[:n] + 1j * poly[n:]) * 2.0 ** -64).astype(np.complex128), tw)
    assert np.array_equal(zt.view(np.uint64), want_t.view(np.uint64))
