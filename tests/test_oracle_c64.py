"""CPU tests pinning the c64 oracle: the reference's golden vector (bit exact) and the
reference's own tolerance tests re-expressed with numpy's pocketfft as the independent FFT.

Mirrors src/unordered.rs:1071-1172 (test_fwd, test_fwd_monomial, test_roundtrip),
src/unordered.rs:1176-9396 (test_equivalency) and src/ordered.rs:389-467 (test_fft).
"""
import hashlib
import os

import numpy as np
import pytest

import oracle_lib as O

GOLD = os.path.join(os.path.dirname(__file__), "golden")


def rand_c(rng, *shape):
    return rng.random(shape) + 1j * rng.random(shape)


def test_golden_blobs_digests():
    a = open(os.path.join(GOLD, "unordered_n2048_dif4_b32_input.f64"), "rb").read()
    b = open(os.path.join(GOLD, "unordered_n2048_dif4_b32_target.f64"), "rb").read()
    assert hashlib.sha256(a).hexdigest() == "efb841a11a3d8325f6fac757487d37282c4ad37cb2801f48afd52a227304ebb1"
    assert hashlib.sha256(b).hexdigest() == "fd92a48d1f5d05edba9530a270931a0dcfed0d41a6b4cd8e77c2cc5d955e1cba"


def test_golden_input_generator_matches_blob():
    import importlib.util

    spec = importlib.util.spec_from_file_location("make_golden", os.path.join(GOLD, "make_golden.py"))
    mg = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mg)
    x = np.array(mg.std_rng_unit_f64(0, 4096))
    blob = np.fromfile(os.path.join(GOLD, "unordered_n2048_dif4_b32_input.f64"), dtype=np.float64)
    assert np.array_equal(x, blob)


def test_equivalency_bit_exact():
    """src/unordered.rs:1176-9396: assert_eq! on all 2048 entries."""
    x = np.fromfile(os.path.join(GOLD, "unordered_n2048_dif4_b32_input.f64"), dtype=np.complex128)
    t = np.fromfile(os.path.join(GOLD, "unordered_n2048_dif4_b32_target.f64"), dtype=np.complex128)
    y = O.UnorderedPlan(2048, O.DIF4, 32).fwd(x)
    assert np.array_equal(y.view(np.uint64), t.view(np.uint64))
    # the -O3 build used as the timed CPU baseline gives the same bits
    y2 = O.UnorderedPlan(2048, O.DIF4, 32, fast=True).fwd(x)
    assert np.array_equal(y2.view(np.uint64), t.view(np.uint64))


def test_sincospi64_exact_points():
    assert O.sincospi64(0.0) == (0.0, 1.0)
    assert O.sincospi64(0.5) == (1.0, 0.0)
    assert O.sincospi64(1.0)[1] == -1.0 and O.sincospi64(1.0)[0] == 0.0
    assert O.sincospi64(-0.5) == (-1.0, 0.0)
    s, c = O.sincospi64(0.25)
    assert abs(s - 2 ** -0.5) < 2e-16 and abs(c - 2 ** -0.5) < 2e-16
    import mpmath

    mpmath.mp.prec = 200
    rng = np.random.default_rng(1)
    for a in rng.uniform(-2, 2, 200):
        s, c = O.sincospi64(float(a))
        es, ec = mpmath.sinpi(mpmath.mpf(float(a))), mpmath.cospi(mpmath.mpf(float(a)))
        # <= 1 ulp (the reference's claim for its minimax polynomials)
        assert abs(s - es) <= np.spacing(abs(float(es))) and abs(c - ec) <= np.spacing(abs(float(ec)))


@pytest.mark.parametrize("algo", range(8))
def test_ordered_all_algos(algo):
    """src/ordered.rs:389-467: every algo, n = 2..1024, fwd vs independent FFT, inv round trip."""
    rng = np.random.default_rng(100 + algo)
    for k in range(1, 11):
        n = 1 << k
        x = rand_c(rng, n)
        plan = O.OrderedPlan(n, algo)
        y = plan.fwd(x)
        assert np.abs(y - np.fft.fft(x)).max() < 1e-12, (O.ALGO_NAMES[algo], n)
        z = plan.inv(y) / n
        assert np.abs(z - x).max() < 1e-14, (O.ALGO_NAMES[algo], n)
        # inverse is the unnormalised adjoint
        assert np.abs(plan.inv(x) - np.fft.ifft(x) * n).max() < 1e-12


def test_ordered_rejects_bad_sizes():
    with pytest.raises(ValueError):
        O.OrderedPlan(2048, O.DIF4)  # src/ordered.rs:244
    with pytest.raises(ValueError):
        O.OrderedPlan(48, O.DIF4)


def test_unordered_fwd_matches_fft_through_permutation():
    """src/unordered.rs:1073-1104"""
    rng = np.random.default_rng(2)
    for n in [128, 256, 512, 1024]:
        x = rand_c(rng, n)
        y = O.UnorderedPlan(n, O.DIF4, 32).fwd(x)
        pi = O.permutation(n, 32)
        assert np.abs(y[pi] - np.fft.fft(x)).max() < 1e-12


@pytest.mark.parametrize("algo", range(8))
def test_unordered_all_bases(algo):
    rng = np.random.default_rng(3 + algo)
    for k in range(5, 15):
        n = 1 << k
        for bk in range(5, min(k, 10) + 1):
            base_n = 1 << bk
            x = rand_c(rng, n)
            plan = O.UnorderedPlan(n, algo, base_n)
            y = plan.fwd(x)
            pi = O.permutation(n, base_n)
            ref = np.fft.fft(x)
            assert np.linalg.norm(y[pi] - ref) / np.linalg.norm(ref) < 1e-15 * k
            z = plan.inv(y) / n
            assert np.abs(z - x).max() < 1e-12


def test_unordered_small_n_equals_base():
    rng = np.random.default_rng(4)
    for n in [1, 2, 4, 8, 16]:
        x = rand_c(rng, n)
        y = O.UnorderedPlan(n, O.DIF2, n).fwd(x)
        assert np.abs(y - np.fft.fft(x)).max() < 1e-13


def test_unordered_rejects_bad_params():
    for args in [(2048, O.DIF4, 16), (2048, O.DIF4, 2048), (64, O.DIF4, 128), (100, O.DIF4, 32)]:
        with pytest.raises(ValueError):
            O.UnorderedPlan(*args)  # src/unordered.rs:660-669


def test_roundtrip():
    """src/unordered.rs:1141-1172"""
    rng = np.random.default_rng(5)
    for n in [32, 64, 256, 512, 1024]:
        x = rand_c(rng, n)
        plan = O.UnorderedPlan(n, O.DIF4, 32)
        assert np.abs(plan.inv(plan.fwd(x)) / n - x).max() < 1e-12


def test_fwd_monomial():
    """src/unordered.rs:1108-1137"""
    rng = np.random.default_rng(6)
    for n in [256, 512, 1024]:
        for base_n in [32, n, n // 2, n // 4, n // 8]:
            plan = O.UnorderedPlan(n, O.DIF4, base_n)
            for _ in range(10):
                d = int(rng.integers(0, n))
                z = np.zeros(n, np.complex128)
                z[d] = 1.0
                assert np.abs(plan.fwd_monomial(d) - plan.fwd(z)).max() < 1e-12


def test_permutation_closed_form():
    """pi(i) = bitrev_L(lo) * base_n + hi."""
    for n, base_n in [(2048, 32), (2048, 256), (4096, 1024), (65536, 512), (256, 256)]:
        L = (n // base_n).bit_length() - 1
        pi = O.permutation(n, base_n)
        i = np.arange(n)
        lo, hi = i & ((1 << L) - 1), i >> L
        rev = np.zeros_like(lo)
        for b in range(L):
            rev |= ((lo >> b) & 1) << (L - 1 - b)
        assert np.array_equal(pi, rev * base_n + hi)
        inv = np.array([O.lib().orc_bit_rev_twice_inv(n.bit_length() - 1, base_n.bit_length() - 1, int(p)) for p in range(n)])
        assert np.array_equal(inv[pi], i)


def test_batch_threads_same_bits():
    rng = np.random.default_rng(7)
    x = rand_c(rng, 64, 2048)
    plan = O.UnorderedPlan(2048, O.DIF16, 256)
    a = plan.fwd(x, threads=1)
    b = plan.fwd(x, threads=4)
    assert np.array_equal(a.view(np.uint64), b.view(np.uint64))
    c = O.UnorderedPlan(2048, O.DIF16, 256, fast=True).fwd(x, threads=3)
    assert np.array_equal(a.view(np.uint64), c.view(np.uint64))


@pytest.mark.parametrize("algo", range(8))
def test_avx2_build_same_bits_as_scalar(algo):
    """The timed CPU baseline (-O3 -march=x86-64-v3: two interleaved complex per AVX2 register, the
    reference's c64x2 form) must give the scalar oracle's bits for every algorithm and direction."""
    rng = np.random.default_rng(900 + algo)
    for n, base in [(32, 32), (64, 64), (256, 256), (512, 512), (1024, 32), (2048, 256), (4096, 1024), (8192, 512)]:
        x = rand_c(rng, 3, n)
        slow, fast = O.UnorderedPlan(n, algo, base), O.UnorderedPlan(n, algo, base, fast=True)
        y = slow.fwd(x)
        assert np.array_equal(y.view(np.uint64), fast.fwd(x).view(np.uint64)), (algo, n, base)
        assert np.array_equal(slow.inv(y).view(np.uint64), fast.inv(y).view(np.uint64)), (algo, n, base)
    for n in [2, 4, 8, 16, 128, 1024]:
        x = rand_c(rng, 2, n)
        assert np.array_equal(O.OrderedPlan(n, algo).fwd(x).view(np.uint64), O.OrderedPlan(n, algo, fast=True).fwd(x).view(np.uint64))
        assert np.array_equal(O.OrderedPlan(n, algo).inv(x).view(np.uint64), O.OrderedPlan(n, algo, fast=True).inv(x).view(np.uint64))


def test_pointwise_product_is_num_complex_arithmetic():
    """orc_c64_pointwise restates what a Rust caller's `a * b` / `acc + a * b` on Complex64 computes
    (num_complex: re = a.re*b.re - a.im*b.im, im = a.re*b.im + a.im*b.re, each operation rounded, no FMA):
    checked against the same expression in Python floats, element by element."""
    rng = np.random.default_rng(12)
    a, b, acc = rand_c(rng, 257) - 0.5, rand_c(rng, 257) * 3.0, rand_c(rng, 257)
    prod, fused = O.c64_pointwise(a, b), O.c64_pointwise(a, b, acc)
    for i in range(a.size):
        ar, ai, br, bi = float(a[i].real), float(a[i].imag), float(b[i].real), float(b[i].imag)
        re, im = ar * br - ai * bi, ar * bi + ai * br
        assert (prod[i].real, prod[i].imag) == (re, im)
        assert (fused[i].real, fused[i].imag) == (float(acc[i].real) + re, float(acc[i].imag) + im)


def test_oracle_external_product_composition_is_exact_negacyclic_sum():
    """The checker the GPU tests hold cfft_c64_fwd_mul_inv against -- oracle fwd per term, num_complex product and sum
    in term order, oracle inv -- computes sum_k a_k * b_k modulo X^N + 1: equal to the exact integer schoolbook
    result after rounding (what the crate exists for, README.md:10-17), and independent of the plan's base size."""
    npoly, n, k = 512, 256, 3
    rng = np.random.default_rng(5)
    a = rng.integers(-(1 << 16), 1 << 16, size=(k, npoly))
    b = rng.integers(-(1 << 10), 1 << 10, size=(k, npoly))
    want = np.zeros(npoly, dtype=object)
    for j in range(k):
        full = np.convolve(a[j].astype(object), b[j].astype(object))
        want += full[:npoly]
        want[: npoly - 1] -= full[npoly:]
    twist = np.exp(1j * np.pi * np.arange(n) / npoly)
    fold = lambda p: (p[..., :n] + 1j * p[..., n:]) * twist
    results = []
    for algo, base in [(O.DIF16, 256), (O.DIF4, 32), (O.DIT8, 64)]:
        plan = O.UnorderedPlan(n, algo, base)
        fa, fb = fold(a), fold(b)
        acc = O.c64_pointwise(plan.fwd(fa[0]), plan.fwd(fb[0]))
        for j in range(1, k):
            acc = O.c64_pointwise(plan.fwd(fa[j]), plan.fwd(fb[j]), acc)
        z = plan.inv(acc) / n * np.conj(twist)
        got = np.concatenate([z.real, z.imag])
        assert np.array_equal(np.rint(got).astype(np.int64), want.astype(np.int64)), (algo, base)
        results.append(got)
    assert np.abs(results[0] - results[1]).max() < 1e-3 and np.abs(results[0] - results[2]).max() < 1e-3
