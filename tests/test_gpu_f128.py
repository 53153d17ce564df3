"""GPU parity tests for fft128, through the C ABI.

Bar: bit-exact against the oracle's FMA variant (the arithmetic the reference's AVX2/AVX-512
paths execute on x86, src/fft128/f128_ops.rs:837-841).  Stated tolerance against the scalar
variant and in absolute terms: |gpu - scalar_ref| <= 32 * log2(n) * 2^-106 * rms|X|, and the
reference's own property bound 1e-30 * N for the negacyclic product (src/fft128/mod.rs:2062).
"""
from fractions import Fraction

import numpy as np
import pytest

import oracle_lib as O
from f128_util import dd_mul_pointwise, dd_to_fraction, negacyclic_schoolbook_exact

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def C():
    import concrete_fft_b200

    return concrete_fft_b200


@pytest.fixture(scope="module")
def torch():
    import torch

    return torch


def planes_random(rng, batch, n, full_width=True):
    re0, im0 = rng.random((batch, n)), rng.random((batch, n))
    if full_width:
        re1 = (rng.random((batch, n)) - 0.5) * np.spacing(re0)
        im1 = (rng.random((batch, n)) - 0.5) * np.spacing(im0)
    else:
        re1, im1 = np.zeros((batch, n)), np.zeros((batch, n))
    return [re0, re1, im0, im1]


def dev_run(torch, fn, planes):
    d = [torch.from_numpy(np.ascontiguousarray(p).copy()).cuda() for p in planes]
    fn(*d)
    torch.cuda.synchronize()
    return [t.cpu().numpy() for t in d]


def bits_equal(a, b):
    return all(np.array_equal(x.view(np.uint64), y.view(np.uint64)) for x, y in zip(a, b))


@pytest.mark.parametrize("logn", range(5, 15))
def test_fwd_inv_bit_exact(C, torch, logn):
    """n = 32 .. 16384 (the reference bench's range, benches/fft.rs:213)."""
    n = 1 << logn
    rng = np.random.default_rng(logn)
    batch = 3 if n <= 4096 else 2
    planes = planes_random(rng, batch, n)
    plan = C.fft128.Plan(n)
    assert plan.fft_size() == n
    ref = O.F128Plan(n)
    y = dev_run(torch, plan.fwd, planes)
    want = ref.fwd(*planes, variant=O.F128_FMA)
    assert bits_equal(y, want)
    z = dev_run(torch, plan.inv, y)
    assert bits_equal(z, ref.inv(*want, variant=O.F128_FMA))
    tw = plan.twiddles()
    assert bits_equal(tw, ref.twiddles())
    # stated ulp bound against the reference's scalar multiply variant
    sc = ref.fwd(*planes, variant=O.F128_SCALAR)
    rms = np.sqrt(np.mean(want[0] ** 2 + want[2] ** 2))
    bound = 32 * logn * 2.0 ** -106 * rms
    for hi, lo, shi, slo in [(y[0], y[1], sc[0], sc[1]), (y[2], y[3], sc[2], sc[3])]:
        diff = (hi - shi) + (lo - slo)
        assert np.abs(diff).max() <= bound


def test_many_small_transforms_per_tile(C, torch):
    rng = np.random.default_rng(20)
    for n, batch in [(32, 129), (64, 70), (512, 9), (2048, 5)]:
        planes = planes_random(rng, batch, n)
        plan = C.fft128.Plan(n)
        assert bits_equal(dev_run(torch, plan.fwd, planes), O.F128Plan(n).fwd(*planes, variant=O.F128_FMA))


@pytest.mark.parametrize("npoly", [64, 512, 4096])
def test_negacyclic_product_property(C, torch, npoly):
    """src/fft128/mod.rs:1972-2065 through the CUDA path (host-memory entry points)."""
    rng = np.random.default_rng(npoly)
    n = npoly // 2
    lhs, rhs = rng.random(npoly), rng.random(npoly)
    exact = negacyclic_schoolbook_exact(lhs, rhs)
    plan = C.fft128.Plan(n)
    L = [lhs[:n].copy(), np.zeros(n), lhs[n:].copy(), np.zeros(n)]
    R = [rhs[:n].copy(), np.zeros(n), rhs[n:].copy(), np.zeros(n)]
    plan.fwd(*L)
    plan.fwd(*R)
    P = [np.ascontiguousarray(p) for p in dd_mul_pointwise(L, R, 2.0 / npoly)]
    plan.inv(*P)
    got_hi = np.concatenate([P[0], P[2]])
    got_lo = np.concatenate([P[1], P[3]])
    err = max(abs(dd_to_fraction(h, l) - e) for h, l, e in zip(got_hi, got_lo, exact))
    assert float(err) < 1e-30 * npoly


def test_length_mismatch_panics(C, torch):
    plan = C.fft128.Plan(64)
    a = np.zeros(64)
    with pytest.raises(C.PanicError):
        plan.fwd(a, a.copy(), a.copy(), np.zeros(63))  # src/fft128/mod.rs:1912-1915
    with pytest.raises(C.PanicError):
        plan.fwd(np.zeros(0), np.zeros(0), np.zeros(0), np.zeros(0))


def test_full_size_n2048_batch16384_roundtrip(C, torch):
    """BASELINE config 4 at full size (1 GiB): round trip to double-double accuracy and sampled
    rows bit-exact against the oracle."""
    n, batch = 2048, 16384
    plan = C.fft128.Plan(n)
    g = torch.Generator(device="cuda").manual_seed(1234)
    re0 = torch.rand(batch, n, dtype=torch.float64, device="cuda", generator=g)
    im0 = torch.rand(batch, n, dtype=torch.float64, device="cuda", generator=g)
    re1, im1 = torch.zeros_like(re0), torch.zeros_like(im0)
    w = [re0.clone(), re1.clone(), im0.clone(), im1.clone()]
    plan.fwd(*w)
    rows = [0, 5000, 16383]
    src = [t[rows].cpu().numpy() for t in (re0, re1, im0, im1)]
    want = O.F128Plan(n).fwd(*src, variant=O.F128_FMA)
    assert bits_equal([t[rows].cpu().numpy() for t in w], want)
    plan.inv(*w)
    torch.cuda.synchronize()
    # (hi + lo) / n - x, evaluated in double-double: hi/n is exact (n is a power of two)
    err_re = ((w[0] / n - re0) + w[1] / n).abs().max()
    err_im = ((w[2] / n - im0) + w[3] / n).abs().max()
    assert float(err_re) < 1e-28 and float(err_im) < 1e-28


def test_unaligned_planes_take_the_scalar_path(C, torch):
    """Planes that are only 8-byte aligned must not use the 128-bit HBM accesses."""
    rng = np.random.default_rng(30)
    n, batch = 64, 5
    planes = planes_random(rng, batch, n)
    plan = C.fft128.Plan(n)
    big = [torch.zeros(batch * n + 1, dtype=torch.float64, device="cuda") for _ in range(4)]
    views = [b[1:] for b in big]  # data_ptr is 8 mod 16
    for v, p in zip(views, planes):
        v.copy_(torch.from_numpy(p.reshape(-1)))
    assert views[0].data_ptr() % 16 == 8
    plan.fwd(*views)
    torch.cuda.synchronize()
    want = O.F128Plan(n).fwd(*planes, variant=O.F128_FMA)
    assert bits_equal([v.cpu().numpy().reshape(batch, n) for v in views], want)


def test_f128_operator_set_bit_exact(C, torch):
    """the scalar f128 operators on device arrays, bit-exact against the oracle's
    restatement of src/fft128/f128_ops.rs (which meets the reference's 2^-104 / 2^-101 bounds)."""
    rng = np.random.default_rng(40)
    n = 100003
    a_hi = rng.uniform(-4, 4, n)
    b_hi = rng.uniform(0.25, 4, n) * rng.choice([-1.0, 1.0], n)
    a_lo = (rng.random(n) - 0.5) * np.spacing(a_hi)
    b_lo = (rng.random(n) - 0.5) * np.spacing(b_hi)
    dev = [torch.from_numpy(x).cuda() for x in (a_hi, a_lo, b_hi, b_lo)]
    for op in ["add", "sub", "mul", "div", "add_estimate", "sub_estimate", "div_estimate"]:
        hi, lo = C.fft128.f128_op(op, *dev)
        torch.cuda.synchronize()
        want_hi, want_lo = O.f128_binary_op(op, a_hi, a_lo, b_hi, b_lo)
        assert np.array_equal(hi.cpu().numpy().view(np.uint64), want_hi.view(np.uint64)), op
        assert np.array_equal(lo.cpu().numpy().view(np.uint64), want_lo.view(np.uint64)), op


def test_f128_mixed_unary_and_compare_operators_bit_exact(C, torch):
    """the rest of the `f128` operator surface (f128_ops.rs:48-274 operator impls with f64 operands, :279-455 mixed forms,
    :404 sqr, :494-511 to_f64 / is_nan / abs, :514-618 sincospi) on device arrays: bit-exact against the oracle"""
    rng = np.random.default_rng(41)
    n = 50021
    a_hi = rng.uniform(-4, 4, n)
    b_hi = rng.uniform(0.25, 4, n) * rng.choice([-1.0, 1.0], n)
    a_lo = (rng.random(n) - 0.5) * np.spacing(a_hi)
    b_lo = (rng.random(n) - 0.5) * np.spacing(b_hi)
    d = {k: torch.from_numpy(v).cuda() for k, v in dict(a_hi=a_hi, a_lo=a_lo, b_hi=b_hi, b_lo=b_lo).items()}
    eq = lambda t, w: np.array_equal(t.cpu().numpy().view(np.uint64), w.view(np.uint64))
    for op in ["add_f128_f64", "sub_f128_f64", "mul_f128_f64", "div_f128_f64"]:
        hi, lo = C.fft128.f128_op(op, d["a_hi"], d["a_lo"], d["b_hi"], None)
        wh, wl = O.f128_binary_op(op, a_hi, a_lo, b_hi, None)
        assert eq(hi, wh) and eq(lo, wl), op
    for op in ["sub_f64_f128", "div_f64_f128"]:
        hi, lo = C.fft128.f128_op(op, d["a_hi"], None, d["b_hi"], d["b_lo"])
        wh, wl = O.f128_binary_op(op, a_hi, None, b_hi, b_lo)
        assert eq(hi, wh) and eq(lo, wl), op
    for op in ["add_f64_f64", "sub_f64_f64", "mul_f64_f64", "div_f64_f64"]:
        hi, lo = C.fft128.f128_op(op, d["a_hi"], None, d["b_hi"], None)
        wh, wl = O.f128_binary_op(op, a_hi, None, b_hi, None)
        assert eq(hi, wh) and eq(lo, wl), op
    for op in ["sqr", "abs", "neg", "is_nan"]:
        hi, lo = C.fft128.f128_unary(op, d["a_hi"], d["a_lo"])
        wh, wl = O.f128_unary_op(op, a_hi, a_lo)
        assert eq(hi, wh) and eq(lo, wl), op
    # sincospi on [-1, 1], including the quadrant / sixteenth boundaries and the end points
    x_hi = np.concatenate([rng.uniform(-1, 1, 20000), np.arange(-32, 33) / 32.0, [1.0, -1.0, 0.0, 0.5, -0.5, 0.03125]])
    x_lo = (rng.random(x_hi.size) - 0.5) * np.spacing(x_hi) * (np.abs(x_hi) < 1)
    (sh, sl), (ch, cl) = C.fft128.f128_unary("sincospi", torch.from_numpy(x_hi).cuda(), torch.from_numpy(x_lo).cuda())
    (wsh, wsl), (wch, wcl) = O.f128_unary_op("sincospi", x_hi, x_lo)
    assert eq(sh, wsh) and eq(sl, wsl) and eq(ch, wch) and eq(cl, wcl)
    with pytest.raises(C.PanicError):  # the reference panics outside [-1, 1]
        C.fft128.f128_unary("sincospi", torch.tensor([0.5, 1.5], dtype=torch.float64, device="cuda"), torch.zeros(2, dtype=torch.float64, device="cuda"))
    # comparisons, NaNs and equal-hi cases included
    c_hi = a_hi.copy()
    c_lo = a_lo.copy()
    c_hi[::7] = b_hi[::7]
    c_lo[::14] = b_lo[::14]
    c_hi[5::1001] = np.nan
    c_lo[9::1003] = np.nan
    dc = [torch.from_numpy(v).cuda() for v in (c_hi, c_lo)]
    got = C.fft128.f128_compare(dc[0], dc[1], d["b_hi"], d["b_lo"]).cpu().numpy()
    assert np.array_equal(got, O.f128_compare(c_hi, c_lo, b_hi, b_lo)) and set(np.unique(got)) == {-1, 0, 1, 2}
    got = C.fft128.f128_compare(dc[0], dc[1], d["b_hi"]).cpu().numpy()  # against f64 values
    assert np.array_equal(got, O.f128_compare(c_hi, c_lo, b_hi))
    hi, _ = C.fft128.f128_unary("is_nan", dc[0], dc[1])
    assert np.array_equal(hi.cpu().numpy() != 0, np.isnan(c_hi) | np.isnan(c_lo))


def test_pointwise_product_bit_exact_and_full_size_convolution(C, torch):
    """fwd -> point-wise product -> inv entirely on the device at BASELINE configs[3] size
    (n = 2048, batch 16384 = 1 GiB per operand): the product kernel is bit-exact against the oracle,
    and sampled rows satisfy the reference's negacyclic-convolution bound 1e-30 * N."""
    n, batch = 2048, 16384
    npoly = 2 * n
    g = torch.Generator(device="cuda").manual_seed(99)
    L = [torch.rand(batch, n, dtype=torch.float64, device="cuda", generator=g), torch.zeros(batch, n, dtype=torch.float64, device="cuda"),
         torch.rand(batch, n, dtype=torch.float64, device="cuda", generator=g), torch.zeros(batch, n, dtype=torch.float64, device="cuda")]
    R = [torch.rand(batch, n, dtype=torch.float64, device="cuda", generator=g), torch.zeros(batch, n, dtype=torch.float64, device="cuda"),
         torch.rand(batch, n, dtype=torch.float64, device="cuda", generator=g), torch.zeros(batch, n, dtype=torch.float64, device="cuda")]
    # torch's CUDA generator yields odd multiples of 2^-54; the exact schoolbook checker wants multiples
    # of 2^-53 (what rand::random produces in the reference test), so quantise to 2^-40
    for t in (L[0], L[2], R[0], R[2]):
        t.mul_(2.0 ** 40).floor_().mul_(2.0 ** -40)
    rows = [0, 4321, 16383]
    lhs = [np.concatenate([L[0][r].cpu().numpy(), L[2][r].cpu().numpy()]) for r in rows]
    rhs = [np.concatenate([R[0][r].cpu().numpy(), R[2][r].cpu().numpy()]) for r in rows]
    plan = C.fft128.Plan(n)
    plan.fwd(*L)
    plan.fwd(*R)
    Ls = [[t[r].cpu().numpy() for t in L] for r in rows]
    Rs = [[t[r].cpu().numpy() for t in R] for r in rows]
    C.fft128.cplx_mul_scale(L, R, 2.0 / npoly)
    torch.cuda.synchronize()
    for k, r in enumerate(rows):
        want = O.f128_cplx_mul_scale(Ls[k], Rs[k], 2.0 / npoly)
        assert bits_equal([t[r].cpu().numpy() for t in L], want)
    plan.inv(*L)
    torch.cuda.synchronize()
    for k, r in enumerate(rows):
        exact = negacyclic_schoolbook_exact(lhs[k], rhs[k])
        hi = np.concatenate([L[0][r].cpu().numpy(), L[2][r].cpu().numpy()])
        lo = np.concatenate([L[1][r].cpu().numpy(), L[3][r].cpu().numpy()])
        err = max(abs(dd_to_fraction(h, l) - e) for h, l, e in zip(hi, lo, exact))
        assert float(err) < 1e-30 * npoly  # src/fft128/mod.rs:2062


@pytest.mark.parametrize("n,batch", [(32, 1), (32, 70), (64, 33), (256, 3), (256, 9), (1024, 3), (2048, 1), (2048, 3), (4096, 2), (8192, 2)])
def test_fwd_mul_inv_bit_exact(C, torch, n, batch):
    """cfft_f128_fwd_mul_inv (one kernel for n <= 4096, the three launches above): the bits of the oracle's fwd (FMA
    butterflies) -> scalar cplx_mul * factor (src/fft128/mod.rs:2033-2047) -> inv, of the library's three separate
    calls and of the composed path; rhs shared by the batch or one row per transform; ragged last tiles."""
    import os

    rng = np.random.default_rng(7 * n + batch)
    plan = C.fft128.Plan(n)
    assert plan.has_fused_mul_kernel() == (n <= 4096)
    ref = O.F128Plan(n)
    factor = 2.0 / (2 * n)
    for shared in (True, False):
        lhs = planes_random(rng, batch, n)
        rhs = [p - 0.5 * (i % 2 == 0) for i, p in enumerate(planes_random(rng, 1 if shared else batch, n))]  # Fourier-domain operand
        F = ref.fwd(*lhs, variant=O.F128_FMA)
        R = [np.broadcast_to(p, (batch, n)).copy() for p in rhs]
        want = ref.inv(*O.f128_cplx_mul_scale(F, R, factor), variant=O.F128_FMA)
        dl = [torch.from_numpy(p.copy()).cuda() for p in lhs]
        dr = [torch.from_numpy(np.ascontiguousarray(p if not shared else p[0]).copy()).cuda() for p in rhs]
        launches = C.launch_count()
        plan.fwd_mul_inv(dl, dr, factor)
        torch.cuda.synchronize()
        if n <= 4096:
            assert C.launch_count() - launches == 1
        assert bits_equal([t.cpu().numpy() for t in dl], want), (n, batch, shared)
        # the library's three calls
        d2 = [torch.from_numpy(p.copy()).cuda() for p in lhs]
        plan.fwd(*d2)
        C.fft128.cplx_mul_scale(d2, [torch.from_numpy(p).cuda() for p in R], factor)
        plan.inv(*d2)
        torch.cuda.synchronize()
        assert bits_equal([t.cpu().numpy() for t in d2], want), (n, batch, shared)
        os.environ["CFFT_B200_FUSED_MUL_COMPOSED"] = "1"
        try:
            d3 = [torch.from_numpy(p.copy()).cuda() for p in lhs]
            plan.fwd_mul_inv(d3, dr, factor)
        finally:
            del os.environ["CFFT_B200_FUSED_MUL_COMPOSED"]
        torch.cuda.synchronize()
        assert bits_equal([t.cpu().numpy() for t in d3], want), (n, batch, shared)
    with pytest.raises(C.PanicError):
        plan.fwd_mul_inv(dl, [t[..., : n // 4].contiguous() for t in dr], factor)


def test_fwd_mul_inv_negacyclic_product_full_size(C, torch):
    """BASELINE configs[3] size (n = 2048, batch 16384) through the one-kernel product: equals the three separate
    launches bit for bit on every row, and sampled rows meet the reference's bound 1e-30 * N against the exact
    schoolbook negacyclic convolution (src/fft128/mod.rs:2062)."""
    n, batch = 2048, 16384
    npoly = 2 * n
    g = torch.Generator(device="cuda").manual_seed(123)
    def operand():
        t = [torch.rand(batch, n, dtype=torch.float64, device="cuda", generator=g), torch.zeros(batch, n, dtype=torch.float64, device="cuda"),
             torch.rand(batch, n, dtype=torch.float64, device="cuda", generator=g), torch.zeros(batch, n, dtype=torch.float64, device="cuda")]
        for x in (t[0], t[2]):
            x.mul_(2.0 ** 40).floor_().mul_(2.0 ** -40)
        return t
    L, R = operand(), operand()
    rows = [0, 777, 16383]
    lhs = [np.concatenate([L[0][r].cpu().numpy(), L[2][r].cpu().numpy()]) for r in rows]
    rhs = [np.concatenate([R[0][r].cpu().numpy(), R[2][r].cpu().numpy()]) for r in rows]
    plan = C.fft128.Plan(n)
    plan.fwd(*R)
    L2 = [t.clone() for t in L]
    plan.fwd_mul_inv(L, R, 2.0 / npoly)
    plan.fwd(*L2)
    C.fft128.cplx_mul_scale(L2, R, 2.0 / npoly)
    plan.inv(*L2)
    torch.cuda.synchronize()
    for a, b in zip(L, L2):
        assert torch.equal(a.view(torch.int64), b.view(torch.int64))
    for k, r in enumerate(rows):
        exact = negacyclic_schoolbook_exact(lhs[k], rhs[k])
        hi = np.concatenate([L[0][r].cpu().numpy(), L[2][r].cpu().numpy()])
        lo = np.concatenate([L[1][r].cpu().numpy(), L[3][r].cpu().numpy()])
        err = max(abs(dd_to_fraction(h, l) - e) for h, l, e in zip(hi, lo, exact))
        assert float(err) < 1e-30 * npoly


def test_random_sizes_fuzz(C, torch):
    rng = np.random.default_rng(424242)
    for trial in range(16):
        logn = int(rng.integers(5, 16))
        n = 1 << logn
        batch = int(rng.integers(1, max(2, min(50, (1 << 17) // n))))
        planes = planes_random(rng, batch, n, full_width=bool(rng.integers(0, 2)))
        plan = C.fft128.Plan(n)
        ref = O.F128Plan(n)
        if rng.random() < 0.5:
            y = dev_run(torch, plan.fwd, planes)
        else:
            y = [p.copy() for p in planes]
            plan.fwd(*y)
        want = ref.fwd(*planes, variant=O.F128_FMA)
        assert bits_equal(y, want), (trial, n, batch)
        assert bits_equal(dev_run(torch, plan.inv, y), ref.inv(*want, variant=O.F128_FMA)), (trial, n, batch)


def test_host_entry_points(C, torch):
    rng = np.random.default_rng(50)
    n, batch = 1024, 300  # > one zero-copy call, exercises the chunked pipeline too
    planes = planes_random(rng, batch, n)
    plan = C.fft128.Plan(n)
    ref = O.F128Plan(n)
    want = ref.fwd(*planes, variant=O.F128_FMA, threads=8)
    h = [p.copy() for p in planes]
    plan.fwd(*h)
    assert bits_equal(h, want)
    back = ref.inv(*want, variant=O.F128_FMA, threads=8)
    plan.inv(*h)
    assert bits_equal(h, back)
    both = [p.copy() for p in planes]
    plan.fwd_inv_host(*both)
    assert bits_equal(both, back)
    one = [p[:1].copy() for p in planes]  # single transform: zero-copy path
    plan.fwd(*one)
    assert bits_equal(one, [w[:1] for w in want])


def test_strided_planes_bit_exact(C, torch):
    """cfft_f128_fwd_strided / _inv_strided: rows row_stride >= n doubles apart in each plane (x[:, j] of [batch, k, n]),
    bit-identical to the packed call, neighbouring rows untouched."""
    rng = np.random.default_rng(99)
    for n, batch, k, j in [(256, 5, 3, 2), (2048, 3, 2, 0), (8192, 2, 2, 1)]:
        plan = C.fft128.Plan(n)
        ref = O.F128Plan(n)
        host = [p.reshape(batch, k, n) for p in planes_random(rng, batch * k, n)]
        dev = [torch.from_numpy(p.copy()).cuda() for p in host]
        plan.fwd_strided(*[d[:, j] for d in dev])
        torch.cuda.synchronize()
        want = [p.copy() for p in host]
        out = ref.fwd(*[np.ascontiguousarray(p[:, j]) for p in host], variant=O.F128_FMA)
        for w, o in zip(want, out):
            w[:, j] = o
        assert bits_equal([d.cpu().numpy() for d in dev], want), (n, batch, k, j)
        plan.inv_strided(*[d[:, j] for d in dev])
        torch.cuda.synchronize()
        back = ref.inv(*[np.ascontiguousarray(w[:, j]) for w in want], variant=O.F128_FMA)
        for w, o in zip(want, back):
            w[:, j] = o
        assert bits_equal([d.cpu().numpy() for d in dev], want), (n, batch, k, j)


def test_host_call_sharded_over_replicas(C, torch):
    """cfft_f128_host_multi: one host call, the batch cut over replicas of the plan; bit-identical to the oracle."""
    from concrete_fft_b200.sharding import MultiGpu

    rng = np.random.default_rng(909)
    n = 1024
    plan = C.fft128.Plan(n)
    ref = O.F128Plan(n)
    for devices in [[0, 0]] + ([[0, 1]] if torch.cuda.device_count() >= 2 else []):
        mg = MultiGpu(plan, devices)
        for batch in (1, 7):
            planes = planes_random(rng, batch, n)
            h = [p.copy() for p in planes]
            mg.fwd(*h)
            want = ref.fwd(*planes, variant=O.F128_FMA)
            assert bits_equal(h, want), (devices, batch)
            mg.inv(*h)
            assert bits_equal(h, ref.inv(*want, variant=O.F128_FMA)), (devices, batch)
