"""GPU parity tests for the c64 path, through the C ABI (via the Python mirror of the Rust API).

Bar: BIT-EXACT against the oracle (= the reference's result for the same plan), which is far
inside the stated tolerance (relative L2 <= 1e-13 * log2 N); the tolerance is asserted as well
against an independent FFT (numpy / pocketfft) so both statements are on record.
"""
import os

import numpy as np
import pytest

import oracle_lib as O

pytestmark = pytest.mark.gpu

GOLD = os.path.join(os.path.dirname(__file__), "golden")


@pytest.fixture(scope="module")
def C():
    import concrete_fft_b200

    return concrete_fft_b200


@pytest.fixture(scope="module")
def torch():
    import torch

    return torch


def rand_c(rng, *shape):
    return rng.random(shape) + 1j * rng.random(shape)


def bits_equal(a, b):
    return np.array_equal(np.ascontiguousarray(a).view(np.uint64), np.ascontiguousarray(b).view(np.uint64))


def dev_run(torch, fn, x):
    d = torch.from_numpy(np.ascontiguousarray(x).copy()).cuda()
    fn(d)
    torch.cuda.synchronize()
    return d.cpu().numpy()


def test_golden_vector_bit_exact_on_gpu(C, torch):
    """src/unordered.rs:1176-9396 (test_equivalency) executed by the CUDA path."""
    x = np.fromfile(os.path.join(GOLD, "unordered_n2048_dif4_b32_input.f64"), dtype=np.complex128)
    t = np.fromfile(os.path.join(GOLD, "unordered_n2048_dif4_b32_target.f64"), dtype=np.complex128)
    plan = C.unordered.Plan(2048, C.unordered.Method.UserProvided(C.ordered.FftAlgo.Dif4, 32))
    assert plan.algo() == (C.ordered.FftAlgo.Dif4, 32)
    assert bits_equal(dev_run(torch, plan.fwd, x), t)
    h = x.copy()
    plan.fwd(h)  # host-memory entry point (the literal Plan::fwd(&mut [c64]) drop-in)
    assert bits_equal(h, t)
    plan.inv(h)
    assert np.abs(h / 2048 - x).max() < 1e-12


@pytest.mark.parametrize("algo", range(8))
def test_ordered_all_algos_bit_exact(C, torch, algo):
    """src/ordered.rs:389-467: all 8 algorithms, n = 2 .. 1024, fwd and inv."""
    rng = np.random.default_rng(200 + algo)
    for k in range(0, 11):
        n = 1 << k
        x = rand_c(rng, 3, n)
        plan = C.ordered.Plan(n, C.ordered.Method.UserProvided(C.ordered.FftAlgo(algo)))
        assert plan.fft_size() == n and plan.algo() == C.ordered.FftAlgo(algo)
        assert plan.fft_scratch().size_bytes == 16 * n
        if n == 1:
            assert bits_equal(dev_run(torch, plan.fwd, x), x)
            continue
        ref = O.OrderedPlan(n, algo)
        y = dev_run(torch, plan.fwd, x)
        assert bits_equal(y, ref.fwd(x)), (algo, n)
        assert np.abs(y - np.fft.fft(x, axis=1)).max() < 1e-12
        assert bits_equal(dev_run(torch, plan.inv, y), ref.inv(y)), (algo, n)
        # twiddle tables are the reference's init_wt tables, bit for bit (NaN slots included)
        if n == 256 and algo == O.DIF16:
            assert plan.kernel_name() == "fast-b256-regs"  # ordered Dif16/256 == the register kernel's base FFT
        tw = np.zeros((2, 2 * n), np.complex128)
        if n >= (2 << (algo >> 1)):
            O.lib().orc_init_wt(2 << (algo >> 1), n, tw[0].ctypes.data, tw[1].ctypes.data)
        assert bits_equal(plan.twiddles(False), tw[0]) and bits_equal(plan.twiddles(True), tw[1])


@pytest.mark.parametrize("algo", range(8))
def test_unordered_plans_bit_exact(C, torch, algo):
    """Every (base_algo, base_n) the reference accepts, n = 32 .. 2^14: same bits, same order."""
    rng = np.random.default_rng(300 + algo)
    A = C.ordered.FftAlgo
    for k in range(5, 15):
        n = 1 << k
        for bk in sorted({5, 6, 8, 9, 10, k} & set(range(5, min(k, 10) + 1))):
            base_n = 1 << bk
            x = rand_c(rng, 2, n)
            plan = C.unordered.Plan(n, C.unordered.Method.UserProvided(A(algo), base_n))
            ref = O.UnorderedPlan(n, algo, base_n)
            y = dev_run(torch, plan.fwd, x)
            want = ref.fwd(x)
            assert bits_equal(y, want), (algo, n, base_n)
            z = dev_run(torch, plan.inv, y)
            assert bits_equal(z, ref.inv(want)), (algo, n, base_n)
            # stated tolerance vs an independent FFT through the permutation
            pi = plan.permutation().astype(np.int64)
            assert np.array_equal(pi, O.permutation(n, base_n))
            f = np.fft.fft(x, axis=1)
            rel = np.linalg.norm(y[:, pi] - f, axis=1) / np.linalg.norm(f, axis=1)
            assert rel.max() <= 1e-13 * k
            assert plan.fft_scratch().size_bytes == 16 * base_n
            if k in (5, 11) and bk == 5:
                assert bits_equal(plan.twiddles(False), ref.twiddles(False))
                assert bits_equal(plan.twiddles(True), ref.twiddles(True))


def test_unordered_small_sizes_base_equals_n(C, torch):
    rng = np.random.default_rng(4)
    for n in [1, 2, 4, 8, 16]:
        x = rand_c(rng, 5, n)
        plan = C.unordered.Plan(n, C.unordered.Method.UserProvided(C.ordered.FftAlgo.Dif4, n))
        y = dev_run(torch, plan.fwd, x)
        if n > 1:
            assert bits_equal(y, O.UnorderedPlan(n, O.DIF4, n).fwd(x))
        assert np.abs(y - np.fft.fft(x, axis=1)).max() < 1e-13


def test_large_n_multi_pass(C, torch):
    """N = 2^16 and 2^17 (levels above one tile run as HBM passes): bit-exact, correct order."""
    rng = np.random.default_rng(5)
    for n, base_n, algo in [(1 << 16, 1024, O.DIF16), (1 << 16, 512, O.DIT8), (1 << 17, 32, O.DIF4)]:
        x = rand_c(rng, 2, n)
        plan = C.unordered.Plan(n, C.unordered.Method.UserProvided(C.ordered.FftAlgo(algo), base_n))
        ref = O.UnorderedPlan(n, algo, base_n)
        y = dev_run(torch, plan.fwd, x)
        want = ref.fwd(x)
        assert bits_equal(y, want)
        assert bits_equal(dev_run(torch, plan.inv, y), ref.inv(want))


def test_measure_method_is_deterministic_and_valid(C, torch):
    rng = np.random.default_rng(6)
    for n in [64, 256, 512, 2048, 8192, 32768]:
        p1 = C.unordered.Plan(n, C.unordered.Method.Measure())  # plan fixed by rule, kernel variant autotuned
        p2 = C.unordered.Plan(n, C.unordered.Method.Measure())
        assert p1.algo() == p2.algo()
        algo, base_n = p1.algo()
        if n <= 256:
            assert base_n == n  # src/unordered.rs:561-564
        else:
            assert base_n == 256
            assert p1.kernel_name().startswith("fast-b256-")
        x = rand_c(rng, 2, n)
        assert bits_equal(dev_run(torch, p1.fwd, x), O.UnorderedPlan(n, int(algo), base_n).fwd(x))


def test_fwd_monomial(C, torch):
    """src/unordered.rs:1108-1137, plus bit-exactness against the oracle's table lookup."""
    rng = np.random.default_rng(7)
    for n in [256, 512, 1024]:
        for base_n in [32, n, n // 2, n // 4, n // 8]:
            plan = C.unordered.Plan(n, C.unordered.Method.UserProvided(C.ordered.FftAlgo.Dif4, base_n))
            ref = O.UnorderedPlan(n, O.DIF4, base_n)
            for _ in range(3):
                d = int(rng.integers(0, n))
                buf = torch.zeros(n, dtype=torch.complex128, device="cuda")
                plan.fwd_monomial(d, buf)
                torch.cuda.synchronize()
                got = buf.cpu().numpy()
                assert bits_equal(got, ref.fwd_monomial(d))
                z = np.zeros(n, np.complex128)
                z[d] = 1.0
                plan.fwd(z)
                assert np.abs(got - z).max() < 1e-12
                h = np.zeros(n, np.complex128)
                plan.fwd_monomial(d, h)
                assert bits_equal(h, got)
    with pytest.raises(C.PanicError):
        plan.fwd_monomial(n, np.zeros(n, np.complex128))  # degree < n, src/unordered.rs:859


def test_serde_standard_order_mapping(C, torch):
    """src/unordered.rs:9399-9467: plan1 (base 32) -> standard order -> plan2 (base 64) -> inv."""
    rng = np.random.default_rng(8)
    A = C.ordered.FftAlgo
    for n in [64, 128, 256, 512, 1024]:
        x = rand_c(rng, n)
        p1 = C.unordered.Plan(n, C.unordered.Method.UserProvided(A.Dif4, 32))
        p2 = C.unordered.Plan(n, C.unordered.Method.UserProvided(A.Dif4, 64))
        f = x.copy()
        p1.fwd(f)
        blob = p1.serialize_bincode(f)
        assert len(blob) == 8 + 16 * n
        g = np.zeros(n, np.complex128)
        p2.deserialize_bincode(blob, g)
        p2.inv(g)
        assert np.abs(g / n - x).max() < 1e-12
        # device gather / scatter agree with the host loops
        fd = torch.from_numpy(f).cuda()
        std = p1.serialize_fourier_buffer(fd)
        assert bits_equal(std.cpu().numpy(), p1.serialize_fourier_buffer(f))
        back = torch.zeros_like(fd)
        p1.deserialize_fourier_buffer(std, back)
        torch.cuda.synchronize()
        assert bits_equal(back.cpu().numpy(), f)
        with pytest.raises(C.InvalidLength):
            p2.deserialize_fourier_buffer(np.zeros(n - 1, np.complex128), g)
        with pytest.raises(C.InvalidLength):
            p2.deserialize_fourier_buffer(np.zeros(n + 1, np.complex128), g)


def test_length_mismatch_panics(C, torch):
    plan = C.unordered.Plan(256, C.unordered.Method.UserProvided(C.ordered.FftAlgo.Dif4, 32))
    with pytest.raises(C.PanicError):
        plan.fwd(np.zeros(255, np.complex128))  # assert_eq!, src/unordered.rs:827
    with pytest.raises(C.PanicError):
        plan.inv(torch.zeros(300, dtype=torch.complex128, device="cuda"))
    with pytest.raises(C.PanicError):
        plan.fwd(np.zeros(0, np.complex128))


def test_host_pipeline_pinned_and_pageable_ragged_batches(C, torch):
    """The host entry chunks the batch through three slots; sizes that do not divide a chunk,
    pageable and pinned memory must all give the oracle's bits."""
    rng = np.random.default_rng(9)
    n = 2048
    plan = C.unordered.Plan(n, C.unordered.Method.UserProvided(C.ordered.FftAlgo.Dif16, 256))
    ref = O.UnorderedPlan(n, O.DIF16, 256)
    for batch in [1, 3, 1024 + 7, 3 * 1024 + 1]:
        x = rand_c(rng, batch, n)
        want = ref.fwd(x, threads=8)
        pageable = x.copy()
        plan.fwd(pageable)
        assert bits_equal(pageable, want)
        pinned = torch.from_numpy(x.copy()).pin_memory()
        plan.fwd(pinned.numpy())
        assert bits_equal(pinned.numpy(), want)
        plan.inv(pinned.numpy())
        assert bits_equal(pinned.numpy(), ref.inv(want, threads=8))
        both = x.copy()
        plan.fwd_inv_host(both)
        assert bits_equal(both, pinned.numpy())


def test_clone_and_concurrent_streams(C, torch):
    rng = np.random.default_rng(10)
    n = 1024
    plan = C.unordered.Plan(n, C.unordered.Method.UserProvided(C.ordered.FftAlgo.Dit4, 64))
    twin = plan.clone()
    assert twin.algo() == plan.algo() and twin.fft_size() == n
    x = rand_c(rng, 64, n)
    want = O.UnorderedPlan(n, O.DIT4, 64).fwd(x)
    streams = [torch.cuda.Stream() for _ in range(4)]
    bufs = [torch.from_numpy(x.copy()).cuda() for _ in streams]
    torch.cuda.synchronize()
    for s, b in zip(streams, bufs):
        with torch.cuda.stream(s):
            (plan if s is streams[0] else twin).fwd(b)
    torch.cuda.synchronize()
    for b in bufs:
        assert bits_equal(b.cpu().numpy(), want)


def test_full_size_properties_n2048_batch65536(C, torch):
    """BASELINE config 2 at full size (2 GiB on the device): size-independent properties --
    round trip, linearity, Parseval -- plus bit-exactness of sampled rows against the oracle."""
    n, batch = 2048, 65536
    plan = C.unordered.Plan(n, C.unordered.Method.UserProvided(C.ordered.FftAlgo.Dif16, 256))
    g = torch.Generator(device="cuda").manual_seed(0x5EED0000)
    x = torch.rand(batch, n, 2, dtype=torch.float64, device="cuda", generator=g)
    x = torch.view_as_complex(x).contiguous()
    # row 0 = the reference's golden-vector input
    gold = np.fromfile(os.path.join(GOLD, "unordered_n2048_dif4_b32_input.f64"), dtype=np.complex128)
    x[0] = torch.from_numpy(gold).cuda()
    y = x.clone()
    plan.fwd(y)
    rows = [0, 1, 777, 32768, 65535]
    ref = O.UnorderedPlan(n, O.DIF16, 256)
    xs = x[rows].cpu().numpy()
    assert bits_equal(y[rows].cpu().numpy(), ref.fwd(xs))
    # Parseval: sum |X|^2 = n sum |x|^2, per row
    ex = (x.real ** 2 + x.imag ** 2).sum(1)
    ey = (y.real ** 2 + y.imag ** 2).sum(1)
    assert float(((ey - n * ex).abs() / (n * ex)).max()) < 1e-13
    # linearity on a slice: F(a x0 + x1) = a F(x0) + F(x1)
    a = 0.75
    lin = (a * x[:1024] + x[1024:2048]).contiguous()
    plan.fwd(lin)
    want = a * y[:1024] + y[1024:2048]
    rel = (lin - want).abs().pow(2).sum(1).sqrt() / want.abs().pow(2).sum(1).sqrt()
    assert float(rel.max()) < 1e-13 * 11
    # round trip
    plan.inv(y)
    torch.cuda.synchronize()
    err = float((y / n - x).abs().max())
    assert err < 1e-12


@pytest.mark.parametrize("npoly,base_n", [(64, 32), (1024, 256), (4096, 256), (4096, 32)])
def test_negacyclic_product_through_unordered_plan(C, torch, npoly, base_n):
    """What the crate is for (README.md:10-17, src/lib.rs:9-16): polynomial products modulo X^N + 1 with
    only element-wise work in the (permuted) Fourier domain.  Fold N real coefficients into N/2 complex
    points, twist by e^{i pi k / N}, unordered fwd, point-wise product on the device, unordered inv,
    untwist, unfold -- equals the integer schoolbook negacyclic convolution.  The c64 counterpart of the
    reference's fft128 `test_product` (src/fft128/mod.rs:1990-2066)."""
    n = npoly // 2
    rng = np.random.default_rng(npoly + base_n)
    batch = 6
    a = rng.integers(-(1 << 20), 1 << 20, size=(batch, npoly))
    b = rng.integers(-(1 << 10), 1 << 10, size=(batch, npoly))
    # schoolbook in exact integer arithmetic: c_k = sum_{i+j=k} a_i b_j - sum_{i+j=k+N} a_i b_j
    want = np.zeros((batch, npoly), dtype=object)
    for r in range(batch):
        full = np.convolve(a[r].astype(object), b[r].astype(object))
        want[r] = full[:npoly]
        want[r][: npoly - 1] -= full[npoly:]
    plan = C.unordered.Plan(n, C.unordered.Method.UserProvided(C.ordered.FftAlgo.Dif16, min(base_n, n)))
    # fold, integer -> f64 conversion and twist ride on the forward transform's loads (cfft_c64_poly_fwd) ...
    fa = plan.fwd_poly(torch.from_numpy(a).cuda())
    fb = plan.fwd_poly(torch.from_numpy(b).cuda())
    C.pointwise.mul_assign(fa, fb)  # same permutation on both operands: the order never has to be undone
    # ... and untwist, 1 / n and the rounding on the inverse transform's stores (cfft_c64_poly_inv): integers out
    got = plan.inv_poly(fa)
    torch.cuda.synchronize()
    assert np.array_equal(got.cpu().numpy(), want.astype(np.int64))  # |c| < 2^42: far inside f64


def test_pointwise_products_bit_exact(C, torch):
    """cfft_c64_mul_assign / cfft_c64_mul_add_assign against the oracle's restatement of num_complex's `*`
    and `+` (no FMA), on random data of ragged lengths plus signed zeros, infinities and NaN."""
    rng = np.random.default_rng(31)
    for length in (1, 7, 4096, 1000003):
        a = rand_c(rng, length) - (0.5 + 0.5j)
        b = rand_c(rng, length) * 1e3
        acc = rand_c(rng, length)
        if length >= 7:
            a[:6] = [0.0, -0.0, complex(np.inf, 1.0), complex(1.0, -np.inf), complex(np.nan, 0.0), 1e308 + 1e308j]
            b[:6] = [-1.0 + 0.0j, complex(0.0, -0.0), 0.0j, 2.0 + 3.0j, 1.0 + 1.0j, 10.0 + 10.0j]
        da, db, dc = (torch.from_numpy(v.copy()).cuda() for v in (a, b, acc))
        C.pointwise.mul_add_assign(dc, da, db)
        C.pointwise.mul_assign(da, db)
        torch.cuda.synchronize()
        def same(got, want):  # bit-exact, except that NaN payloads / signs are not IEEE-specified (x86 vs GPU differ)
            g, w = got.view(np.float64), want.view(np.float64)
            nan = np.isnan(w)
            return np.array_equal(np.isnan(g), nan) and np.array_equal(g[~nan].view(np.uint64), w[~nan].view(np.uint64))

        with np.errstate(all="ignore"):
            assert same(da.cpu().numpy(), O.c64_pointwise(a, b)), length
            assert same(dc.cpu().numpy(), O.c64_pointwise(a, b, acc)), length
    with pytest.raises(C.PanicError):
        C.pointwise.mul_assign(da, db[:5])


def _oracle_fwd_mul_inv(ref, a, b):
    """inv(sum_k fwd(a[r, k]) * b[r or shared, k]) composed from the oracle's pieces, terms added in order."""
    batch, k, n = a.shape
    out = np.empty((batch, n), np.complex128)
    for r in range(batch):
        br = b if b.ndim == 2 else b[r]
        acc = O.c64_pointwise(ref.fwd(a[r, 0]), br[0])
        for j in range(1, k):
            acc = O.c64_pointwise(ref.fwd(a[r, j]), br[j], acc)
        out[r] = ref.inv(acc)
    return out


@pytest.mark.parametrize("n", [256, 512, 1024, 2048, 4096, 8192])
def test_fwd_mul_inv_fused_kernel_bit_exact(C, torch, n):
    """cfft_c64_fwd_mul_inv on plans with the one-kernel path (c64_fwd_mul_inv_kernel): bit-identical to the oracle's
    fwd -> num_complex product / sum -> inv, to the composition of the library's own calls and to the composed
    device path, for 1..5 terms, whole and ragged CTA tiles, b shared by the batch or per row, and in place."""
    rng = np.random.default_rng(900 + n)
    plan = C.unordered.Plan(n, C.unordered.Method.UserProvided(C.ordered.FftAlgo.Dif16, 256))
    assert plan.has_fused_mul_kernel()
    ref = O.UnorderedPlan(n, O.DIF16, 256)
    for batch, k, shared in [(1, 1, True), (7, 1, False), (8, 2, True), (5, 3, False), (33, 2, True), (3, 5, True)]:
        a = rand_c(rng, batch, k, n) - (0.5 + 0.5j)
        b = (rand_c(rng, k, n) if shared else rand_c(rng, batch, k, n)) - (0.5 + 0.5j)
        want = _oracle_fwd_mul_inv(ref, a, b)
        da, db = torch.from_numpy(a).cuda(), torch.from_numpy(b).cuda()
        launches = C._native.launch_count()
        got = plan.fwd_mul_inv(da, db)
        assert C._native.launch_count() - launches == (1 if n <= 4096 else k)  # one kernel (n = 8192: one per term, chained through out)
        torch.cuda.synchronize()
        assert bits_equal(got.cpu().numpy(), want), (n, batch, k, shared)
        assert bits_equal(da.cpu().numpy(), a)  # inputs untouched
        # the library's separate calls
        fa = da.clone()
        plan.fwd(fa)
        bb = db if not shared else db.unsqueeze(0).expand(batch, k, n).contiguous()
        acc = fa[:, 0].contiguous()
        C.pointwise.mul_assign(acc, bb[:, 0].contiguous())
        for j in range(1, k):
            C.pointwise.mul_add_assign(acc, fa[:, j].contiguous(), bb[:, j].contiguous())
        plan.inv(acc)
        torch.cuda.synchronize()
        assert bits_equal(acc.cpu().numpy(), want), (n, batch, k, shared)
        # the composed device path (what plans without the fused kernel run)
        os.environ["CFFT_B200_FUSED_MUL_COMPOSED"] = "1"
        try:
            got2 = plan.fwd_mul_inv(da, db)
        finally:
            del os.environ["CFFT_B200_FUSED_MUL_COMPOSED"]
        torch.cuda.synchronize()
        assert bits_equal(got2.cpu().numpy(), want), (n, batch, k, shared)
        if k == 1:  # in place
            inplace = da.clone().reshape(batch, n)
            assert plan.fwd_mul_inv(inplace, db, out=inplace) is inplace
            torch.cuda.synchronize()
            assert bits_equal(inplace.cpu().numpy(), want), (n, batch)


@pytest.mark.parametrize("kind,n,algo,base_n", [("unordered", 2048, "Dif4", 32), ("unordered", 8192, "Dif16", 256),
                                                ("unordered", 16384, "Dif16", 256), ("unordered", 64, "Dit8", 64),
                                                ("ordered", 512, "Dif8", 0), ("ordered", 256, "Dif16", 0)])
def test_fwd_mul_inv_any_plan_bit_exact(C, torch, kind, n, algo, base_n):
    """cfft_c64_fwd_mul_inv on plans without the one-kernel path (and the ordered 256-point plan, which has it):
    the same bits as the oracle composition."""
    rng = np.random.default_rng(n + base_n)
    A = getattr(C.ordered.FftAlgo, algo)
    oa = getattr(O, algo.upper())
    if kind == "unordered":
        plan, ref = C.unordered.Plan(n, C.unordered.Method.UserProvided(A, base_n)), O.UnorderedPlan(n, oa, base_n)
    else:
        plan, ref = C.ordered.Plan(n, C.ordered.Method.UserProvided(A)), O.OrderedPlan(n, oa)
    assert plan.has_fused_mul_kernel() == (n in (256, 8192))
    for batch, k, shared in [(3, 1, True), (5, 3, False), (2, 2, True)]:
        a = rand_c(rng, batch, k, n) - (0.5 + 0.5j)
        b = (rand_c(rng, k, n) if shared else rand_c(rng, batch, k, n)) - (0.5 + 0.5j)
        got = plan.fwd_mul_inv(torch.from_numpy(a).cuda(), torch.from_numpy(b).cuda())
        torch.cuda.synchronize()
        assert bits_equal(got.cpu().numpy(), _oracle_fwd_mul_inv(ref, a, b)), (kind, n, batch, k, shared)


@pytest.mark.parametrize("n,algo,base_n", [(256, "Dif16", 256), (1024, "Dif16", 256), (2048, "Dif16", 256), (4096, "Dif16", 256),
                                           (8192, "Dif16", 256), (2048, "Dif4", 32), (16384, "Dif16", 256)])
def test_fwd_mul_add_bit_exact(C, torch, n, algo, base_n):
    """cfft_c64_fwd_mul_add (forward transform + multiply[-accumulate] into a Fourier-domain accumulator, no inverse): term by
    term it builds exactly the oracle's sum, and cfft_c64_inv of it equals cfft_c64_fwd_mul_inv; strided views of a
    [batch, k, n] array as inputs, b shared or per row, one kernel per call where the plan has the fused kernel."""
    rng = np.random.default_rng(31 * n + base_n)
    plan = C.unordered.Plan(n, C.unordered.Method.UserProvided(getattr(C.ordered.FftAlgo, algo), base_n))
    ref = O.UnorderedPlan(n, getattr(O, algo.upper()), base_n)
    for batch, k, shared in [(5, 3, True), (2, 2, False), (1, 1, True)]:
        a = rand_c(rng, batch, k, n) - (0.5 + 0.5j)
        b = (rand_c(rng, k, n) if shared else rand_c(rng, batch, k, n)) - (0.5 + 0.5j)
        want = np.empty((batch, n), np.complex128)
        for r in range(batch):
            br = b if shared else b[r]
            acc = O.c64_pointwise(ref.fwd(a[r, 0]), br[0])
            for j in range(1, k):
                acc = O.c64_pointwise(ref.fwd(a[r, j]), br[j], acc)
            want[r] = acc
        da, db = torch.from_numpy(a).cuda(), torch.from_numpy(b).cuda()
        dacc = torch.full((batch, n), float("nan"), dtype=torch.complex128, device="cuda")
        launches = C._native.launch_count()
        for j in range(k):
            plan.fwd_mul_add(da[:, j], db[j] if shared else db[:, j], dacc, accumulate=j > 0)
        if plan.has_fused_mul_kernel():
            assert C._native.launch_count() - launches == k
        torch.cuda.synchronize()
        assert bits_equal(dacc.cpu().numpy(), want), (n, batch, k, shared)
        assert bits_equal(da.cpu().numpy(), a)
        plan.inv(dacc)
        full = plan.fwd_mul_inv(da, db)
        torch.cuda.synchronize()
        assert bits_equal(dacc.cpu().numpy(), full.cpu().numpy()), (n, batch, k, shared)
    with pytest.raises(C.PanicError):
        plan.fwd_mul_add(da[:, 0], db[0], dacc[:, : n // 2].contiguous())
    with pytest.raises(C.PanicError):  # acc aliasing an input
        C._native.check(C._native.lib.cfft_c64_fwd_mul_add(plan._h, dacc.data_ptr(), n, db.data_ptr(), 0, dacc.data_ptr(), 0, 1, 0))


def test_fwd_mul_inv_negacyclic_external_product(C, torch):
    """The shape the call exists for: out = sum_k a_k * b_k modulo X^N + 1 (N = 4096, fft size 2048, k = 4 terms, b
    shared by the batch like a bootstrapping-key GGSW row), one call on the device, against the exact integer
    schoolbook product."""
    npoly, n, k, batch = 4096, 2048, 4, 5
    rng = np.random.default_rng(77)
    a = rng.integers(-(1 << 16), 1 << 16, size=(batch, k, npoly))
    b = rng.integers(-(1 << 10), 1 << 10, size=(k, npoly))
    want = np.zeros((batch, npoly), dtype=object)
    for r in range(batch):
        for j in range(k):
            full = np.convolve(a[r, j].astype(object), b[j].astype(object))
            want[r] += full[:npoly]
            want[r][: npoly - 1] -= full[npoly:]
    twist = np.exp(1j * np.pi * np.arange(n) / npoly)
    fold = lambda p: (p[..., :n] + 1j * p[..., n:]) * twist
    plan = C.unordered.Plan(n, C.unordered.Method.Measure())
    assert plan.has_fused_mul_kernel()
    fb = torch.from_numpy(fold(b)).cuda()
    plan.fwd(fb)  # the key is kept in the Fourier domain, in this plan's order
    out = plan.fwd_mul_inv(torch.from_numpy(fold(a)).cuda(), fb)
    torch.cuda.synchronize()
    z = out.cpu().numpy() / n * np.conj(twist)
    got = np.concatenate([z.real, z.imag], axis=1)
    assert np.array_equal(np.rint(got).astype(np.int64), want.astype(np.int64))


def test_fwd_mul_inv_full_size_matches_separate_calls(C, torch):
    """BASELINE.json configs[1] shape through the fused call: 16384 rows of N = 2048 with 2 terms each against the
    library's separate fwd / product / inv launches, bit for bit (the oracle pins those at small sizes)."""
    n, k, batch = 2048, 2, 16384
    plan = C.unordered.Plan(n, C.unordered.Method.UserProvided(C.ordered.FftAlgo.Dif16, 256))
    g = torch.Generator(device="cuda").manual_seed(5)
    a = torch.view_as_complex(torch.rand((batch, k, n, 2), generator=g, device="cuda", dtype=torch.float64) - 0.5)
    b = torch.view_as_complex(torch.rand((k, n, 2), generator=g, device="cuda", dtype=torch.float64) - 0.5)
    got = plan.fwd_mul_inv(a, b)
    fa = a.clone()
    plan.fwd(fa)
    acc = fa[:, 0].contiguous()
    C.pointwise.mul_assign(acc, b[0].expand(batch, n).contiguous())
    C.pointwise.mul_add_assign(acc, fa[:, 1].contiguous(), b[1].expand(batch, n).contiguous())
    plan.inv(acc)
    torch.cuda.synchronize()
    assert torch.equal(torch.view_as_real(got).view(torch.int64), torch.view_as_real(acc).view(torch.int64))


def test_fwd_mul_inv_argument_checks(C, torch):
    plan = C.unordered.Plan(512, C.unordered.Method.UserProvided(C.ordered.FftAlgo.Dif16, 256))
    a = torch.zeros((4, 2, 512), dtype=torch.complex128, device="cuda")
    b = torch.zeros((2, 512), dtype=torch.complex128, device="cuda")
    with pytest.raises(C.PanicError):
        plan.fwd_mul_inv(a, b[:1].contiguous())  # wrong number of terms
    with pytest.raises(C.PanicError):
        plan.fwd_mul_inv(a[:, :, :256].contiguous(), b)  # wrong fft size
    with pytest.raises(C.PanicError):
        plan.fwd_mul_inv(a, b, out=torch.zeros((3, 512), dtype=torch.complex128, device="cuda"))
    with pytest.raises(C.PanicError):  # aliasing a with more than one term
        C._native.check(C._native.lib.cfft_c64_fwd_mul_inv(plan._h, a.data_ptr(), 2, b.data_ptr(), 0, a.data_ptr(), 4, 0))
    with pytest.raises(C.PanicError):  # zero terms
        C._native.check(C._native.lib.cfft_c64_fwd_mul_inv(plan._h, a.data_ptr(), 0, b.data_ptr(), 0, a.data_ptr(), 4, 0))
    f128 = C.fft128.Plan(64)
    with pytest.raises(C.PanicError):
        C._native.check(C._native.lib.cfft_c64_fwd_mul_inv(f128._h, a.data_ptr(), 1, b.data_ptr(), 0, a.data_ptr(), 1, 0))


@pytest.mark.parametrize("n", [256, 512, 1024, 2048, 4096, 8192])
def test_fast_register_kernel_bit_exact(C, torch, n):
    """c64_fast.cu (plans with base (Dif16, 256)): same bits and order as the reference plan, for
    whole and ragged CTA tiles, and identical to the exact tile kernel."""
    rng = np.random.default_rng(n)
    plan = C.unordered.Plan(n, C.unordered.Method.UserProvided(C.ordered.FftAlgo.Dif16, 256))
    assert plan.kernel_name() == "fast-b256-regs"
    ref = O.UnorderedPlan(n, O.DIF16, 256)
    for batch in [1, 7, 8, 33]:
        x = rand_c(rng, batch, n)
        y = dev_run(torch, plan.fwd, x)
        want = ref.fwd(x)
        assert bits_equal(y, want), (n, batch)
        assert bits_equal(dev_run(torch, plan.inv, y), ref.inv(want)), (n, batch)
    os.environ["CFFT_B200_FORCE_EXACT"] = "1"
    try:
        exact = C.unordered.Plan(n, C.unordered.Method.UserProvided(C.ordered.FftAlgo.Dif16, 256))
    finally:
        del os.environ["CFFT_B200_FORCE_EXACT"]
    assert exact.kernel_name() == "exact-regs"
    assert bits_equal(dev_run(torch, exact.fwd, x), y)


@pytest.mark.parametrize("logn", [14, 15, 16, 17, 18, 19, 20, 21, 22])
def test_fast_large_n_column_passes_bit_exact(C, torch, logn):
    """n = 2^14 .. 2^20 with base (Dif16, 256): levels as column passes (c64_column.cu) + base FFTs on
    rows; same bits and same permuted order as the reference plan."""
    n = 1 << logn
    rng = np.random.default_rng(logn)
    plan = C.unordered.Plan(n, C.unordered.Method.UserProvided(C.ordered.FftAlgo.Dif16, 256))
    # default: base FFTs on rows of 256 up to 2^16; from 2^17 the last levels + base FFTs run as the fused
    # kernel of 512 .. 4096 points, which saves one HBM pass
    assert plan.kernel_name() == ("fast-b256-column+rows" if logn <= 16 else "fast-b256-column+fused-rows")
    plans = [plan]
    for var, name in (("2", "fast-b256-column+rows"), ("9", "fast-b256-column+fused-rows")):
        os.environ["CFFT_B200_FAST_VARIANT"] = var
        try:
            plans.append(C.unordered.Plan(n, C.unordered.Method.UserProvided(C.ordered.FftAlgo.Dif16, 256)))
        finally:
            del os.environ["CFFT_B200_FAST_VARIANT"]
        assert plans[-1].kernel_name() == name
    ref = O.UnorderedPlan(n, O.DIF16, 256)
    for batch in ([1, 3] if logn <= 17 else [2]):
        x = rand_c(rng, batch, n)
        want = ref.fwd(x, threads=8)
        back = ref.inv(want, threads=8)
        for pl in plans:
            y = dev_run(torch, pl.fwd, x)
            assert bits_equal(y, want), (n, batch, pl.kernel_name())
            assert bits_equal(dev_run(torch, pl.inv, y), back), (n, batch, pl.kernel_name())
    pi = plan.permutation().astype(np.int64)
    f = np.fft.fft(x[0])
    assert np.linalg.norm(y[0][pi] - f) / np.linalg.norm(f) <= 1e-13 * logn


@pytest.mark.parametrize("logn", [11, 12, 13, 14, 15, 16, 17, 18, 19, 20])
def test_ordered_above_reference_cap(C, torch, logn):
    """BASELINE configs[2]: standard-order transforms for n > 2^10.  The reference cannot build these
    (src/ordered.rs:244), so the oracle is the DFT definition: the unordered reference plan
    (Dif16, 256) un-permuted (SURVEY.md 8c caveat) -- bit-exact -- and numpy's FFT within tolerance."""
    n = 1 << logn
    rng = np.random.default_rng(1000 + logn)
    with pytest.raises(C.PanicError):
        C.ordered.Plan(n, C.ordered.Method.UserProvided(C.ordered.FftAlgo.Dif16))  # reference behaviour
    plan = C.ordered.Plan(n, C.ordered.Method.Measure(), allow_large=True)
    want_kernel = "ordered-b256-regs-std" if logn <= 13 else "ordered-b256-column+rows-std"
    assert plan.fft_size() == n and plan.kernel_name().startswith(("ordered-b256-regs-std", "ordered-b256-column+rows-std"))
    assert C.ordered.Plan(n, C.ordered.Method.UserProvided(C.ordered.FftAlgo.Dif16), allow_large=True).kernel_name() == want_kernel
    ref = O.UnorderedPlan(n, O.DIF16, 256)
    pi = O.permutation(n, 256)
    batch = 3 if logn <= 16 else 2
    x = rand_c(rng, batch, n)
    y = dev_run(torch, plan.fwd, x)
    want = ref.fwd(x, threads=8)[:, pi]
    assert bits_equal(y, want)
    f = np.fft.fft(x, axis=1)
    assert (np.linalg.norm(y - f, axis=1) / np.linalg.norm(f, axis=1)).max() <= 1e-13 * logn
    z = dev_run(torch, plan.inv, y)
    perm_in = np.empty_like(y)
    perm_in[:, pi] = y
    assert bits_equal(z, ref.inv(perm_in, threads=8))
    assert np.abs(z / n - x).max() < 1e-11
    h = x.copy()
    plan.fwd(h)  # host-memory entry
    assert bits_equal(h, want)


@pytest.mark.parametrize("n", [2048, 4096, 8192])
def test_ordered_fused_standard_order_kernel(C, torch, n):
    """Standard-order plans 2^11 <= n <= 2^13 run as ONE kernel (levels + base FFTs + the transposing
    exchange in shared memory).  Same bits as the unordered reference plan un-permuted, for ragged
    batches, in place, fwd and inv; and the same bits as the multi-pass variant."""
    rng = np.random.default_rng(1500 + n)
    plan = C.ordered.Plan(n, C.ordered.Method.UserProvided(C.ordered.FftAlgo.Dif16), allow_large=True)
    assert plan.kernel_name() == "ordered-b256-regs-std"
    ref = O.UnorderedPlan(n, O.DIF16, 256)
    pi = O.permutation(n, 256)
    for batch in (1, 5, 37):
        x = rand_c(rng, batch, n)
        y = dev_run(torch, plan.fwd, x)
        assert bits_equal(y, ref.fwd(x, threads=8)[:, pi]), (n, batch)
        perm_in = np.empty_like(y)
        perm_in[:, pi] = y
        assert bits_equal(dev_run(torch, plan.inv, y), ref.inv(perm_in, threads=8)), (n, batch)
    os.environ["CFFT_B200_FAST_VARIANT"] = "3"
    try:
        multi = C.ordered.Plan(n, C.ordered.Method.UserProvided(C.ordered.FftAlgo.Dif16), allow_large=True)
    finally:
        del os.environ["CFFT_B200_FAST_VARIANT"]
    assert multi.kernel_name() == "ordered-b256-column+rows-std"
    x = rand_c(rng, 9, n)
    assert bits_equal(dev_run(torch, multi.fwd, x), dev_run(torch, plan.fwd, x))
    assert bits_equal(dev_run(torch, multi.inv, x), dev_run(torch, plan.inv, x))


@pytest.mark.parametrize("n", [16, 32, 64, 128, 512, 1024])
def test_ord16_register_kernel_bit_exact(C, torch, n):
    """Whole-transform Dif16 plans (ordered, and unordered with base_n == n) on the register kernel of
    c64_ord16.cu: bit-exact vs the oracle for ragged batches (tail CTAs, single rows), fwd and inv,
    and identical to the exact tile kernel."""
    rng = np.random.default_rng(1600 + n)
    A = C.ordered.FftAlgo
    po = C.ordered.Plan(n, C.ordered.Method.UserProvided(A.Dif16))
    pu = C.unordered.Plan(n, C.unordered.Method.UserProvided(A.Dif16, n))
    assert po.kernel_name() == "ord16-regs" and pu.kernel_name() == "ord16-regs"
    assert C.ordered.Plan(n, C.ordered.Method.UserProvided(A.Dit16)).kernel_name() == ("exact-regs-spec" if n == 1024 else "exact-regs")
    assert np.array_equal(pu.permutation(), np.arange(n))
    ref = O.OrderedPlan(n, O.DIF16)
    uref = O.UnorderedPlan(n, O.DIF16, n)
    for batch in (1, 2, 3, 63, 64, 65, 1000):
        x = rand_c(rng, batch, n)
        y = dev_run(torch, po.fwd, x)
        want = ref.fwd(x)
        assert bits_equal(y, want), (n, batch)
        assert bits_equal(want, uref.fwd(x))
        assert bits_equal(dev_run(torch, pu.fwd, x), want), (n, batch)
        z = dev_run(torch, po.inv, y)
        assert bits_equal(z, ref.inv(y)), (n, batch)
        assert bits_equal(dev_run(torch, pu.inv, y), z)
        assert np.abs(z / n - x).max() < 1e-12
    f = np.fft.fft(x, axis=1)
    assert (np.linalg.norm(y - f, axis=1) / np.linalg.norm(f, axis=1)).max() <= 1e-13 * np.log2(n)
    os.environ["CFFT_B200_FORCE_EXACT"] = "1"
    try:
        exact = C.ordered.Plan(n, C.ordered.Method.UserProvided(A.Dif16))
    finally:
        del os.environ["CFFT_B200_FORCE_EXACT"]
    assert exact.kernel_name() == "exact-regs"
    assert bits_equal(dev_run(torch, exact.fwd, x), y)
    h = x.copy()
    po.fwd(h)  # host-memory entry
    assert bits_equal(h, y)


@pytest.mark.parametrize("algo", range(8))
def test_generic_register_kernel_matches_tile_kernel_and_oracle(C, torch, algo):
    """c64_regs.cu (any plan, 16 c64 per thread, planar twiddles) against the oracle AND against the
    shared-memory tile kernel it replaces (CFFT_B200_EXACT_TILE=1), ragged batches that leave partial
    tiles, n below / at / above the tile, ordered and unordered."""
    rng = np.random.default_rng(1700 + algo)
    A = C.ordered.FftAlgo
    cases = [(64, 64), (256, 32), (1024, 1024), (2048, 512), (4096, 128), (8192, 1024), (32768, 64)]
    for n, base_n in cases:
        plan = C.unordered.Plan(n, C.unordered.Method.UserProvided(A(algo), base_n))
        if not (algo == O.DIF16 and (base_n == 256 or base_n == n)):
            assert plan.kernel_name() in ("exact-regs", "exact-regs-spec"), (algo, n, base_n, plan.kernel_name())
        ref = O.UnorderedPlan(n, algo, base_n)
        for batch in (1, 3, 2048 // min(n, 2048) + 1):
            x = rand_c(rng, batch, n)
            y = dev_run(torch, plan.fwd, x)
            want = ref.fwd(x, threads=8)
            assert bits_equal(y, want), (algo, n, base_n, batch)
            assert bits_equal(dev_run(torch, plan.inv, y), ref.inv(want, threads=8)), (algo, n, base_n, batch)
            os.environ["CFFT_B200_EXACT_TILE"] = "1"
            try:
                assert bits_equal(dev_run(torch, plan.fwd, x), y)
            finally:
                del os.environ["CFFT_B200_EXACT_TILE"]
    for n in (2, 4, 8, 16, 32, 512, 1024):
        plan = C.ordered.Plan(n, C.ordered.Method.UserProvided(A(algo)))
        x = rand_c(rng, 131, n)
        assert bits_equal(dev_run(torch, plan.fwd, x), O.OrderedPlan(n, algo).fwd(x)), (algo, n)


SPEC_PLANS = [(2048, O.DIF16, 1024), (2048, O.DIF16, 512), (2048, O.DIF8, 512), (2048, O.DIF4, 32), (2048, O.DIT16, 1024),
              (1024, O.DIF16, 512), (1024, O.DIF8, 512), (4096, O.DIF16, 1024), (4096, O.DIF8, 512),
              (2048, O.DIT8, 512), (2048, O.DIT16, 512), (1024, O.DIT16, 512), (1024, O.DIT8, 512), (4096, O.DIT16, 1024), (4096, O.DIT8, 512)]


@pytest.mark.parametrize("n,algo,base_n", SPEC_PLANS)
def test_compile_time_schedule_kernels_bit_exact(C, torch, n, algo, base_n):
    """c64_regs_spec_kernel: the plans the reference's own Method::Measure tends to produce (src/unordered.rs:568-630) and the
    golden-vector plan with their stage schedule built at compile time -- against the oracle and against the interpreter
    (CFFT_B200_REGS_NO_SPEC=1), whole and ragged batches, fwd and inv."""
    rng = np.random.default_rng(4100 + n + algo + base_n)
    A = C.ordered.FftAlgo
    plan = C.unordered.Plan(n, C.unordered.Method.UserProvided(A(algo), base_n))
    assert plan.kernel_name() == "exact-regs-spec"
    ref = O.UnorderedPlan(n, algo, base_n)
    for batch in (1, 2, 37):
        x = rand_c(rng, batch, n)
        y = dev_run(torch, plan.fwd, x)
        want = ref.fwd(x, threads=8)
        assert bits_equal(y, want), (n, algo, base_n, batch)
        back = dev_run(torch, plan.inv, y)
        assert bits_equal(back, ref.inv(want, threads=8)), (n, algo, base_n, batch)
        os.environ["CFFT_B200_REGS_NO_SPEC"] = "1"
        try:
            assert bits_equal(dev_run(torch, plan.fwd, x), y) and bits_equal(dev_run(torch, plan.inv, y), back)
        finally:
            del os.environ["CFFT_B200_REGS_NO_SPEC"]


@pytest.mark.parametrize("n,algo", [(1024, O.DIF8), (1024, O.DIT8), (1024, O.DIT16), (512, O.DIF8), (512, O.DIT8), (1024, O.DIF4)])
def test_compile_time_schedule_ordered_plans_bit_exact(C, torch, n, algo):
    """The same for whole-transform plans: ordered::Plan with a radix-4 / 8 algorithm or Dit16 (Dif16 has its own kernels), and
    the unordered plan with base_n == n, whose output order is the natural one."""
    rng = np.random.default_rng(4200 + n + algo)
    A = C.ordered.FftAlgo
    po = C.ordered.Plan(n, C.ordered.Method.UserProvided(A(algo)))
    pu = C.unordered.Plan(n, C.unordered.Method.UserProvided(A(algo), n))
    assert po.kernel_name() == "exact-regs-spec" and pu.kernel_name() == "exact-regs-spec"
    ref = O.OrderedPlan(n, algo)
    for batch in (1, 3, 129):
        x = rand_c(rng, batch, n)
        y = dev_run(torch, po.fwd, x)
        want = ref.fwd(x)
        assert bits_equal(y, want) and bits_equal(dev_run(torch, pu.fwd, x), want), (n, algo, batch)
        back = dev_run(torch, po.inv, y)
        assert bits_equal(back, ref.inv(want)), (n, algo, batch)
        os.environ["CFFT_B200_REGS_NO_SPEC"] = "1"
        try:
            assert bits_equal(dev_run(torch, po.fwd, x), y) and bits_equal(dev_run(torch, po.inv, y), back)
        finally:
            del os.environ["CFFT_B200_REGS_NO_SPEC"]


def test_autotune_keeps_bits_and_order(C, torch):
    """cfft_plan_autotune (the Method::Measure replacement) switches kernel VARIANTS only: every
    variant of a plan gives the same bits in the same order."""
    rng = np.random.default_rng(77)
    A = C.ordered.FftAlgo
    for n in [512, 2048, 8192]:
        x = rand_c(rng, 5, n)
        want = O.UnorderedPlan(n, O.DIF16, 256).fwd(x)
        os.environ["CFFT_B200_FAST_VARIANT"] = "2"
        try:
            multi = C.unordered.Plan(n, C.unordered.Method.UserProvided(A.Dif16, 256))
        finally:
            del os.environ["CFFT_B200_FAST_VARIANT"]
        assert multi.kernel_name() == "fast-b256-column+rows"
        y = dev_run(torch, multi.fwd, x)
        assert bits_equal(y, want)
        assert bits_equal(dev_run(torch, multi.inv, y), O.UnorderedPlan(n, O.DIF16, 256).inv(want))
        report = multi.autotune()
        assert "selected:" in report and "fast-b256-regs" in report and "fast-b256-column+rows" in report
        assert multi.algo() == (A.Dif16, 256)
        assert bits_equal(dev_run(torch, multi.fwd, x), want)
        twin = multi.clone()
        assert twin.kernel_name() == multi.kernel_name()
    # tile-size variants of the exact kernel and of fft128
    p = C.unordered.Plan(128, C.unordered.Method.Measure())
    assert "exact-tile/" in p.tuning_report() and "exact-regs" in p.tuning_report() and "selected:" in p.tuning_report()
    x = rand_c(rng, 300, 128)
    assert bits_equal(dev_run(torch, p.fwd, x), O.UnorderedPlan(128, O.DIF16, 128).fwd(x))
    fp = C.fft128.Plan(256)
    rep = fp.autotune()
    assert "f128-radix8-tile/" in rep and "selected:" in rep
    planes = [rng.random((37, 256)), np.zeros((37, 256)), rng.random((37, 256)), np.zeros((37, 256))]
    d = [torch.from_numpy(a.copy()).cuda() for a in planes]
    fp.fwd(*d)
    torch.cuda.synchronize()
    ref = O.F128Plan(256).fwd(*planes, variant=O.F128_FMA)
    assert all(np.array_equal(a.cpu().numpy().view(np.uint64), b.view(np.uint64)) for a, b in zip(d, ref))


def test_concurrent_host_calls_share_one_plan(C, torch):
    """`&self` semantics: many host threads call fwd on ONE plan at the same time (TFHE-rs keeps plans
    in a global cache and calls them from worker threads)."""
    import threading

    rng = np.random.default_rng(55)
    n = 1024
    plan = C.unordered.Plan(n, C.unordered.Method.UserProvided(C.ordered.FftAlgo.Dif16, 256))
    ref = O.UnorderedPlan(n, O.DIF16, 256)
    inputs = [rand_c(rng, 1 + 37 * i, n) for i in range(8)]
    wants = [ref.fwd(x) for x in inputs]
    outs = [x.copy() for x in inputs]
    errors = []

    def work(i):
        try:
            for _ in range(3):
                outs[i][...] = inputs[i]
                plan.fwd(outs[i])
        except Exception as e:  # pragma: no cover
            errors.append(e)

    threads = [threading.Thread(target=work, args=(i,)) for i in range(8)]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    assert not errors
    for got, want in zip(outs, wants):
        assert bits_equal(got, want)


def test_plan_on_second_device(C, torch):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    rng = np.random.default_rng(56)
    n = 2048
    plan = C.unordered.Plan(n, C.unordered.Method.UserProvided(C.ordered.FftAlgo.Dif16, 256), device=1)
    assert plan.device() == 1
    x = rand_c(rng, 4, n)
    d = torch.from_numpy(x.copy()).to("cuda:1")
    plan.fwd(d)  # current device stays cuda:0
    torch.cuda.synchronize(1)
    assert torch.cuda.current_device() == 0
    assert bits_equal(d.cpu().numpy(), O.UnorderedPlan(n, O.DIF16, 256).fwd(x))
    with pytest.raises(ValueError):
        plan.fwd(torch.from_numpy(x.copy()).to("cuda:0"))
    h = x.copy()
    plan.fwd(h)
    assert bits_equal(h, d.cpu().numpy())


@pytest.mark.parametrize("n", [8192, 16384])
def test_cluster_kernel_bit_exact(C, torch, n):
    """One transform per thread-block cluster (DSMEM exchange after the first level): same bits."""
    rng = np.random.default_rng(n + 1)
    os.environ["CFFT_B200_FAST_VARIANT"] = "4"
    try:
        plan = C.unordered.Plan(n, C.unordered.Method.UserProvided(C.ordered.FftAlgo.Dif16, 256))
    finally:
        del os.environ["CFFT_B200_FAST_VARIANT"]
    assert plan.kernel_name() == "fast-b256-cluster"
    ref = O.UnorderedPlan(n, O.DIF16, 256)
    for batch in [1, 2, 37]:
        x = rand_c(rng, batch, n)
        y = dev_run(torch, plan.fwd, x)
        want = ref.fwd(x, threads=8)
        assert bits_equal(y, want), (n, batch)
        assert bits_equal(dev_run(torch, plan.inv, y), ref.inv(want, threads=8)), (n, batch)
    assert "fast-b256-cluster" in plan.autotune()


@pytest.mark.parametrize("logn", [14, 15, 16])
def test_persistent_two_phase_kernel_bit_exact(C, torch, logn):
    """n = 2^14 .. 2^16 with both HBM passes in one persistent kernel (work queue, per-transform
    release / acquire counters): same bits as the reference plan for batches smaller than, equal to and
    larger than the lag between the phases, and larger than the number of resident CTAs' items."""
    n = 1 << logn
    rng = np.random.default_rng(logn)
    os.environ["CFFT_B200_FAST_VARIANT"] = "8"
    try:
        plan = C.unordered.Plan(n, C.unordered.Method.UserProvided(C.ordered.FftAlgo.Dif16, 256))
    finally:
        del os.environ["CFFT_B200_FAST_VARIANT"]
    assert plan.kernel_name() == "fast-b256-persistent-2pass"
    ref = O.UnorderedPlan(n, O.DIF16, 256)
    for batch in [1, 2, 37, 1 << (22 - logn)]:
        x = rand_c(rng, batch, n)
        y = dev_run(torch, plan.fwd, x)
        want = ref.fwd(x, threads=8)
        assert bits_equal(y, want), (n, batch)
        assert bits_equal(dev_run(torch, plan.inv, y), ref.inv(want, threads=8)), (n, batch)
    # opt-in only (it spin-waits on other CTAs): never an autotune candidate unless the caller allows it
    assert "fast-b256-persistent-2pass" not in plan.autotune()
    os.environ["CFFT_B200_ALLOW_PERSISTENT"] = "1"
    try:
        assert "fast-b256-persistent-2pass" in plan.autotune()
    finally:
        del os.environ["CFFT_B200_ALLOW_PERSISTENT"]
    lib = C._native.lib
    import ctypes
    count = ctypes.c_uint32(123)
    C._native.check(lib.cfft_twopass_timeouts(0, ctypes.byref(count)))
    assert count.value == 0  # no wait ever expired


def test_persistent_kernels_share_the_gpu_without_deadlock(C, torch):
    """Several persistent two-phase kernels (each sized to fill the GPU, each with spin-waits on its own
    counters) launched back to back on different streams, on ONE shared plan, together with ordinary kernels:
    items only wait for items handed out earlier to CTAs that are already running, so co-scheduling cannot
    deadlock; results stay bit-exact."""
    n = 1 << 15
    rng = np.random.default_rng(77)
    os.environ["CFFT_B200_FAST_VARIANT"] = "8"
    try:
        plan = C.unordered.Plan(n, C.unordered.Method.UserProvided(C.ordered.FftAlgo.Dif16, 256))
    finally:
        del os.environ["CFFT_B200_FAST_VARIANT"]
    small = C.unordered.Plan(2048, C.unordered.Method.UserProvided(C.ordered.FftAlgo.Dif16, 256))
    x = rand_c(rng, 96, n)
    ref = O.UnorderedPlan(n, O.DIF16, 256)
    want = ref.fwd(x, threads=8)
    back = ref.inv(want, threads=8)
    streams = [torch.cuda.Stream() for _ in range(4)]
    bufs = [torch.from_numpy(x.copy()).cuda() for _ in streams]
    filler = torch.view_as_complex(torch.rand(4096, 2048, 2, dtype=torch.float64, device="cuda")).contiguous()
    torch.cuda.synchronize()
    for _ in range(3):  # stress rounds: three persistent launches per stream, an ordinary kernel in between
        for s, b in zip(streams, bufs):
            with torch.cuda.stream(s):
                plan.fwd(b)
                if s is streams[0]:
                    small.fwd(filler)
                plan.inv(b)
                plan.fwd(b)
        torch.cuda.synchronize()
        for b in bufs:
            b.copy_(torch.from_numpy(x))
    # one checked round
    for s, b in zip(streams, bufs):
        with torch.cuda.stream(s):
            plan.fwd(b)
    torch.cuda.synchronize()
    for b in bufs:
        assert bits_equal(b.cpu().numpy(), want)
    for s, b in zip(streams, bufs):
        with torch.cuda.stream(s):
            plan.inv(b)
    torch.cuda.synchronize()
    for b in bufs:
        assert bits_equal(b.cpu().numpy(), back)


@pytest.mark.parametrize("n,measure", [(2048, False), (1 << 15, True), (1 << 18, True)])
def test_calls_can_be_captured_into_a_cuda_graph(C, torch, n, measure):
    """The device entry points are stream-capture safe (kernel launches, event fork / join over the
    library's auxiliary streams, nothing synchronous): a caller can record fwd / inv into a CUDA graph and
    replay it -- 0.05 ms of host time per call instead of ~1 ms for the multi-launch schedules
    (tools/graph_probe.py) -- with the same bits."""
    rng = np.random.default_rng(n)
    A = C.ordered.FftAlgo
    plan = C.unordered.Plan(n, C.unordered.Method.Measure() if measure else C.unordered.Method.UserProvided(A.Dif16, 256))
    batch = max(3, (96 << 20) // (16 * n)) if measure else 37  # large enough for the chunked schedule to fork
    x = rand_c(rng, batch, n)
    d = torch.from_numpy(x.copy()).cuda()
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        plan.fwd(d)  # warm-up outside the capture: creates the auxiliary streams
        d.copy_(torch.from_numpy(x))
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=s):
            plan.fwd(d)
        d.copy_(torch.from_numpy(x))  # capture does not execute; start from x again
        g.replay()
        torch.cuda.synchronize()
    want = dev_run(torch, plan.fwd, x)
    assert bits_equal(d.cpu().numpy(), want)
    sub = x[:2]
    assert bits_equal(want[:2], O.UnorderedPlan(n, O.DIF16, 256).fwd(sub, threads=8))


def test_random_plans_fuzz(C, torch):
    """Seeded fuzz over everything Plan::new accepts: random n, algo, base_n, batch and entry point
    (device / host-pageable), bit-exact against the oracle in both directions."""
    rng = np.random.default_rng(20261017)
    A = C.ordered.FftAlgo
    for trial in range(48):
        logn = int(rng.integers(0, 16))
        n = 1 << logn
        algo = int(rng.integers(0, 8))
        if rng.random() < 0.35 and logn <= 10:
            kind = "ordered"
            plan = C.ordered.Plan(n, C.ordered.Method.UserProvided(A(algo)))
            ref = O.OrderedPlan(n, algo) if n > 1 else None
        else:
            kind = "unordered"
            choices = [b for b in range(5, 11) if b <= logn] + [logn] if logn <= 10 else [b for b in range(5, 11)]
            base_n = 1 << int(rng.choice(choices))
            plan = C.unordered.Plan(n, C.unordered.Method.UserProvided(A(algo), base_n))
            ref = O.UnorderedPlan(n, algo, base_n) if n > 1 else None
        batch = int(rng.integers(1, max(2, min(40, (1 << 18) // n))))
        x = rand_c(rng, batch, n)
        if n == 1:
            assert bits_equal(dev_run(torch, plan.fwd, x), x)
            continue
        want = ref.fwd(x)
        if rng.random() < 0.5:
            y = dev_run(torch, plan.fwd, x)
            z = dev_run(torch, plan.inv, y)
        else:
            y = x.copy(); plan.fwd(y)
            z = y.copy(); plan.inv(z)
        assert bits_equal(y, want), (trial, kind, n, algo, batch)
        assert bits_equal(z, ref.inv(want)), (trial, kind, n, algo, batch)


def test_alignment_rules(C, torch):
    """Device buffers must be 16-byte aligned (128-bit accesses); host slices may be 8-byte aligned
    like a Rust &mut [Complex64] -- pageable or pinned."""
    rng = np.random.default_rng(31)
    n = 2048
    plan = C.unordered.Plan(n, C.unordered.Method.UserProvided(C.ordered.FftAlgo.Dif16, 256))
    ref = O.UnorderedPlan(n, O.DIF16, 256)
    x = rand_c(rng, 2, n)
    want = ref.fwd(x)
    raw = np.zeros(2 * (2 * n) + 1, np.float64)
    view = raw[1:].view(np.complex128).reshape(2, n)  # 8 mod 16
    if view.ctypes.data % 16 == 0:
        raw = np.zeros(2 * (2 * n) + 2, np.float64)
        view = raw[2:][: 4 * n].view(np.complex128).reshape(2, n)
    view[...] = x
    if view.ctypes.data % 16 == 8:
        plan.fwd(view)
        assert bits_equal(view, want)
    pinned = torch.zeros(2 * (2 * n) + 1, dtype=torch.float64).pin_memory()
    pv = pinned.numpy()[1:].view(np.complex128).reshape(2, n)
    pv[...] = x
    plan.fwd(pv)
    assert bits_equal(pv, want)
    dev = torch.zeros(2 * n * 2 + 2, dtype=torch.float64, device="cuda")
    st = C._native.lib.cfft_c64_fwd(plan._h, dev.data_ptr() + 8, 2, None)
    assert st == C._native.EINVAL


def test_strided_rows_bit_exact(C, torch):
    """cfft_c64_fwd_strided / _inv_strided (SURVEY.md 8b stride_elems): rows row_stride >= n apart, transformed where they are --
    x[:, j] of a [batch, k, n] record -- bit-identical to the packed call, neighbours untouched; fused-kernel plans take the
    stride in the kernel, other plans (generic register kernel, n = 2^14 passes, ordered) go through the packed workspace."""
    rng = np.random.default_rng(515)
    A = C.ordered.FftAlgo
    U = C.unordered
    plans = [(U.Plan(2048, U.Method.UserProvided(A.Dif16, 256)), O.UnorderedPlan(2048, O.DIF16, 256)),
             (U.Plan(256, U.Method.UserProvided(A.Dif16, 256)), O.UnorderedPlan(256, O.DIF16, 256)),
             (U.Plan(8192, U.Method.UserProvided(A.Dif16, 256)), O.UnorderedPlan(8192, O.DIF16, 256)),
             (U.Plan(2048, U.Method.UserProvided(A.Dif4, 32)), O.UnorderedPlan(2048, O.DIF4, 32)),
             (U.Plan(16384, U.Method.UserProvided(A.Dif16, 256)), O.UnorderedPlan(16384, O.DIF16, 256)),
             (C.ordered.Plan(512, C.ordered.Method.UserProvided(A.Dif8)), O.OrderedPlan(512, O.DIF8))]
    for plan, ref in plans:
        n = plan.fft_size()
        for batch, k, j in [(5, 3, 1), (1, 2, 1), (4, 1, 0)]:
            x = rand_c(rng, batch, k, n)
            d = torch.from_numpy(x.copy()).cuda()
            plan.fwd_strided(d[:, j])
            torch.cuda.synchronize()
            got = d.cpu().numpy()
            want = x.copy()
            want[:, j] = ref.fwd(np.ascontiguousarray(x[:, j]))
            assert bits_equal(got, want), (plan.kernel_name(), n, batch, k, j)
            plan.inv_strided(d[:, j])
            torch.cuda.synchronize()
            want[:, j] = ref.inv(np.ascontiguousarray(want[:, j]))
            assert bits_equal(d.cpu().numpy(), want), (plan.kernel_name(), n, batch, k, j)
    p = plans[0][0]
    d = torch.zeros((4, 2, 2048), dtype=torch.complex128, device="cuda")
    with pytest.raises(C.PanicError):
        p.fwd_strided(d[:, :, ::2][:, 0])  # inner stride 2
    with pytest.raises(C.PanicError):
        p.fwd_strided(d.reshape(8, 2048)[:, :1024])  # wrong row length
    # the C entry itself rejects overlapping rows
    import ctypes
    assert C._native.lib.cfft_c64_fwd_strided(p._h, d.data_ptr(), 1024, 2, None) == C._native.EINVAL


def test_host_call_sharded_over_replicas(C, torch):
    """cfft_c64_host_multi: ONE host call whose batch the library cuts over replicas of the plan (north_star item 5 inside the
    library).  Two replicas on cuda:0 always; cuda:0 + cuda:1 when the box has them.  Bit-identical to the oracle."""
    from concrete_fft_b200.sharding import MultiGpu

    rng = np.random.default_rng(808)
    A = C.ordered.FftAlgo
    plan = C.unordered.Plan(2048, C.unordered.Method.UserProvided(A.Dif16, 256))
    ref = O.UnorderedPlan(2048, O.DIF16, 256)
    device_sets = [[0, 0], [0, 0, 0]] + ([[0, 1]] if torch.cuda.device_count() >= 2 else [])
    for devices in device_sets:
        mg = MultiGpu(plan, devices)
        for batch in (1, 5, 1031):  # fewer rows than replicas, a ragged split, a multi-chunk split
            x = rand_c(rng, batch, 2048)
            h = x.copy()
            mg.fwd(h)
            want = ref.fwd(x, threads=8)
            assert bits_equal(h, want), (devices, batch)
            mg.inv(h)
            assert bits_equal(h, ref.inv(want, threads=8)), (devices, batch)
            h = x.copy()
            mg.fwd_inv(h)
            assert bits_equal(h, ref.inv(want, threads=8)), (devices, batch)
    with pytest.raises(C.PanicError):
        MultiGpu(plan, [0, 0]).fwd(np.zeros(100, np.complex128))
    other = C.unordered.Plan(2048, C.unordered.Method.UserProvided(A.Dif4, 32))
    import ctypes
    arr = (ctypes.c_void_p * 2)(plan._h.value if hasattr(plan._h, "value") else plan._h, other._h.value if hasattr(other._h, "value") else other._h)
    buf = np.zeros(2 * 2048, np.complex128)
    assert C._native.lib.cfft_c64_host_multi(arr, 2, 0, buf.ctypes.data, buf.size, 2) == C._native.EINVAL  # different transforms


@pytest.mark.parametrize("n", [256, 512, 1024, 2048, 4096])
def test_fwd_mul_inv_several_outputs_bit_exact(C, torch, n):
    """cfft_c64_fwd_mul_inv_multi (the GLWE external product: every forward transform feeds all outputs): bit-identical to
    cfft_c64_fwd_mul_inv once per output and to the oracle composition; one kernel for two outputs at n = 512 .. 2048, the
    output-by-output path elsewhere, b shared by the batch or per row, ragged tiles, the composed path."""
    rng = np.random.default_rng(7700 + n)
    A = C.ordered.FftAlgo
    plan = C.unordered.Plan(n, C.unordered.Method.UserProvided(A.Dif16, 256))
    ref = O.UnorderedPlan(n, O.DIF16, 256)
    assert C._native.lib.cfft_plan_has_fused_mul2_kernel(plan._h) == (1 if n in (512, 1024, 2048) else 0)
    for batch, k, n_out, per_row in [(5, 3, 2, False), (1, 1, 2, False), (700 if n <= 1024 else 310, 2, 2, False), (4, 2, 2, True), (3, 2, 3, False)]:
        a = rand_c(rng, batch, k, n) - (0.5 + 0.5j)
        b = (rand_c(rng, batch, k, n_out, n) if per_row else rand_c(rng, k, n_out, n)) - (0.5 + 0.5j)
        da, db = torch.from_numpy(a).cuda(), torch.from_numpy(b).cuda()
        got = plan.fwd_mul_inv_multi(da, db)
        torch.cuda.synchronize()
        for o in range(n_out):
            bo = torch.from_numpy(np.ascontiguousarray(b[:, :, o] if per_row else b[:, o])).cuda()
            single = plan.fwd_mul_inv(da, bo)
            torch.cuda.synchronize()
            assert bits_equal(got[:, o].cpu().numpy(), single.cpu().numpy()), (n, batch, k, n_out, per_row, o)
        if batch <= 5:  # the oracle composition, row by row
            for r in range(batch):
                for o in range(n_out):
                    acc = None
                    for i in range(k):
                        acc = O.c64_pointwise(ref.fwd(a[r, i]), (b[r, i, o] if per_row else b[i, o]), acc)
                    assert bits_equal(got[r, o].cpu().numpy(), ref.inv(acc)), (n, r, o)
    os.environ["CFFT_B200_FUSED_MUL_COMPOSED"] = "1"
    try:
        a = rand_c(rng, 3, 2, n)
        b = rand_c(rng, 2, 2, n)
        da, db = torch.from_numpy(a).cuda(), torch.from_numpy(b).cuda()
        composed = plan.fwd_mul_inv_multi(da, db)
        torch.cuda.synchronize()
    finally:
        del os.environ["CFFT_B200_FUSED_MUL_COMPOSED"]
    fused = plan.fwd_mul_inv_multi(da, db)
    torch.cuda.synchronize()
    assert bits_equal(composed.cpu().numpy(), fused.cpu().numpy())
    with pytest.raises(C.PanicError):
        plan.fwd_mul_inv_multi(da, torch.zeros((3, 2, n), dtype=torch.complex128, device="cuda"))  # k mismatch
    lib = C._native.lib
    assert lib.cfft_c64_fwd_mul_inv_multi(plan._h, da.data_ptr(), 2, db.data_ptr(), 0, 2, da.data_ptr(), 3, None) == C._native.EINVAL  # out overlaps a
