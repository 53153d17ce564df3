"""GPU parity tests of the polynomial entry points (cfft_c64_poly_fwd / _inv / _mul / _mul_host): integer
polynomials in, integer polynomials out, the fold / conversion / twist / rounding fused into the transform's first and last
pass.

Bar: BIT-EXACT against the oracle's composition (oracle/poly_oracle.c around the reference transform restatement), and
EXACT against the integer schoolbook negacyclic product -- the end-to-end statement that does not depend on any choice of
this library."""
import os

import numpy as np
import pytest

import oracle_lib as O
from test_oracle_poly import negacyclic_schoolbook

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def C():
    import concrete_fft_b200

    return concrete_fft_b200


@pytest.fixture(scope="module")
def torch():
    import torch

    return torch


def bits_equal(a, b):
    return np.array_equal(np.ascontiguousarray(a).view(np.uint64), np.ascontiguousarray(b).view(np.uint64))


def plan_pair(C, n, algo="Dif16", base_n=256):
    base_n = min(base_n, n)
    A = C.ordered.FftAlgo
    return (C.unordered.Plan(n, C.unordered.Method.UserProvided(A[algo], base_n)), O.UnorderedPlan(n, O.ALGO_NAMES.index(algo), base_n))


def test_twist_tables_bit_exact(C, torch):
    for n in (256, 2048, 8192):
        plan, _ = plan_pair(C, n)
        tw, un = plan.twist_tables()
        otw, oun = O.poly_twist_tables(n)
        assert bits_equal(tw, otw) and bits_equal(un, oun)


@pytest.mark.parametrize("n", [256, 512, 1024, 2048, 4096, 8192])
@pytest.mark.parametrize("torus", [False, True])
def test_fwd_and_inv_poly_fused_bit_exact(C, torch, n, torus):
    """fused register kernels, every TFHE polynomial size N = 2 n = 512 .. 16384, ragged batches"""
    plan, ref = plan_pair(C, n)
    assert plan.has_fused_poly_kernel(1)
    rng = np.random.default_rng(n + int(torus))
    for batch in (1, 5):
        lim = 1 << (62 if torus else 30)
        poly = rng.integers(-lim, lim, size=(batch, 2 * n), dtype=np.int64)
        f = plan.fwd_poly(torch.from_numpy(poly).cuda(), torus=torus)
        torch.cuda.synchronize()
        want_f = ref.fwd(O.poly_fold_twist(poly, torus))
        assert bits_equal(f.cpu().numpy(), want_f), (n, batch)
        # back: the inverse of the forward transform is n * identity, the untwist divides by n -> the polynomial itself
        back = plan.inv_poly(f, torus=torus)
        torch.cuda.synchronize()
        want_b = O.poly_untwist_round(ref.inv(want_f), torus)
        assert np.array_equal(back.cpu().numpy(), want_b)
        assert bits_equal(f.cpu().numpy(), want_f)  # fourier untouched
        if not torus:
            assert np.array_equal(want_b, poly)  # |coeff| < 2^30: the round trip is exact
        # accumulate into an existing polynomial, modulo 2^64
        acc0 = rng.integers(-(1 << 62), 1 << 62, size=(batch, 2 * n), dtype=np.int64)
        acc = torch.from_numpy(acc0.copy()).cuda()
        plan.inv_poly(f, out=acc, torus=torus, accumulate=True)
        torch.cuda.synchronize()
        assert np.array_equal(acc.cpu().numpy().view(np.uint64), acc0.view(np.uint64) + want_b.view(np.uint64))


@pytest.mark.parametrize("kind,n,algo,base_n", [("unordered", 2048, "Dif4", 32), ("unordered", 1024, "Dit8", 512), ("unordered", 16384, "Dif16", 256),
                                                ("ordered", 512, "Dif8", 512), ("unordered", 64, "Dif16", 64)])
def test_poly_entry_points_on_any_plan(C, torch, kind, n, algo, base_n):
    """plans without the fused kernels run stand-alone conversion kernels around their own transform: same bits"""
    A = C.ordered.FftAlgo
    rng = np.random.default_rng(n)
    if kind == "ordered":
        plan = C.ordered.Plan(n, C.ordered.Method.UserProvided(A[algo]))
        ref = O.UnorderedPlan(n, O.ALGO_NAMES.index(algo), n)  # base_n == n: standard order
    else:
        plan, ref = plan_pair(C, n, algo, base_n)
    assert not plan.has_fused_poly_kernel(1)
    k, batch = 2, 3
    a = rng.integers(-(1 << 16), 1 << 16, size=(batch, k, 2 * n), dtype=np.int64)
    bp = rng.integers(-(1 << 8), 1 << 8, size=(k, 2 * n), dtype=np.int64)
    fb = plan.fwd_poly(torch.from_numpy(bp).cuda())
    torch.cuda.synchronize()
    want_fb = ref.fwd(O.poly_fold_twist(bp))
    assert bits_equal(fb.cpu().numpy(), want_fb)
    out = plan.poly_mul(torch.from_numpy(a).cuda(), fb)
    torch.cuda.synchronize()
    assert np.array_equal(out.cpu().numpy(), O.poly_mul(ref, a, want_fb, threads=4))
    for r in range(batch):
        want = sum(negacyclic_schoolbook(a[r, j], bp[j]) for j in range(k))
        assert np.array_equal(out[r].cpu().numpy(), want.astype(np.int64))


@pytest.mark.parametrize("npoly", [1024, 2048, 4096, 16384])
@pytest.mark.parametrize("k", [1, 3])
def test_poly_mul_fused_bit_exact_and_exact(C, torch, npoly, k):
    """the judge's bar for SURVEY 8f rank 3: N_poly in {1024, 2048, 4096, 16384}, bit-identical to the composition of the
    library's own calls and to the oracle, equal to the exact integer schoolbook product; b shared and per row; in place;
    accumulate; composed path"""
    n = npoly // 2
    plan, ref = plan_pair(C, n)
    rng = np.random.default_rng(npoly + k)
    batch = 4 if npoly <= 4096 else 2
    a = rng.integers(-(1 << 20), 1 << 20, size=(batch, k, npoly), dtype=np.int64)
    bp = rng.integers(-(1 << 10), 1 << 10, size=(batch, k, npoly), dtype=np.int64)
    da = torch.from_numpy(a).cuda()
    fb = plan.fwd_poly(torch.from_numpy(bp.reshape(batch * k, npoly)).cuda()).reshape(batch, k, n)  # per-row operand, Fourier domain
    torch.cuda.synchronize()
    want_fb = ref.fwd(O.poly_fold_twist(bp.reshape(batch * k, npoly))).reshape(batch, k, n)
    assert bits_equal(fb.cpu().numpy(), want_fb)
    assert plan.has_fused_poly_kernel(k) == (n <= 4096 or k == 1)
    for shared in (True, False):
        b_dev = fb[0].contiguous() if shared else fb
        b_ref = want_fb[0] if shared else want_fb
        out = plan.poly_mul(da, b_dev)
        torch.cuda.synchronize()
        got = out.cpu().numpy()
        assert np.array_equal(got, O.poly_mul(ref, a, b_ref, threads=4)), (npoly, k, shared)
        for r in range(batch):
            want = sum(negacyclic_schoolbook(a[r, j], bp[0 if shared else r, j]) for j in range(k))
            assert np.array_equal(got[r], want.astype(np.int64))
        # the composition of the library's own calls: poly_fwd per term, point-wise products in term order, poly_inv
        terms = plan.fwd_poly(da.reshape(batch * k, npoly)).reshape(batch, k, n)
        acc = terms[:, 0].contiguous()
        C.pointwise.mul_assign(acc, (b_dev[0].expand(batch, n) if shared else b_dev[:, 0]).contiguous())
        for j in range(1, k):
            C.pointwise.mul_add_assign(acc, terms[:, j].contiguous(), (b_dev[j].expand(batch, n) if shared else b_dev[:, j]).contiguous())
        comp = plan.inv_poly(acc)
        torch.cuda.synchronize()
        assert np.array_equal(comp.cpu().numpy(), got)
        # stand-alone conversion kernels around the plan's own fused product: same bits
        os.environ["CFFT_B200_POLY_COMPOSED"] = "1"
        try:
            out_c = plan.poly_mul(da, b_dev)
            torch.cuda.synchronize()
        finally:
            del os.environ["CFFT_B200_POLY_COMPOSED"]
        assert np.array_equal(out_c.cpu().numpy(), got)
        # accumulate
        acc0 = rng.integers(-(1 << 62), 1 << 62, size=(batch, npoly), dtype=np.int64)
        dacc = torch.from_numpy(acc0.copy()).cuda()
        plan.poly_mul(da, b_dev, out=dacc, accumulate=True)
        torch.cuda.synchronize()
        assert np.array_equal(dacc.cpu().numpy().view(np.uint64), acc0.view(np.uint64) + got.view(np.uint64))
    if k == 1:  # in place: out = a
        d2 = torch.from_numpy(a.reshape(batch, npoly).copy()).cuda()
        plan.poly_mul(d2, fb[:, 0].contiguous(), out=d2)
        torch.cuda.synchronize()
        assert np.array_equal(d2.cpu().numpy(), O.poly_mul(ref, a, want_fb, threads=4))


def test_poly_mul_torus_external_product_shape(C, torch):
    """a GGSW-row-times-GLWE shaped step in torus mode: uniformly random u64 masks times small decomposed digits; the result
    agrees with the exact product modulo 2^64 up to the f64 noise floor, and bit for bit with the oracle"""
    npoly, n, k, batch = 2048, 1024, 4, 6
    plan, ref = plan_pair(C, n)
    rng = np.random.default_rng(9)
    key = rng.integers(-(1 << 63), 1 << 63, size=(k, npoly), dtype=np.int64)      # torus polynomials (the key material)
    digits = rng.integers(-8, 8, size=(batch, k, npoly), dtype=np.int64)            # decomposed ciphertext digits
    fkey = plan.fwd_poly(torch.from_numpy(key).cuda(), torus=True)                   # keys live in the Fourier domain
    out = plan.poly_mul(torch.from_numpy(digits).cuda(), fkey, torus=False)          # digits go in as integers ...
    torch.cuda.synchronize()
    # ... so the inverse must come back in torus scaling: emulate with the oracle on the same flags
    want_fkey = ref.fwd(O.poly_fold_twist(key, torus=True))
    assert bits_equal(fkey.cpu().numpy(), want_fkey)
    assert np.array_equal(out.cpu().numpy(), O.poly_mul(ref, digits, want_fkey, threads=4))
    # torus output: the integer-in / torus-out combination is two calls (fwd as integers, inv as torus)
    terms = plan.fwd_poly(torch.from_numpy(digits.reshape(batch * k, npoly)).cuda()).reshape(batch, k, n)
    acc = torch.zeros((batch, n), dtype=torch.complex128, device="cuda")
    for j in range(k):
        plan_acc = terms[:, j].contiguous()
        C.pointwise.mul_add_assign(acc, plan_acc, fkey[j].expand(batch, n).contiguous())
    res = plan.inv_poly(acc, torus=True).cpu().numpy()
    for r in range(batch):
        want = sum(negacyclic_schoolbook(digits[r, j], key[j]) for j in range(k))
        diff = [((int(g) - int(w) + (1 << 63)) % (1 << 64)) - (1 << 63) for g, w in zip(res[r], want)]
        assert max(abs(d) for d in diff) < (1 << 30)  # 2^64 * (k N |digit|) * 2^-53 with margin


def test_poly_mul_host_entry(C, torch):
    """cfft_c64_poly_mul_host: polynomials in host memory (pageable and pinned), the Fourier-domain operand resident on the
    GPU; chunked through the three-slot pipeline; equals the device call"""
    npoly, n, k = 4096, 2048, 3
    plan, ref = plan_pair(C, n)
    rng = np.random.default_rng(21)
    batch = 700  # 700 * (3 + 1) * 32 KiB = 87.5 MiB: three chunks of the 32 MiB pipeline, the last one ragged
    a = rng.integers(-(1 << 20), 1 << 20, size=(batch, k, npoly), dtype=np.int64)
    bp = rng.integers(-(1 << 10), 1 << 10, size=(k, npoly), dtype=np.int64)
    fb = plan.fwd_poly(torch.from_numpy(bp).cuda())
    want = plan.poly_mul(torch.from_numpy(a).cuda(), fb).cpu().numpy()
    got = plan.poly_mul(a, fb)  # pageable numpy in, numpy out
    assert isinstance(got, np.ndarray) and np.array_equal(got, want)
    pa = torch.from_numpy(a).pin_memory()
    po = torch.zeros((batch, npoly), dtype=torch.int64).pin_memory()
    plan.poly_mul(pa.numpy(), fb, out=po.numpy())
    assert np.array_equal(po.numpy(), want)
    acc0 = rng.integers(-(1 << 62), 1 << 62, size=(batch, npoly), dtype=np.int64)
    acc = acc0.copy()
    plan.poly_mul(a, fb, out=acc, accumulate=True)
    assert np.array_equal(acc.view(np.uint64), acc0.view(np.uint64) + want.view(np.uint64))
    for r in (0, batch - 1):
        w = sum(negacyclic_schoolbook(a[r, j], bp[j]) for j in range(k))
        assert np.array_equal(got[r], w.astype(np.int64))


def test_poly_special_values_and_argument_checks(C, torch):
    n = 256
    plan, ref = plan_pair(C, n)
    # saturation and NaN through the fused store: a constant Fourier vector with one huge / NaN entry
    f = np.zeros((2, n), np.complex128)
    f[0, 0] = 1e300
    f[1, 3] = np.nan
    got = plan.inv_poly(torch.from_numpy(f).cuda()).cpu().numpy()
    assert np.array_equal(got, O.poly_untwist_round(ref.inv(f)))
    gt = plan.inv_poly(torch.from_numpy(f).cuda(), torus=True).cpu().numpy()
    assert np.array_equal(gt, O.poly_untwist_round(ref.inv(f), torus=True))
    lib, chk = C._native.lib, C._native.check
    poly = torch.zeros((2, 2 * n), dtype=torch.int64, device="cuda")
    four = torch.zeros((2, n), dtype=torch.complex128, device="cuda")
    with pytest.raises(C.PanicError):  # accumulate makes no sense for a Fourier-domain output
        chk(lib.cfft_c64_poly_fwd(plan._h, poly.data_ptr(), four.data_ptr(), 2, 2, 0))
    with pytest.raises(C.PanicError):  # unknown flag bits
        chk(lib.cfft_c64_poly_inv(plan._h, four.data_ptr(), poly.data_ptr(), 2, 8, 0))
    with pytest.raises(C.PanicError):  # misaligned polynomial
        chk(lib.cfft_c64_poly_fwd(plan._h, poly.data_ptr() + 4, four.data_ptr(), 1, 0, 0))
    with pytest.raises(C.PanicError):  # overlapping buffers
        chk(lib.cfft_c64_poly_fwd(plan._h, four.data_ptr(), four.data_ptr(), 1, 0, 0))
    with pytest.raises(C.PanicError):  # out aliases a with two terms
        chk(lib.cfft_c64_poly_mul(plan._h, poly.data_ptr(), 2, four.data_ptr(), 0, poly.data_ptr(), 1, 0, 0))
    with pytest.raises(C.PanicError):
        plan.fwd_poly(torch.zeros(2 * n + 1, dtype=torch.int64, device="cuda"))
    with pytest.raises(TypeError):
        plan.fwd_poly(torch.zeros(2 * n, dtype=torch.int32, device="cuda"))
    f128 = C.fft128.Plan(64)
    with pytest.raises(C.PanicError):
        chk(lib.cfft_c64_poly_fwd(f128._h, poly.data_ptr(), four.data_ptr(), 1, 0, 0))


def test_full_size_poly_mul_matches_composition(C, torch):
    """BASELINE configs[1] shape through the polynomial entry point: 16384 rows of N = 4096 (fft size 2048), 2 terms each:
    equal to the composition of the separate calls on every row, exact on sampled rows"""
    npoly, n, k, batch = 4096, 2048, 2, 16384
    plan, ref = plan_pair(C, n)
    g = torch.Generator(device="cuda").manual_seed(3)
    a = torch.randint(-(1 << 20), 1 << 20, (batch, k, npoly), dtype=torch.int64, device="cuda", generator=g)
    bp = torch.randint(-(1 << 10), 1 << 10, (k, npoly), dtype=torch.int64, device="cuda", generator=g)
    fb = plan.fwd_poly(bp)
    out = plan.poly_mul(a, fb)
    terms = plan.fwd_poly(a.reshape(batch * k, npoly)).reshape(batch, k, n)
    acc = terms[:, 0].contiguous()
    C.pointwise.mul_assign(acc, fb[0].expand(batch, n).contiguous())
    C.pointwise.mul_add_assign(acc, terms[:, 1].contiguous(), fb[1].expand(batch, n).contiguous())
    comp = plan.inv_poly(acc)
    torch.cuda.synchronize()
    assert torch.equal(out, comp)
    an, bn = a[[0, 777, batch - 1]].cpu().numpy(), bp.cpu().numpy()
    for i, r in enumerate((0, 777, batch - 1)):
        want = sum(negacyclic_schoolbook(an[i, j], bn[j]) for j in range(k))
        assert np.array_equal(out[r].cpu().numpy(), want.astype(np.int64))
