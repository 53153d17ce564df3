"""Mirror of concrete_fft::fft128 (src/fft128/mod.rs): negacyclic transform on double-double
complex data held in four planar f64 arrays (re hi, re lo, im hi, im lo)."""
import ctypes

from . import _native as N
from ._buffers import current_stream_ptr, f64_view


class f128:
    """src/fft128/mod.rs:3-7: value = hi + lo."""

    __slots__ = ("hi", "lo")

    def __init__(self, hi=0.0, lo=0.0):
        self.hi, self.lo = float(hi), float(lo)

    def __repr__(self):
        return "f128(%r, %r)" % (self.hi, self.lo)


class Plan:
    """fft128::Plan, src/fft128/mod.rs:1832-1961."""

    def __init__(self, n, device=0):
        h = ctypes.c_void_p()
        N.check(N.lib.cfft_f128_plan_create(ctypes.byref(h), device, n))
        self._h = h

    new = classmethod(lambda cls, n, **kw: cls(n, **kw))

    def __del__(self):
        try:
            if getattr(self, "_h", None):
                N.lib.cfft_plan_destroy(self._h)
                self._h = None
        except Exception:  # interpreter shutdown: the library handle may already be gone
            pass

    def fft_size(self):
        return int(N.lib.cfft_plan_fft_size(self._h))

    def device(self):
        return int(N.lib.cfft_plan_device(self._h))

    def kernel_name(self):
        return N.lib.cfft_plan_kernel_name(self._h).decode()

    def __repr__(self):
        return "Plan { fft_size: %d }" % self.fft_size()

    def autotune(self, batch_hint=0):
        N.check(N.lib.cfft_plan_autotune(self._h, batch_hint))
        buf = ctypes.create_string_buffer(4096)
        N.lib.cfft_plan_tuning_report(self._h, buf, 4096)
        return buf.value.decode()

    def _run(self, planes, inverse):
        n = self.fft_size()
        views = [f64_view(p) for p in planes]
        kinds = {v[0] for v in views}
        if len(kinds) != 1:
            raise TypeError("all four planes must live in the same memory space")
        for v in views:  # four assert_eq!, src/fft128/mod.rs:1912-1915
            if v[2] == 0 or v[2] % n or v[2] != views[0][2]:
                raise N.PanicError("assertion failed: buf.len() == fft_size")
        length, batch = views[0][2], views[0][2] // n
        ptrs = [v[1] for v in views]
        if "host" in kinds:
            fn = N.lib.cfft_f128_inv_host if inverse else N.lib.cfft_f128_fwd_host
            N.check(fn(self._h, *ptrs, length, batch))
        else:
            dev = views[0][3]
            if any(v[3] != self.device() for v in views):
                raise ValueError("planes must live on the plan's device cuda:%d" % self.device())
            fn = N.lib.cfft_f128_inv if inverse else N.lib.cfft_f128_fwd
            N.check(fn(self._h, *ptrs, batch, current_stream_ptr(dev)))

    def _run_strided(self, planes, inverse):
        """Four CUDA float64 VIEWS [batch, n] with unit inner stride and one common row stride >= n (cfft_f128_*_strided)."""
        import torch

        n = self.fft_size()
        for p in planes:
            if not (isinstance(p, torch.Tensor) and p.is_cuda and p.dtype == torch.float64):
                raise TypeError("planes must be CUDA float64 tensors")
            if p.dim() != 2 or p.shape != planes[0].shape or p.shape[1] != n or p.stride(1) != 1 or p.stride(0) != planes[0].stride(0) \
                    or (p.shape[0] > 1 and p.stride(0) < n):
                raise N.PanicError("assertion failed: planes have shape [batch, fft_size], unit inner stride, one row stride >= fft_size")
            if p.device.index != self.device():
                raise ValueError("planes must live on the plan's device cuda:%d" % self.device())
        batch = int(planes[0].shape[0])
        stride = int(planes[0].stride(0)) if batch > 1 else n
        fn = N.lib.cfft_f128_inv_strided if inverse else N.lib.cfft_f128_fwd_strided
        N.check(fn(self._h, *[p.data_ptr() for p in planes], stride, batch, current_stream_ptr(self.device())))

    def fwd_strided(self, buf_re0, buf_re1, buf_im0, buf_im1):
        self._run_strided((buf_re0, buf_re1, buf_im0, buf_im1), False)

    def inv_strided(self, buf_re0, buf_re1, buf_im0, buf_im1):
        self._run_strided((buf_re0, buf_re1, buf_im0, buf_im1), True)

    def fwd(self, buf_re0, buf_re1, buf_im0, buf_im1):
        """src/fft128/mod.rs:1905-1928: standard order in, bit-reversed order out."""
        self._run((buf_re0, buf_re1, buf_im0, buf_im1), False)

    def inv(self, buf_re0, buf_re1, buf_im0, buf_im1):
        """src/fft128/mod.rs:1938-1960: bit-reversed order in, standard order out, unnormalised."""
        self._run((buf_re0, buf_re1, buf_im0, buf_im1), True)

    def fwd_inv_host(self, buf_re0, buf_re1, buf_im0, buf_im1):
        """fwd then inv on the device between one upload and one download (bench `e2e` step)."""
        n = self.fft_size()
        views = [f64_view(p) for p in (buf_re0, buf_re1, buf_im0, buf_im1)]
        if any(v[0] != "host" for v in views) or any(v[2] != views[0][2] or v[2] % n for v in views):
            raise N.PanicError("fwd_inv_host needs four host planes of batch * n doubles")
        N.check(N.lib.cfft_f128_fwd_inv_host(self._h, *[v[1] for v in views], views[0][2], views[0][2] // n))

    def fwd_mul_inv(self, lhs, rhs, factor):
        """lhs <- inv((fwd(lhs) * rhs) * factor)  (cfft_f128_fwd_mul_inv): the negacyclic product of the reference's
        tests (src/fft128/mod.rs:2018-2053) in one call.  `lhs`: four CUDA float64 planes of batch * n, in place; `rhs`:
        four Fourier-domain planes of n (shared by every row) or batch * n doubles."""
        n = self.fft_size()
        lv = [f64_view(t) for t in lhs]
        rv = [f64_view(t) for t in rhs]
        if any(v[0] != "device" for v in lv + rv) or len({v[2] for v in lv}) != 1 or len({v[2] for v in rv}) != 1:
            raise TypeError("fwd_mul_inv takes two sets of four CUDA float64 planes of equal length")
        length = lv[0][2]
        if length == 0 or length % n:
            raise N.PanicError("assertion failed: buf.len() == fft_size")
        if rv[0][2] not in (n, length):
            raise N.PanicError("assertion failed: rhs holds fft_size or batch * fft_size points")
        stride = 0 if rv[0][2] == n else n
        dev = lv[0][3]
        if any(v[3] != self.device() for v in lv + rv):
            raise ValueError("planes must live on the plan's device cuda:%d" % self.device())
        N.check(N.lib.cfft_f128_fwd_mul_inv(self._h, *[v[1] for v in lv], *[v[1] for v in rv], stride, float(factor),
                                            length // n, current_stream_ptr(dev)))

    def has_fused_mul_kernel(self):
        return bool(N.lib.cfft_plan_has_fused_mul_kernel(self._h))

    def twiddles(self):
        import numpy as np

        out = []
        for w in range(4):
            a = np.empty(self.fft_size(), np.float64)
            N.check(N.lib.cfft_plan_copy_twiddles(self._h, w, a.ctypes.data, a.nbytes))
            out.append(a)
        return out


# ---- the scalar f128 operators on device arrays (src/fft128/f128_ops.rs) ---------------------------
_OPS = {"add": 0, "sub": 1, "mul": 2, "div": 3, "add_estimate": 4, "sub_estimate": 5, "div_estimate": 6,
        # mixed-operand forms (f128_ops.rs:279-455): the f64 operand is passed as its value plane, lo = None
        "add_f128_f64": 7, "sub_f128_f64": 8, "sub_f64_f128": 9, "mul_f128_f64": 10, "div_f128_f64": 11, "div_f64_f128": 12,
        "add_f64_f64": 13, "sub_f64_f64": 14, "mul_f64_f64": 15, "div_f64_f64": 16}
_A_F64 = {9, 12, 13, 14, 15, 16}
_B_F64 = {7, 8, 10, 11, 13, 14, 15, 16}
_UNARY = {"sqr": 0, "abs": 1, "neg": 2, "sincospi": 3, "is_nan": 4}


def _plane(t, name):
    v = f64_view(t)
    if v[0] != "device":
        raise TypeError("%s must be a CUDA float64 tensor" % name)
    return v


def f128_op(op, a_hi, a_lo, b_hi, b_lo):
    """Element-wise f128 operator on CUDA float64 tensors; returns (hi, lo), bit-exact with the reference's scalar functions.
    `op`: add, sub, mul, div (f128::add_f128_f128 ... div_f128_f128), add_estimate, sub_estimate, div_estimate, or a
    mixed-operand form add_f128_f64, sub_f128_f64, sub_f64_f128, mul_f128_f64, div_f128_f64, div_f64_f128, add_f64_f64,
    sub_f64_f64, mul_f64_f64, div_f64_f64 -- pass None as the lo plane of an f64 operand (add_f64_f128 / mul_f64_f128 are the
    f128_f64 forms with the operands swapped, f128_ops.rs:294-298, 388-391)."""
    import torch

    code = _OPS[op]
    planes = [("a_hi", a_hi, False), ("a_lo", a_lo, code in _A_F64), ("b_hi", b_hi, False), ("b_lo", b_lo, code in _B_F64)]
    ptrs, length, dev = [], None, None
    for name, t, optional in planes:
        if t is None and optional:
            ptrs.append(None)
            continue
        v = _plane(t, name)
        if length is None:
            length, dev = v[2], v[3]
        if v[2] != length or v[3] != dev:
            raise TypeError("f128_op operands must have equal length and live on one device")
        ptrs.append(v[1])
    out_hi, out_lo = torch.empty_like(a_hi), torch.empty_like(a_hi)
    N.check(N.lib.cfft_f128_binary_op(dev, code, *ptrs, out_hi.data_ptr(), out_lo.data_ptr(), length, current_stream_ptr(dev)))
    return out_hi, out_lo


def f128_unary(op, a_hi, a_lo):
    """sqr, abs, neg, is_nan -> (hi, lo); sincospi -> ((sin_hi, sin_lo), (cos_hi, cos_lo)) for inputs in [-1, 1] (raises
    PanicError otherwise, like the reference's panic)  (f128_ops.rs:404-409, 494-575)."""
    import torch

    va, vb = _plane(a_hi, "a_hi"), _plane(a_lo, "a_lo")
    if va[2] != vb[2] or va[3] != vb[3]:
        raise TypeError("f128_unary operands must have equal length and live on one device")
    dev = va[3]
    code = _UNARY[op]
    out_hi, out_lo = torch.empty_like(a_hi), torch.empty_like(a_hi)
    o2h = torch.empty_like(a_hi) if code == 3 else None
    o2l = torch.empty_like(a_hi) if code == 3 else None
    N.check(N.lib.cfft_f128_unary_op(dev, code, va[1], vb[1], out_hi.data_ptr(), out_lo.data_ptr(), o2h.data_ptr() if code == 3 else None,
                                     o2l.data_ptr() if code == 3 else None, va[2], current_stream_ptr(dev)))
    if code == 3:
        return (out_hi, out_lo), (o2h, o2l)
    return out_hi, out_lo


def f128_compare(a_hi, a_lo, b_hi, b_lo=None):
    """PartialOrd of f128 element-wise: int8 tensor of -1 (Less), 0 (Equal), 1 (Greater), 2 (None / unordered); b_lo = None
    compares with the f64 values b_hi (f128_ops.rs:240-274)."""
    import torch

    vs = [_plane(a_hi, "a_hi"), _plane(a_lo, "a_lo"), _plane(b_hi, "b_hi")] + ([_plane(b_lo, "b_lo")] if b_lo is not None else [])
    if len({v[2] for v in vs}) != 1 or len({v[3] for v in vs}) != 1:
        raise TypeError("f128_compare operands must have equal length and live on one device")
    dev = vs[0][3]
    out = torch.empty(a_hi.shape, dtype=torch.int8, device=a_hi.device)
    N.check(N.lib.cfft_f128_compare(dev, vs[0][1], vs[1][1], vs[2][1], vs[3][1] if b_lo is not None else None, out.data_ptr(), vs[0][2],
                                    current_stream_ptr(dev)))
    return out


def cplx_mul_scale(lhs, rhs, factor):
    """lhs <- (lhs * rhs) * factor point-wise on planar (re0, re1, im0, im1) CUDA tensors: the step
    between fwd and inv of a negacyclic product (src/fft128/mod.rs:2033-2047)."""
    lv = [f64_view(t) for t in lhs]
    rv = [f64_view(t) for t in rhs]
    if any(v[0] != "device" for v in lv + rv) or len({v[2] for v in lv + rv}) != 1:
        raise TypeError("cplx_mul_scale takes eight CUDA float64 tensors of equal length")
    dev = lv[0][3]
    N.check(N.lib.cfft_f128_cplx_mul_scale(dev, *[v[1] for v in lv], *[v[1] for v in rv], float(factor), lv[0][2],
                                           current_stream_ptr(dev)))
