"""Element-wise c64 work in the Fourier domain -- what the reference leaves to the caller between `fwd` and
`inv` ("The only operations that are performed in the Fourier domain are elementwise", README.md:10-17 of
the reference).  Semantics: num_complex's `*` / `+` on Complex64, every operation individually rounded
(no FMA), so a GPU-resident pipeline gives the bits a Rust caller's `a * b` / `acc + a * b` gives.
Operands are CUDA complex128 tensors in ANY element order, as long as both came out of the same plan."""
from . import _native as N
from ._buffers import current_stream_ptr


def _dev_view(t, name):
    import torch

    if not (isinstance(t, torch.Tensor) and t.is_cuda and t.dtype == torch.complex128 and t.is_contiguous()):
        raise TypeError("%s must be a contiguous CUDA complex128 tensor" % name)
    return t.data_ptr(), t.numel(), t.device.index


def mul_assign(lhs, rhs):
    """lhs[i] *= rhs[i]  (cfft_c64_mul_assign)."""
    lp, ln, dev = _dev_view(lhs, "lhs")
    rp, rn, rdev = _dev_view(rhs, "rhs")
    if ln != rn or dev != rdev:
        raise N.PanicError("operands differ in length or device")
    N.check(N.lib.cfft_c64_mul_assign(dev, lp, rp, ln, current_stream_ptr(dev)))


def mul_add_assign(acc, a, b):
    """acc[i] += a[i] * b[i]  (cfft_c64_mul_add_assign): the accumulation step of an external product."""
    cp, cn, dev = _dev_view(acc, "acc")
    ap, an, adev = _dev_view(a, "a")
    bp, bn, bdev = _dev_view(b, "b")
    if not (cn == an == bn and dev == adev == bdev):
        raise N.PanicError("operands differ in length or device")
    N.check(N.lib.cfft_c64_mul_add_assign(dev, cp, ap, bp, cn, current_stream_ptr(dev)))
