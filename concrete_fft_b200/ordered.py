"""Mirror of concrete_fft::ordered (src/ordered.rs): standard-order forward / inverse FFT.

    plan = Plan(n, Method.UserProvided(FftAlgo.Dif4))     # Plan::new, src/ordered.rs:242
    plan.fwd(buf)                                          # Plan::fwd, src/ordered.rs:342

`buf` is a numpy complex128 array (host memory, synchronous, like the Rust call) or a torch
complex128 CUDA tensor (device memory, stream ordered).  A buffer of batch * n elements is
treated as `batch` independent transforms.  Neither direction is normalised.
"""
import ctypes
import enum

from . import _native as N
from ._buffers import c64_view, current_stream_ptr


class FftAlgo(enum.IntEnum):
    """src/ordered.rs:28-45"""
    Dif2 = 0
    Dit2 = 1
    Dif4 = 2
    Dit4 = 3
    Dif8 = 4
    Dit8 = 5
    Dif16 = 6
    Dit16 = 7


class Method:
    """src/ordered.rs:50-59"""

    def __init__(self, kind, algo=None, duration=None):
        self.kind, self.algo, self.duration = kind, algo, duration

    @staticmethod
    def UserProvided(algo):
        return Method(N.METHOD_USER, FftAlgo(algo))

    @staticmethod
    def Measure(duration=None):
        """The duration is accepted for source compatibility; selection happens on the device."""
        return Method(N.METHOD_MEASURE, None, duration)

    def __eq__(self, other):
        return isinstance(other, Method) and (self.kind, self.algo) == (other.kind, other.algo)

    def __repr__(self):
        return "UserProvided(%s)" % self.algo.name if self.kind == N.METHOD_USER else "Measure(%r)" % (self.duration,)


class StackReq:
    """What Plan::fft_scratch returns in the reference (dyn_stack::StackReq): size + alignment."""

    def __init__(self, size_bytes, align_bytes):
        self.size_bytes, self.align_bytes = size_bytes, align_bytes

    def __repr__(self):
        return "StackReq(size_bytes=%d, align_bytes=%d)" % (self.size_bytes, self.align_bytes)


class _PlanBase:
    _h = None

    def __del__(self):
        try:
            if getattr(self, "_h", None):
                N.lib.cfft_plan_destroy(self._h)
                self._h = None
        except Exception:  # interpreter shutdown: the library handle may already be gone
            pass

    def fft_size(self):
        return int(N.lib.cfft_plan_fft_size(self._h))

    def kernel_name(self):
        return N.lib.cfft_plan_kernel_name(self._h).decode()

    def device(self):
        return int(N.lib.cfft_plan_device(self._h))

    def autotune(self, batch_hint=0):
        """Time this plan's kernel variants on the device and keep the fastest (never changes the
        plan's (base_algo, base_n) or any output bit).  Returns the report text."""
        N.check(N.lib.cfft_plan_autotune(self._h, batch_hint))
        return self.tuning_report()

    def tuning_report(self):
        import ctypes as _c

        buf = _c.create_string_buffer(4096)
        N.lib.cfft_plan_tuning_report(self._h, buf, 4096)
        return buf.value.decode()

    def _algo(self):
        a, b = ctypes.c_int(), ctypes.c_uint64()
        N.check(N.lib.cfft_plan_algo(self._h, ctypes.byref(a), ctypes.byref(b)))
        return FftAlgo(a.value), int(b.value)

    def fft_scratch(self):
        size, align = ctypes.c_uint64(), ctypes.c_uint64()
        N.check(N.lib.cfft_plan_scratch_req(self._h, ctypes.byref(size), ctypes.byref(align)))
        return StackReq(int(size.value), int(align.value))

    def _c64(self, buf, inverse):
        n = self.fft_size()
        kind, ptr, length, batch, dev = c64_view(buf, n)
        if length == 0 or length % n != 0:
            raise N.PanicError("assertion failed: buf.len() == fft_size (got %d, fft size %d)" % (length, n))
        if kind == "host":
            fn = N.lib.cfft_c64_inv_host if inverse else N.lib.cfft_c64_fwd_host
            N.check(fn(self._h, ptr, length, batch))
        else:
            if dev != self.device():
                raise ValueError("buffer is on cuda:%d but the plan lives on cuda:%d" % (dev, self.device()))
            fn = N.lib.cfft_c64_inv if inverse else N.lib.cfft_c64_fwd
            N.check(fn(self._h, ptr, batch, current_stream_ptr(dev)))

    def fwd(self, buf, stack=None):
        """In-place forward transform.  `stack` (the reference's PodStack scratch) is accepted and ignored."""
        self._c64(buf, False)

    def inv(self, buf, stack=None):
        """In-place unnormalised inverse transform."""
        self._c64(buf, True)

    def _c64_strided(self, view, inverse):
        """Rows of a 2-D CUDA tensor VIEW [batch, n] with unit inner stride and row stride >= n (cfft_c64_*_strided): the
        polynomials of a larger record, e.g. x[:, j] of a contiguous [batch, k, n] tensor, transformed where they are."""
        import torch

        n = self.fft_size()
        if not (isinstance(view, torch.Tensor) and view.is_cuda and view.dtype == torch.complex128):
            raise TypeError("buf must be a CUDA complex128 tensor")
        if view.dim() != 2 or view.shape[1] != n or view.stride(1) != 1 or (view.shape[0] > 1 and view.stride(0) < n):
            raise N.PanicError("assertion failed: buf has shape [batch, fft_size], unit inner stride, row stride >= fft_size")
        if view.device.index != self.device():
            raise ValueError("buffer is on cuda:%d but the plan lives on cuda:%d" % (view.device.index, self.device()))
        batch = int(view.shape[0])
        stride = int(view.stride(0)) if batch > 1 else n
        fn = N.lib.cfft_c64_inv_strided if inverse else N.lib.cfft_c64_fwd_strided
        N.check(fn(self._h, view.data_ptr(), stride, batch, current_stream_ptr(self.device())))

    def fwd_strided(self, view):
        """Plan::fwd on every row of a strided [batch, n] view, in place; elements between the rows are not touched."""
        self._c64_strided(view, False)

    def inv_strided(self, view):
        self._c64_strided(view, True)

    def fwd_inv_host(self, buf):
        """fwd then inv on the device between one upload and one download (bench `e2e` step)."""
        n = self.fft_size()
        kind, ptr, length, batch, _ = c64_view(buf, n)
        if kind != "host" or length % n:
            raise N.PanicError("fwd_inv_host needs a host buffer of batch * n elements")
        N.check(N.lib.cfft_c64_fwd_inv_host(self._h, ptr, length, batch))

    def fwd_mul_inv(self, a, b, out=None):
        """out[r] = inv(sum_k fwd(a[r, k]) * b[r, k])  (cfft_c64_fwd_mul_inv): a convolution / external-product step in
        one call, bit-identical to fwd + pointwise.mul_assign / mul_add_assign + inv.  CUDA complex128 tensors:
        `a` [batch, k, n] (or [batch, n] for k = 1), `b` [k, n] shared by every row or [batch, k, n], `out`
        [batch, n] (allocated when omitted; may be `a` itself when k = 1).  Returns out."""
        import torch

        n = self.fft_size()
        for name, t in (("a", a), ("b", b)):
            if not (isinstance(t, torch.Tensor) and t.is_cuda and t.dtype == torch.complex128 and t.is_contiguous()):
                raise TypeError("%s must be a contiguous CUDA complex128 tensor" % name)
        if a.dim() == 2:
            a = a.unsqueeze(1)
        if a.dim() != 3 or a.shape[2] != n or a.shape[1] < 1:
            raise N.PanicError("assertion failed: a has shape [batch, k, fft_size]")
        batch, k = int(a.shape[0]), int(a.shape[1])
        if tuple(b.shape) in ((k, n), (n,) if k == 1 else None):
            stride = 0
        elif tuple(b.shape) in ((batch, k, n), (batch, n) if k == 1 else None):
            stride = k * n
        else:
            raise N.PanicError("assertion failed: b has shape [k, fft_size] or [batch, k, fft_size]")
        if out is None:
            out = torch.empty((batch, n), dtype=torch.complex128, device=a.device)
        if not (isinstance(out, torch.Tensor) and out.is_cuda and out.dtype == torch.complex128 and out.is_contiguous()
                and out.numel() == batch * n):
            raise N.PanicError("assertion failed: out holds batch * fft_size elements")
        dev = a.device.index
        if dev != self.device() or b.device.index != dev or out.device.index != dev:
            raise ValueError("operands must live on the plan's device cuda:%d" % self.device())
        N.check(N.lib.cfft_c64_fwd_mul_inv(self._h, a.data_ptr(), k, b.data_ptr(), stride, out.data_ptr(), batch,
                                           current_stream_ptr(dev)))
        return out

    def fwd_mul_inv_multi(self, a, b, out=None):
        """out[r, o] = inv(sum_k fwd(a[r, k]) * b[r, k, o])  (cfft_c64_fwd_mul_inv_multi): the GLWE external product -- every
        forward transform feeds all n_out outputs.  `a` [batch, k, n]; `b` [k, n_out, n] shared by every row or [batch, k, n_out, n];
        returns `out` [batch, n_out, n].  Bit-identical to fwd_mul_inv once per output."""
        import torch

        n = self.fft_size()
        for name, t in (("a", a), ("b", b)):
            if not (isinstance(t, torch.Tensor) and t.is_cuda and t.dtype == torch.complex128 and t.is_contiguous()):
                raise TypeError("%s must be a contiguous CUDA complex128 tensor" % name)
        if a.dim() != 3 or a.shape[2] != n or a.shape[1] < 1:
            raise N.PanicError("assertion failed: a has shape [batch, k, fft_size]")
        batch, k = int(a.shape[0]), int(a.shape[1])
        if b.dim() == 3 and b.shape[0] == k and b.shape[2] == n:
            n_out, stride = int(b.shape[1]), 0
        elif b.dim() == 4 and tuple(b.shape[:2]) == (batch, k) and b.shape[3] == n:
            n_out, stride = int(b.shape[2]), k * int(b.shape[2]) * n
        else:
            raise N.PanicError("assertion failed: b has shape [k, n_out, fft_size] or [batch, k, n_out, fft_size]")
        if out is None:
            out = torch.empty((batch, n_out, n), dtype=torch.complex128, device=a.device)
        if not (isinstance(out, torch.Tensor) and out.is_cuda and out.dtype == torch.complex128 and out.is_contiguous()
                and tuple(out.shape) == (batch, n_out, n)):
            raise N.PanicError("assertion failed: out is a contiguous [batch, n_out, fft_size] complex128 CUDA tensor")
        dev = a.device.index
        if dev != self.device() or b.device.index != dev or out.device.index != dev:
            raise ValueError("all buffers must live on the plan's device cuda:%d" % self.device())
        N.check(N.lib.cfft_c64_fwd_mul_inv_multi(self._h, a.data_ptr(), k, b.data_ptr(), stride, n_out, out.data_ptr(), batch, current_stream_ptr(dev)))
        return out

    def fwd_mul_add(self, a, b, acc, accumulate=True):
        """acc[r] (Fourier domain) <- [acc[r] +] fwd(a[r]) * b[r]  (cfft_c64_fwd_mul_add), no inverse.  `a`: [batch, n] CUDA
        complex128, possibly a strided view a3[:, j] of a contiguous [batch, k, n] tensor; `b`: [n] / [batch, n] (or such a
        view of [batch, k, n]); `acc`: contiguous [batch, n]."""
        import torch

        n = self.fft_size()
        if not (isinstance(acc, torch.Tensor) and acc.is_cuda and acc.dtype == torch.complex128 and acc.is_contiguous()
                and acc.dim() == 2 and acc.shape[1] == n):
            raise N.PanicError("assertion failed: acc is a contiguous [batch, fft_size] complex128 CUDA tensor")
        batch = int(acc.shape[0])

        def rows(t, name, may_share):
            if not (isinstance(t, torch.Tensor) and t.is_cuda and t.dtype == torch.complex128):
                raise TypeError("%s must be a CUDA complex128 tensor" % name)
            if may_share and t.dim() == 1 and t.shape[0] == n and t.stride(0) == 1:
                return 0
            if t.dim() != 2 or tuple(t.shape) != (batch, n) or t.stride(1) != 1 or (batch > 1 and t.stride(0) < n):
                raise N.PanicError("assertion failed: %s has shape [batch, fft_size] with unit inner stride" % name)
            return int(t.stride(0)) if batch > 1 else n

        sa, sb = rows(a, "a", False), rows(b, "b", True)
        dev = acc.device.index
        if dev != self.device() or a.device.index != dev or b.device.index != dev:
            raise ValueError("operands must live on the plan's device cuda:%d" % self.device())
        N.check(N.lib.cfft_c64_fwd_mul_add(self._h, a.data_ptr(), sa, b.data_ptr(), sb, acc.data_ptr(), int(bool(accumulate)),
                                           batch, current_stream_ptr(dev)))
        return acc

    def has_fused_mul_kernel(self):
        return bool(N.lib.cfft_plan_has_fused_mul_kernel(self._h))

    # ---- integer polynomials <-> the Fourier domain (cfft_c64_poly_*, include/cfft_b200.h) --------------------------
    @staticmethod
    def _poly_flags(torus, accumulate=False):
        return (N.POLY_TORUS if torus else 0) | (N.POLY_ACCUMULATE if accumulate else 0)

    def _poly_tensor(self, t, name, inner):
        import torch

        if not (isinstance(t, torch.Tensor) and t.is_cuda and t.dtype == torch.int64 and t.is_contiguous()):
            raise TypeError("%s must be a contiguous CUDA int64 tensor" % name)
        if t.numel() == 0 or t.numel() % inner:
            raise N.PanicError("assertion failed: %s holds whole polynomials of %d coefficients" % (name, inner))
        if t.device.index != self.device():
            raise ValueError("%s is on cuda:%d but the plan lives on cuda:%d" % (name, t.device.index, self.device()))
        return t

    def _fourier_tensor(self, t, name):
        import torch

        if not (isinstance(t, torch.Tensor) and t.is_cuda and t.dtype == torch.complex128 and t.is_contiguous()):
            raise TypeError("%s must be a contiguous CUDA complex128 tensor" % name)
        if t.device.index != self.device():
            raise ValueError("%s is on cuda:%d but the plan lives on cuda:%d" % (name, t.device.index, self.device()))
        return t

    def fwd_poly(self, poly, out=None, torus=False):
        """fourier[r] = fwd(twist(fold(poly[r])))  (cfft_c64_poly_fwd).  `poly`: CUDA int64 [batch, 2 n] (torus: the u64
        torus elements viewed as int64); returns complex128 [batch, n] in this plan's order."""
        import torch

        n = self.fft_size()
        poly = self._poly_tensor(poly, "poly", 2 * n)
        batch = poly.numel() // (2 * n)
        if out is None:
            out = torch.empty((batch, n), dtype=torch.complex128, device=poly.device)
        out = self._fourier_tensor(out, "out")
        if out.numel() != batch * n:
            raise N.PanicError("assertion failed: out holds batch * fft_size elements")
        N.check(N.lib.cfft_c64_poly_fwd(self._h, poly.data_ptr(), out.data_ptr(), batch, self._poly_flags(torus), current_stream_ptr(self.device())))
        return out

    def inv_poly(self, fourier, out=None, torus=False, accumulate=False):
        """poly[r] (+)= round(untwist(inv(fourier[r])))  (cfft_c64_poly_inv); `fourier` is left untouched."""
        import torch

        n = self.fft_size()
        fourier = self._fourier_tensor(fourier, "fourier")
        if fourier.numel() == 0 or fourier.numel() % n:
            raise N.PanicError("assertion failed: fourier holds batch * fft_size elements")
        batch = fourier.numel() // n
        if out is None:
            if accumulate:
                raise N.PanicError("accumulate needs an output polynomial to add to")
            out = torch.empty((batch, 2 * n), dtype=torch.int64, device=fourier.device)
        out = self._poly_tensor(out, "out", 2 * n)
        if out.numel() != batch * 2 * n:
            raise N.PanicError("assertion failed: out holds batch polynomials")
        N.check(N.lib.cfft_c64_poly_inv(self._h, fourier.data_ptr(), out.data_ptr(), batch, self._poly_flags(torus, accumulate),
                                        current_stream_ptr(self.device())))
        return out

    def poly_mul(self, a, b, out=None, torus=False, accumulate=False):
        """out[r] (+)= round(untwist(inv(sum_k fwd(twist(fold(a[r, k]))) * b[r, k])))  (cfft_c64_poly_mul[_host]): a negacyclic
        product / external-product step with integer polynomials in and out.  `a`: int64 [batch, k, 2 n] (or [batch, 2 n]) --
        a CUDA tensor, or a numpy array in host memory (then `out` is a numpy array too and the call streams the batch
        through the GPU); `b`: CUDA complex128 Fourier-domain operand [k, n] (shared) or [batch, k, n]."""
        import numpy as np
        import torch

        n = self.fft_size()
        b = self._fourier_tensor(b, "b")
        host = isinstance(a, np.ndarray)
        if host:
            if a.dtype != np.int64 or not a.flags["C_CONTIGUOUS"]:
                raise TypeError("a must be a C-contiguous numpy int64 array")
        else:
            a = self._poly_tensor(a, "a", 2 * n)
        if a.ndim == 2:
            a = a.reshape(a.shape[0], 1, a.shape[1])
        if a.ndim != 3 or a.shape[2] != 2 * n or a.shape[1] < 1:
            raise N.PanicError("assertion failed: a has shape [batch, k, 2 * fft_size]")
        batch, k = int(a.shape[0]), int(a.shape[1])
        if tuple(b.shape) in ((k, n), (n,) if k == 1 else None):
            stride = 0
        elif tuple(b.shape) in ((batch, k, n), (batch, n) if k == 1 else None):
            stride = k * n
        else:
            raise N.PanicError("assertion failed: b has shape [k, fft_size] or [batch, k, fft_size]")
        flags = self._poly_flags(torus, accumulate)
        if host:
            if out is None:
                if accumulate:
                    raise N.PanicError("accumulate needs an output polynomial to add to")
                out = np.empty((batch, 2 * n), np.int64)
            if not (isinstance(out, np.ndarray) and out.dtype == np.int64 and out.flags["C_CONTIGUOUS"] and out.size == batch * 2 * n):
                raise N.PanicError("assertion failed: out is a C-contiguous numpy int64 array of batch polynomials")
            N.check(N.lib.cfft_c64_poly_mul_host(self._h, a.ctypes.data, k, b.data_ptr(), stride, out.ctypes.data, batch, flags))
            return out
        if out is None:
            if accumulate:
                raise N.PanicError("accumulate needs an output polynomial to add to")
            out = torch.empty((batch, 2 * n), dtype=torch.int64, device=a.device)
        out = self._poly_tensor(out, "out", 2 * n)
        if out.numel() != batch * 2 * n:
            raise N.PanicError("assertion failed: out holds batch polynomials")
        N.check(N.lib.cfft_c64_poly_mul(self._h, a.data_ptr(), k, b.data_ptr(), stride, out.data_ptr(), batch, flags,
                                        current_stream_ptr(self.device())))
        return out

    def has_fused_poly_kernel(self, k_terms=1):
        return bool(N.lib.cfft_plan_has_fused_poly_kernel(self._h, k_terms))

    def twist_tables(self):
        """(twist, untwist): e^{+i pi j / 2n} and conj / n, copied back from the device (tests)."""
        import numpy as np

        n = self.fft_size()
        out = np.empty(2 * n, np.complex128)
        N.check(N.lib.cfft_plan_copy_twist(self._h, out.ctypes.data, out.nbytes))
        return out[:n].copy(), out[n:].copy()

    def twiddles(self, inverse=False):
        """Device twiddle table copied back to the host (tests)."""
        import numpy as np

        n = self.fft_size()
        count = 2 * n if isinstance(self, Plan) else n + self._algo()[1]
        out = np.empty(count, np.complex128)
        N.check(N.lib.cfft_plan_copy_twiddles(self._h, int(inverse), out.ctypes.data, out.nbytes))
        return out


class Plan(_PlanBase):
    """ordered::Plan, src/ordered.rs:187-374."""

    def __init__(self, n, method, device=0, allow_large=False):
        if not isinstance(method, Method):
            raise TypeError("method must be an ordered.Method")
        h = ctypes.c_void_p()
        algo = int(method.algo) if method.algo is not None else 0
        N.check(N.lib.cfft_ordered_plan_create(ctypes.byref(h), device, n, method.kind, algo, int(allow_large)))
        self._h = h

    new = classmethod(lambda cls, n, method, **kw: cls(n, method, **kw))

    def algo(self):
        return self._algo()[0]

    def clone(self):
        h = ctypes.c_void_p()
        N.check(N.lib.cfft_plan_clone(self._h, ctypes.byref(h)))
        p = object.__new__(Plan)
        p._h = h
        return p

    def __repr__(self):
        return "Plan { algo: %s, fft_size: %d }" % (self.algo().name, self.fft_size())
