"""Mirror of concrete_fft::unordered (src/unordered.rs): forward output / inverse input live in
a plan-specific permuted order, reproduced index for index (src/unordered.rs:1046-1051)."""
import ctypes
import struct

import numpy as np

from . import _native as N
from ._buffers import c64_view, current_stream_ptr, is_torch
from .ordered import FftAlgo, StackReq, _PlanBase  # noqa: F401  (re-exported like the Rust module)


class Method:
    """src/unordered.rs:526-537"""

    def __init__(self, kind, base_algo=None, base_n=0, duration=None):
        self.kind, self.base_algo, self.base_n, self.duration = kind, base_algo, base_n, duration

    @staticmethod
    def UserProvided(base_algo, base_n):
        return Method(N.METHOD_USER, FftAlgo(base_algo), int(base_n))

    @staticmethod
    def Measure(duration=None):
        return Method(N.METHOD_MEASURE, None, 0, duration)

    def __repr__(self):
        if self.kind == N.METHOD_USER:
            return "UserProvided { base_algo: %s, base_n: %d }" % (self.base_algo.name, self.base_n)
        return "Measure(%r)" % (self.duration,)


class Plan(_PlanBase):
    """unordered::Plan, src/unordered.rs:496-1037."""

    def __init__(self, n, method, device=0):
        if not isinstance(method, Method):
            raise TypeError("method must be an unordered.Method")
        h = ctypes.c_void_p()
        algo = int(method.base_algo) if method.base_algo is not None else 0
        N.check(N.lib.cfft_unordered_plan_create(ctypes.byref(h), device, n, method.kind, algo, method.base_n))
        self._h = h

    new = classmethod(lambda cls, n, method, **kw: cls(n, method, **kw))

    def algo(self):
        """(base_algo, base_n), src/unordered.rs:783-785"""
        return self._algo()

    def clone(self):
        h = ctypes.c_void_p()
        N.check(N.lib.cfft_plan_clone(self._h, ctypes.byref(h)))
        p = object.__new__(Plan)
        p._h = h
        return p

    def __repr__(self):
        a, b = self.algo()
        return "Plan { base_algo: %s, base_size: %d, fft_size: %d }" % (a.name, b, self.fft_size())

    def fwd_monomial(self, degree, buf):
        """src/unordered.rs:844-900"""
        n = self.fft_size()
        kind, ptr, length, _, dev = c64_view(buf, n)
        if length != n:
            raise N.PanicError("assertion failed: fft_size == buf.len()")
        if not 0 <= degree < n:
            raise N.PanicError("assertion failed: degree < fft_size")
        if kind == "host":
            N.check(N.lib.cfft_unordered_fwd_monomial_host(self._h, degree, ptr, length))
        else:
            self._check_device(dev)
            N.check(N.lib.cfft_unordered_fwd_monomial(self._h, degree, ptr, current_stream_ptr(dev)))

    def _check_device(self, dev):
        if dev != self.device():
            raise ValueError("buffer is on cuda:%d but the plan lives on cuda:%d" % (dev, self.device()))

    def permutation(self):
        """perm[i] = index in the plan's buffer of Fourier coefficient i (bit_rev_twice)."""
        out = np.empty(self.fft_size(), np.uint64)
        N.check(N.lib.cfft_unordered_permutation(self._h, out.ctypes.data))
        return out

    # ---- serde mapping, src/unordered.rs:942-1036 -------------------------------------------
    def serialize_fourier_buffer(self, buf):
        """Standard-order copy of a permuted Fourier-domain buffer (what the reference hands to
        the serde serializer element by element).  numpy in -> numpy out; CUDA tensor in ->
        CUDA tensor out (batched: any multiple of n)."""
        n = self.fft_size()
        kind, ptr, length, batch, dev = c64_view(buf, n)
        if length == 0 or length % n:
            raise N.PanicError("assertion failed: n == buf.len()")
        if kind == "host":
            out = np.empty_like(buf)
            src, dst = buf.reshape(-1, n), out.reshape(-1, n)
            for r in range(batch):
                N.check(N.lib.cfft_unordered_to_standard_host(self._h, src[r].ctypes.data, dst[r].ctypes.data))
            return out
        import torch

        self._check_device(dev)
        out = torch.empty_like(buf)
        N.check(N.lib.cfft_unordered_to_standard(self._h, ptr, out.data_ptr(), batch, current_stream_ptr(dev)))
        return out

    def deserialize_fourier_buffer(self, seq, buf):
        """Scatter a standard-order sequence into `buf` in the plan's order; raises InvalidLength
        when the sequence does not hold exactly n elements (src/unordered.rs:1027-1031)."""
        n = self.fft_size()
        kind, ptr, length, batch, dev = c64_view(buf, n)
        if kind == "device":
            self._check_device(dev)
            if not is_torch(seq):
                raise TypeError("a CUDA buffer needs its sequence as a CUDA complex128 tensor on the same device")
            # the sequence is validated exactly like the buffer: a host tensor, another dtype or another GPU would hand the
            # kernel a pointer it cannot read (a sticky illegal-address error that poisons the CUDA context)
            skind, sptr, slen, _, sdev = c64_view(seq.contiguous(), n)
            if sdev != dev:
                raise ValueError("sequence is on cuda:%d but the buffer is on cuda:%d" % (sdev, dev))
            if slen != length or length % n:
                raise N.InvalidLength("invalid length %d, expected a sequence of %d 64-bit complex numbers" % (slen, n))
            if sptr == ptr:
                raise N.PanicError("sequence and buffer must not be the same memory")
            N.check(N.lib.cfft_unordered_from_standard(self._h, sptr, ptr, batch, current_stream_ptr(dev)))
            return
        if length != n:
            raise N.PanicError("assertion failed: n == buf.len()")
        seq = np.ascontiguousarray(seq, dtype=np.complex128)
        st = N.lib.cfft_unordered_from_standard_host(self._h, seq.ctypes.data, seq.size, ptr)
        if st == N.ELENGTH:
            raise N.InvalidLength("invalid length %d, expected a sequence of %d 64-bit complex numbers" % (seq.size, n))
        N.check(st)

    def serialize_bincode(self, buf):
        """bincode 1.3 framing of serialize_fourier_buffer (u64 LE length, then re, im f64 LE per
        element), the format the reference's serde test round-trips (src/unordered.rs:9447-9454)."""
        if is_torch(buf):  # CUDA tensor: gather on the device, frame the host copy
            std = self.serialize_fourier_buffer(buf.contiguous()).cpu().numpy()
        else:
            std = self.serialize_fourier_buffer(np.ascontiguousarray(buf))
        return struct.pack("<Q", std.size) + std.astype("<c16").tobytes()

    def deserialize_bincode(self, blob, buf):
        (count,) = struct.unpack_from("<Q", blob, 0)
        avail = (len(blob) - 8) // 16
        seq = np.frombuffer(blob, dtype="<c16", count=min(count, avail), offset=8)
        if count != seq.size:
            raise N.InvalidLength("truncated bincode sequence")
        self.deserialize_fourier_buffer(seq, buf)
