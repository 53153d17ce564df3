"""Batch sharding across GPUs: polynomials are independent (the reference API is one polynomial
per call, src/unordered.rs:826), so ranks take contiguous row ranges and never communicate on
the data path.  torch.distributed is used only for the barrier and the max-over-ranks timing."""


def shard_rows(batch, world_size, rank):
    """Contiguous split of `batch` rows: returns (first_row, row_count) for `rank`.
    Ranks differ by at most one row; every row is owned by exactly one rank."""
    if world_size < 1 or not 0 <= rank < world_size:
        raise ValueError("bad rank / world size")
    base, extra = divmod(batch, world_size)
    first = rank * base + min(rank, extra)
    return first, base + (1 if rank < extra else 0)


def max_over_ranks(value, dist=None, device=None):
    """Whole-job time = max over ranks (all-reduce MAX); identity without a process group."""
    if dist is None or not dist.is_initialized():
        return float(value)
    import torch

    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def job_throughput(units_per_rank, seconds_per_rank, dist=None, device=None):
    """units all ranks processed / max-over-ranks time (bench.py's `value`)."""
    if dist is None or not dist.is_initialized():
        return units_per_rank / seconds_per_rank
    import torch

    u = torch.tensor([float(units_per_rank)], dtype=torch.float64, device=device)
    dist.all_reduce(u, op=dist.ReduceOp.SUM)
    return float(u.item()) / max_over_ranks(seconds_per_rank, dist, device)


class MultiGpu:
    """One plan, several GPUs, ONE process: host-memory calls whose batch the library itself cuts into contiguous row ranges, one
    per replica of the plan (cfft_plan_clone_to_device + cfft_c64_host_multi / cfft_f128_host_multi).  `devices`: CUDA
    device indices (default: all visible; an index may repeat).  Same bits as the plan's own host entry points."""

    def __init__(self, plan, devices=None):
        import ctypes

        from . import _native as N

        if devices is None:
            import torch

            devices = list(range(torch.cuda.device_count()))
        if not devices:
            raise ValueError("no devices")
        self._N, self._plan, self._replicas = N, plan, []
        for d in devices:
            h = ctypes.c_void_p()
            N.check(N.lib.cfft_plan_clone_to_device(plan._h, int(d), ctypes.byref(h)))
            self._replicas.append(h)
        self._arr = (ctypes.c_void_p * len(self._replicas))(*[h.value for h in self._replicas])
        self.devices = [int(d) for d in devices]

    def __del__(self):
        lib = getattr(getattr(self, "_N", None), "lib", None)  # None while the interpreter shuts down
        if lib is not None:
            for h in getattr(self, "_replicas", []):
                lib.cfft_plan_destroy(h)
        self._replicas = []

    def _is_f128(self):
        return self._N.lib.cfft_plan_kind(self._plan._h) == 2

    def _c64(self, buf, op):
        import numpy as np

        n = self._plan.fft_size()
        if not isinstance(buf, np.ndarray) or buf.dtype != np.complex128 or not buf.flags["C_CONTIGUOUS"] or not buf.flags["WRITEABLE"]:
            raise TypeError("buf must be a writeable C-contiguous numpy complex128 array (host memory)")
        if buf.size == 0 or buf.size % n:
            raise self._N.PanicError("assertion failed: buf.len() == batch * fft_size")
        self._N.check(self._N.lib.cfft_c64_host_multi(self._arr, len(self._replicas), op, buf.ctypes.data, buf.size, buf.size // n))

    def _f128(self, planes, op):
        import numpy as np

        n = self._plan.fft_size()
        for p in planes:
            if not isinstance(p, np.ndarray) or p.dtype != np.float64 or not p.flags["C_CONTIGUOUS"] or p.size != planes[0].size:
                raise TypeError("planes must be four C-contiguous numpy float64 arrays of one size (host memory)")
        if planes[0].size == 0 or planes[0].size % n:
            raise self._N.PanicError("assertion failed: buf.len() == batch * fft_size")
        self._N.check(self._N.lib.cfft_f128_host_multi(self._arr, len(self._replicas), op, *[p.ctypes.data for p in planes], planes[0].size,
                                                       planes[0].size // n))

    def fwd(self, *bufs):
        self._f128(bufs, 0) if self._is_f128() else self._c64(bufs[0], 0)

    def inv(self, *bufs):
        self._f128(bufs, 1) if self._is_f128() else self._c64(bufs[0], 1)

    def fwd_inv(self, *bufs):
        self._f128(bufs, 2) if self._is_f128() else self._c64(bufs[0], 2)
