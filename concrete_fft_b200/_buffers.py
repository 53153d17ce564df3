"""Argument adaptation shared by the Plan mirrors: numpy host arrays go to the *_host entry
points, torch CUDA tensors go to the device entry points on torch's current stream."""
import numpy as np

try:  # torch is plumbing only (device memory + streams)
    import torch
except Exception:  # pragma: no cover
    torch = None


def is_torch(x):
    return torch is not None and isinstance(x, torch.Tensor)


def current_stream_ptr(device_index):
    return int(torch.cuda.current_stream(device_index).cuda_stream)


def c64_view(buf, n):
    """-> (kind, ptr, length_in_c64, batch, device_index)."""
    if is_torch(buf):
        if buf.dtype != torch.complex128 or not buf.is_contiguous():
            raise TypeError("buf must be a contiguous complex128 tensor")
        if not buf.is_cuda:
            raise TypeError("torch buffers must live on a CUDA device (pass numpy arrays for host memory)")
        length = buf.numel()
        return "device", buf.data_ptr(), length, (length // n if n else 0), buf.device.index
    if not isinstance(buf, np.ndarray) or buf.dtype != np.complex128 or not buf.flags["C_CONTIGUOUS"]:
        raise TypeError("buf must be a C-contiguous numpy complex128 array")
    if not buf.flags["WRITEABLE"]:
        raise TypeError("buf must be writeable (the transform is in place)")
    length = buf.size
    return "host", buf.ctypes.data, length, (length // n if n else 0), None


def f64_view(buf):
    if is_torch(buf):
        if buf.dtype != torch.float64 or not buf.is_contiguous() or not buf.is_cuda:
            raise TypeError("planes must be contiguous float64 CUDA tensors")
        return "device", buf.data_ptr(), buf.numel(), buf.device.index
    if not isinstance(buf, np.ndarray) or buf.dtype != np.float64 or not buf.flags["C_CONTIGUOUS"]:
        raise TypeError("planes must be C-contiguous numpy float64 arrays")
    return "host", buf.ctypes.data, buf.size, None
