"""concrete_fft_b200 -- B200-native drop-in for concrete-fft's transform hot path.

Host-side mirror of the reference's Rust API (same module / type / method names):

    concrete_fft_b200.ordered.{Plan, Method, FftAlgo}
    concrete_fft_b200.unordered.{Plan, Method}
    concrete_fft_b200.fft128.{Plan, f128}
    concrete_fft_b200.pointwise.{mul_assign, mul_add_assign}   (the caller-side Fourier-domain step)

over the C ABI in include/cfft_b200.h (concrete_fft_b200/libcfft_b200.so, hand-written CUDA
for sm_100a).  Importing this package fails loudly if the CUDA library has not been built.
"""
import numpy as _np

from . import _native  # noqa: F401  (raises ImportError if libcfft_b200.so is missing)
from . import fft128, ordered, pointwise, unordered  # noqa: F401
from ._native import CfftError, InvalidLength, PanicError, launch_count, probe_fp64_issue_rate, version  # noqa: F401

c64 = _np.complex128  # src/lib.rs:84
__all__ = ["ordered", "unordered", "fft128", "pointwise", "c64", "CfftError", "PanicError", "InvalidLength", "launch_count", "version", "probe_fp64_issue_rate"]
