"""ctypes binding of the C ABI (include/cfft_b200.h) -> concrete_fft_b200/libcfft_b200.so.

There is no fallback: if the CUDA library is missing this module raises at import, and every
compute call raises if no CUDA device is present.
"""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libcfft_b200.so")

OK, EINVAL, ECUDA, ENOMEM, EUNSUPPORTED, ELENGTH = 0, -1, -2, -3, -4, -5
METHOD_USER, METHOD_MEASURE = 0, 1
POLY_INTEGER, POLY_TORUS, POLY_ACCUMULATE = 0, 1, 2


class CfftError(RuntimeError):
    def __init__(self, status, message):
        super().__init__("cfft_b200 status %d: %s" % (status, message))
        self.status = status


class PanicError(AssertionError):
    """Raised where the reference's Rust API would panic (assert! / assert_eq!)."""


class InvalidLength(ValueError):
    """serde::de::Error::invalid_length (src/unordered.rs:1027-1028)."""


if not os.path.exists(LIB_PATH):
    raise ImportError(
        "%s is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` or "
        "concrete_fft_b200/csrc/build.sh (there is no CPU fallback)" % LIB_PATH
    )

lib = ctypes.CDLL(LIB_PATH)

_vp, _u64, _int = ctypes.c_void_p, ctypes.c_uint64, ctypes.c_int
_pp = ctypes.POINTER(ctypes.c_void_p)
_SIGNATURES = {
    "cfft_ordered_plan_create": (ctypes.c_int32, [_pp, _int, _u64, _int, _int, _int]),
    "cfft_unordered_plan_create": (ctypes.c_int32, [_pp, _int, _u64, _int, _int, _u64]),
    "cfft_f128_plan_create": (ctypes.c_int32, [_pp, _int, _u64]),
    "cfft_plan_destroy": (None, [_vp]),
    "cfft_plan_clone": (ctypes.c_int32, [_vp, _pp]),
    "cfft_plan_clone_to_device": (ctypes.c_int32, [_vp, _int, _pp]),
    "cfft_c64_host_multi": (ctypes.c_int32, [_pp, _int, _int, _vp, _u64, _u64]),
    "cfft_f128_host_multi": (ctypes.c_int32, [_pp, _int, _int, _vp, _vp, _vp, _vp, _u64, _u64]),
    "cfft_plan_fft_size": (_u64, [_vp]),
    "cfft_plan_algo": (ctypes.c_int32, [_vp, ctypes.POINTER(_int), ctypes.POINTER(_u64)]),
    "cfft_plan_scratch_req": (ctypes.c_int32, [_vp, ctypes.POINTER(_u64), ctypes.POINTER(_u64)]),
    "cfft_plan_kind": (_int, [_vp]),
    "cfft_plan_device": (_int, [_vp]),
    "cfft_plan_kernel_name": (ctypes.c_char_p, [_vp]),
    "cfft_plan_autotune": (ctypes.c_int32, [_vp, _u64]),
    "cfft_plan_tuning_report": (_u64, [_vp, ctypes.c_char_p, _u64]),
    "cfft_c64_fwd": (ctypes.c_int32, [_vp, _vp, _u64, _vp]),
    "cfft_c64_inv": (ctypes.c_int32, [_vp, _vp, _u64, _vp]),
    "cfft_c64_fwd_strided": (ctypes.c_int32, [_vp, _vp, _u64, _u64, _vp]),
    "cfft_c64_inv_strided": (ctypes.c_int32, [_vp, _vp, _u64, _u64, _vp]),
    "cfft_f128_fwd_strided": (ctypes.c_int32, [_vp, _vp, _vp, _vp, _vp, _u64, _u64, _vp]),
    "cfft_f128_inv_strided": (ctypes.c_int32, [_vp, _vp, _vp, _vp, _vp, _u64, _u64, _vp]),
    "cfft_c64_fwd_host": (ctypes.c_int32, [_vp, _vp, _u64, _u64]),
    "cfft_c64_inv_host": (ctypes.c_int32, [_vp, _vp, _u64, _u64]),
    "cfft_c64_fwd_inv_host": (ctypes.c_int32, [_vp, _vp, _u64, _u64]),
    "cfft_unordered_fwd_monomial": (ctypes.c_int32, [_vp, _u64, _vp, _vp]),
    "cfft_unordered_fwd_monomial_host": (ctypes.c_int32, [_vp, _u64, _vp, _u64]),
    "cfft_unordered_permutation": (ctypes.c_int32, [_vp, _vp]),
    "cfft_unordered_to_standard": (ctypes.c_int32, [_vp, _vp, _vp, _u64, _vp]),
    "cfft_unordered_from_standard": (ctypes.c_int32, [_vp, _vp, _vp, _u64, _vp]),
    "cfft_unordered_to_standard_host": (ctypes.c_int32, [_vp, _vp, _vp]),
    "cfft_unordered_from_standard_host": (ctypes.c_int32, [_vp, _vp, _u64, _vp]),
    "cfft_f128_fwd": (ctypes.c_int32, [_vp, _vp, _vp, _vp, _vp, _u64, _vp]),
    "cfft_f128_inv": (ctypes.c_int32, [_vp, _vp, _vp, _vp, _vp, _u64, _vp]),
    "cfft_f128_fwd_host": (ctypes.c_int32, [_vp, _vp, _vp, _vp, _vp, _u64, _u64]),
    "cfft_f128_inv_host": (ctypes.c_int32, [_vp, _vp, _vp, _vp, _vp, _u64, _u64]),
    "cfft_f128_fwd_inv_host": (ctypes.c_int32, [_vp, _vp, _vp, _vp, _vp, _u64, _u64]),
    "cfft_f128_binary_op": (ctypes.c_int32, [_int, _int, _vp, _vp, _vp, _vp, _vp, _vp, _u64, _vp]),
    "cfft_f128_unary_op": (ctypes.c_int32, [_int, _int, _vp, _vp, _vp, _vp, _vp, _vp, _u64, _vp]),
    "cfft_f128_compare": (ctypes.c_int32, [_int, _vp, _vp, _vp, _vp, _vp, _u64, _vp]),
    "cfft_f128_cplx_mul_scale": (ctypes.c_int32, [_int, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, ctypes.c_double, _u64, _vp]),
    "cfft_c64_mul_assign": (ctypes.c_int32, [_int, _vp, _vp, _u64, _vp]),
    "cfft_c64_mul_add_assign": (ctypes.c_int32, [_int, _vp, _vp, _vp, _u64, _vp]),
    "cfft_f128_fwd_mul_inv": (ctypes.c_int32, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _u64, ctypes.c_double, _u64, _vp]),
    "cfft_c64_fwd_mul_inv": (ctypes.c_int32, [_vp, _vp, _u64, _vp, _u64, _vp, _u64, _vp]),
    "cfft_c64_fwd_mul_add": (ctypes.c_int32, [_vp, _vp, _u64, _vp, _u64, _vp, _int, _u64, _vp]),
    "cfft_c64_fwd_mul_inv_multi": (ctypes.c_int32, [_vp, _vp, _u64, _vp, _u64, _u64, _vp, _u64, _vp]),
    "cfft_plan_has_fused_mul2_kernel": (_int, [_vp]),
    "cfft_plan_has_fused_mul_kernel": (_int, [_vp]),
    "cfft_status_string": (ctypes.c_char_p, [ctypes.c_int32]),
    "cfft_last_error": (ctypes.c_char_p, []),
    "cfft_launch_count": (_u64, []),
    "cfft_version": (ctypes.c_char_p, []),
    "cfft_plan_copy_twiddles": (ctypes.c_int32, [_vp, _int, _vp, _u64]),
    "cfft_c64_poly_fwd": (ctypes.c_int32, [_vp, _vp, _vp, _u64, ctypes.c_uint32, _vp]),
    "cfft_c64_poly_inv": (ctypes.c_int32, [_vp, _vp, _vp, _u64, ctypes.c_uint32, _vp]),
    "cfft_c64_poly_mul": (ctypes.c_int32, [_vp, _vp, _u64, _vp, _u64, _vp, _u64, ctypes.c_uint32, _vp]),
    "cfft_c64_poly_mul_host": (ctypes.c_int32, [_vp, _vp, _u64, _vp, _u64, _vp, _u64, ctypes.c_uint32]),
    "cfft_plan_has_fused_poly_kernel": (_int, [_vp, _u64]),
    "cfft_plan_copy_twist": (ctypes.c_int32, [_vp, _vp, _u64]),
    "cfft_twopass_timeouts": (ctypes.c_int32, [_int, ctypes.POINTER(ctypes.c_uint32)]),
    "cfft_probe_fp64_issue_rate": (ctypes.c_int32, [_int] + [ctypes.POINTER(ctypes.c_double)] * 4 + [ctypes.POINTER(_int)]),
}
EXPORTED_SYMBOLS = sorted(_SIGNATURES)
for _name, (_res, _args) in _SIGNATURES.items():
    _f = getattr(lib, _name)  # AttributeError here = the .so does not export a declared symbol
    _f.restype, _f.argtypes = _res, _args


def check(status, panic_on=(EINVAL, ELENGTH)):
    """Map a non-zero status to the exception the Rust shim would turn into a panic."""
    if status == OK:
        return
    msg = lib.cfft_last_error().decode() or lib.cfft_status_string(status).decode()
    if status in panic_on:
        raise PanicError(msg)
    raise CfftError(status, msg)


def launch_count():
    return int(lib.cfft_launch_count())


def probe_fp64_issue_rate(device=0):
    """Measured FP64 issue rate of `device`: {"dfma_per_s", "dadd_per_s", "mix_per_s", "sm_mhz", "sm_count"}."""
    vals = [ctypes.c_double() for _ in range(4)]
    sms = _int()
    check(lib.cfft_probe_fp64_issue_rate(device, *[ctypes.byref(v) for v in vals], ctypes.byref(sms)))
    return {"dfma_per_s": vals[0].value, "dadd_per_s": vals[1].value, "mix_per_s": vals[2].value, "sm_mhz": vals[3].value,
            "sm_count": sms.value}


def version():
    return lib.cfft_version().decode()
