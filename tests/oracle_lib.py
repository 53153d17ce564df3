"""ctypes loader for the CPU oracle (oracle/).  Test infrastructure only.

Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs import this.
"""
import ctypes
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")

DIF2, DIT2, DIF4, DIT4, DIF8, DIT8, DIF16, DIT16 = range(8)
ALGO_NAMES = ["Dif2", "Dit2", "Dif4", "Dit4", "Dif8", "Dit8", "Dif16", "Dit16"]
F128_SCALAR, F128_FMA = 0, 1

_libs = {}


def build():
    subprocess.run(["make", "-s", "-C", ORACLE_DIR], check=True, env=dict(os.environ, CC="gcc"))


def lib(fast=False):
    name = "liboracle_fast.so" if fast else "liboracle.so"
    if name in _libs:
        return _libs[name]
    path = os.path.join(ORACLE_DIR, "_build", name)
    srcs = [os.path.join(ORACLE_DIR, f) for f in os.listdir(ORACLE_DIR) if f.endswith((".c", ".h"))]
    if not os.path.exists(path) or any(os.path.getmtime(s) > os.path.getmtime(path) for s in srcs):
        build()
    L = ctypes.CDLL(path)
    vp, sz, ci, dbl = ctypes.c_void_p, ctypes.c_size_t, ctypes.c_int, ctypes.c_double
    sig = {
        "orc_sincospi64": (None, [dbl, vp, vp]),
        "orc_init_wt": (None, [sz, sz, vp, vp]),
        "orc_ordered_plan_new": (vp, [sz, ci]),
        "orc_ordered_plan_free": (None, [vp]),
        "orc_ordered_fwd": (None, [vp, vp, vp]),
        "orc_ordered_inv": (None, [vp, vp, vp]),
        "orc_unordered_plan_new": (vp, [sz, ci, sz]),
        "orc_unordered_plan_free": (None, [vp]),
        "orc_unordered_fwd": (None, [vp, vp, vp]),
        "orc_unordered_inv": (None, [vp, vp, vp]),
        "orc_unordered_fwd_monomial": (None, [vp, sz, vp]),
        "orc_unordered_twiddles": (vp, [vp, ci]),
        "orc_unordered_fwd_batch": (None, [vp, vp, sz, ci]),
        "orc_unordered_inv_batch": (None, [vp, vp, sz, ci]),
        "orc_bit_rev_twice": (sz, [ctypes.c_uint, ctypes.c_uint, sz]),
        "orc_bit_rev_twice_inv": (sz, [ctypes.c_uint, ctypes.c_uint, sz]),
        "orc_f128_init_twiddles": (None, [sz, vp, vp, vp, vp]),
        "orc_f128_plan_new": (vp, [sz]),
        "orc_f128_plan_free": (None, [vp]),
        "orc_f128_fwd": (None, [vp, vp, vp, vp, vp, ci]),
        "orc_f128_inv": (None, [vp, vp, vp, vp, vp, ci]),
        "orc_f128_fwd_batch": (None, [vp, vp, vp, vp, vp, sz, ci, ci]),
        "orc_f128_inv_batch": (None, [vp, vp, vp, vp, vp, sz, ci, ci]),
        "orc_f128_twiddles": (vp, [vp, ci]),
        "orc_f128_binary_op": (None, [ci, vp, vp, vp, vp, vp, vp, sz]),
        "orc_f128_cplx_mul_scale": (None, [vp, vp, vp, vp, vp, vp, vp, vp, dbl, sz]),
        "orc_f128_unary_op": (None, [ci, vp, vp, vp, vp, vp, vp, sz]),
        "orc_f128_compare": (None, [vp, vp, vp, vp, vp, sz]),
        "orc_c64_pointwise": (None, [vp, vp, vp, sz]),
        "orc_poly_twist_tables": (None, [sz, vp, vp]),
        "orc_poly_fold_twist": (None, [sz, ci, vp, vp, vp]),
        "orc_poly_untwist_round": (None, [sz, ci, ci, vp, vp, vp]),
        "orc_poly_mul_batch": (None, [vp, sz, sz, vp, sz, vp, sz, vp, sz, ci, ci, ci]),
    }
    for k, (res, args) in sig.items():
        f = getattr(L, k)
        f.restype, f.argtypes = res, args
    _libs[name] = L
    return L


def _ptr(a):
    assert a.flags["C_CONTIGUOUS"]
    return a.ctypes.data


def sincospi64(a):
    s, c = ctypes.c_double(), ctypes.c_double()
    lib().orc_sincospi64(a, ctypes.addressof(s), ctypes.addressof(c))
    return s.value, c.value


def bit_rev_twice(n, base_n, i):
    return lib().orc_bit_rev_twice(n.bit_length() - 1, base_n.bit_length() - 1, i)


def permutation(n, base_n):
    """pi[i] = position of frequency i in the unordered plan's output."""
    return np.array([bit_rev_twice(n, base_n, i) for i in range(n)], dtype=np.int64)


class OrderedPlan:
    def __init__(self, n, algo, fast=False):
        self.L = lib(fast)
        self.n, self.algo = n, algo
        self.h = self.L.orc_ordered_plan_new(n, algo)
        if not self.h:
            raise ValueError("invalid ordered plan parameters")

    def __del__(self):
        if getattr(self, "h", None):
            self.L.orc_ordered_plan_free(self.h)

    def _run(self, fn, x):
        buf = np.ascontiguousarray(x, dtype=np.complex128).copy()
        scr = np.zeros(self.n, np.complex128)
        for row in buf.reshape(-1, self.n):
            fn(self.h, _ptr(row), _ptr(scr))
        return buf

    def fwd(self, x):
        return self._run(self.L.orc_ordered_fwd, x)

    def inv(self, x):
        return self._run(self.L.orc_ordered_inv, x)


class UnorderedPlan:
    def __init__(self, n, base_algo, base_n, fast=False):
        self.L = lib(fast)
        self.n, self.base_algo, self.base_n = n, base_algo, base_n
        self.h = self.L.orc_unordered_plan_new(n, base_algo, base_n)
        if not self.h:
            raise ValueError("invalid unordered plan parameters")

    def __del__(self):
        if getattr(self, "h", None):
            self.L.orc_unordered_plan_free(self.h)

    def _run(self, fn, x, threads):
        buf = np.ascontiguousarray(x, dtype=np.complex128).copy()
        fn(self.h, _ptr(buf), buf.size // self.n, threads)
        return buf

    def fwd(self, x, threads=1):
        return self._run(self.L.orc_unordered_fwd_batch, x, threads)

    def inv(self, x, threads=1):
        return self._run(self.L.orc_unordered_inv_batch, x, threads)

    def fwd_inplace(self, buf, threads=1):
        self.L.orc_unordered_fwd_batch(self.h, _ptr(buf), buf.size // self.n, threads)

    def inv_inplace(self, buf, threads=1):
        self.L.orc_unordered_inv_batch(self.h, _ptr(buf), buf.size // self.n, threads)

    def fwd_monomial(self, degree):
        buf = np.zeros(self.n, np.complex128)
        self.L.orc_unordered_fwd_monomial(self.h, degree, _ptr(buf))
        return buf

    def twiddles(self, inverse=False):
        p = self.L.orc_unordered_twiddles(self.h, int(inverse))
        cnt = self.n + self.base_n
        return np.ctypeslib.as_array(ctypes.cast(p, ctypes.POINTER(ctypes.c_double)), (cnt * 2,)).view(np.complex128).copy()


class F128Plan:
    def __init__(self, n, fast=False):
        self.L = lib(fast)
        self.n = n
        self.h = self.L.orc_f128_plan_new(n)
        if not self.h:
            raise ValueError("invalid fft128 plan size")

    def __del__(self):
        if getattr(self, "h", None):
            self.L.orc_f128_plan_free(self.h)

    def _run(self, fn, planes, variant, threads):
        out = [np.ascontiguousarray(p, dtype=np.float64).copy() for p in planes]
        fn(self.h, *[_ptr(p) for p in out], out[0].size // self.n, variant, threads)
        return out

    def fwd(self, re0, re1, im0, im1, variant=F128_FMA, threads=1):
        return self._run(self.L.orc_f128_fwd_batch, (re0, re1, im0, im1), variant, threads)

    def inv(self, re0, re1, im0, im1, variant=F128_FMA, threads=1):
        return self._run(self.L.orc_f128_inv_batch, (re0, re1, im0, im1), variant, threads)

    def fwd_inplace(self, planes, variant=F128_FMA, threads=1):
        self.L.orc_f128_fwd_batch(self.h, *[_ptr(p) for p in planes], planes[0].size // self.n, variant, threads)

    def inv_inplace(self, planes, variant=F128_FMA, threads=1):
        self.L.orc_f128_inv_batch(self.h, *[_ptr(p) for p in planes], planes[0].size // self.n, variant, threads)

    def twiddles(self):
        out = []
        for w in range(4):
            p = self.L.orc_f128_twiddles(self.h, w)
            out.append(np.ctypeslib.as_array(ctypes.cast(p, ctypes.POINTER(ctypes.c_double)), (self.n,)).copy())
        return out


F128_OPS = {"add": 0, "sub": 1, "mul": 2, "div": 3, "add_estimate": 4, "sub_estimate": 5, "div_estimate": 6,
            "add_f128_f64": 7, "sub_f128_f64": 8, "sub_f64_f128": 9, "mul_f128_f64": 10, "div_f128_f64": 11, "div_f64_f128": 12,
            "add_f64_f64": 13, "sub_f64_f64": 14, "mul_f64_f64": 15, "div_f64_f64": 16}
F128_UNARY = {"sqr": 0, "abs": 1, "neg": 2, "sincospi": 3, "is_nan": 4}


def f128_binary_op(op, a_hi, a_lo, b_hi, b_lo):
    """a_lo / b_lo may be None for the f64 operands of the mixed forms"""
    arrs = [None if x is None else np.ascontiguousarray(x, dtype=np.float64) for x in (a_hi, a_lo, b_hi, b_lo)]
    out_hi, out_lo = np.empty_like(arrs[0]), np.empty_like(arrs[0])
    lib().orc_f128_binary_op(F128_OPS[op], *[None if x is None else _ptr(x) for x in arrs], _ptr(out_hi), _ptr(out_lo), arrs[0].size)
    return out_hi, out_lo


def f128_unary_op(op, a_hi, a_lo):
    a_hi, a_lo = np.ascontiguousarray(a_hi, dtype=np.float64), np.ascontiguousarray(a_lo, dtype=np.float64)
    o = [np.empty_like(a_hi) for _ in range(4)]
    lib().orc_f128_unary_op(F128_UNARY[op], _ptr(a_hi), _ptr(a_lo), *[_ptr(x) for x in o], a_hi.size)
    return ((o[0], o[1]), (o[2], o[3])) if op == "sincospi" else (o[0], o[1])


def f128_compare(a_hi, a_lo, b_hi, b_lo=None):
    arrs = [np.ascontiguousarray(x, dtype=np.float64) for x in (a_hi, a_lo, b_hi)]
    bl = None if b_lo is None else np.ascontiguousarray(b_lo, dtype=np.float64)
    out = np.empty(arrs[0].shape, np.int8)
    lib().orc_f128_compare(*[_ptr(x) for x in arrs], None if bl is None else _ptr(bl), _ptr(out), arrs[0].size)
    return out


def f128_cplx_mul_scale(lhs, rhs, factor):
    L = [np.ascontiguousarray(x, dtype=np.float64).copy() for x in lhs]
    R = [np.ascontiguousarray(x, dtype=np.float64) for x in rhs]
    lib().orc_f128_cplx_mul_scale(*[_ptr(x) for x in L], *[_ptr(x) for x in R], float(factor), L[0].size)
    return L


def c64_pointwise(a, b, acc=None):
    """a * b (acc is None) or acc + a * b, element-wise, num_complex semantics (no FMA)."""
    A = np.ascontiguousarray(a, dtype=np.complex128).copy()
    B = np.ascontiguousarray(b, dtype=np.complex128)
    if acc is None:
        lib().orc_c64_pointwise(None, _ptr(A), _ptr(B), A.size)
        return A
    C = np.ascontiguousarray(acc, dtype=np.complex128).copy()
    lib().orc_c64_pointwise(_ptr(C), _ptr(A), _ptr(B), A.size)
    return C


# ---- caller-side steps around the c64 transform (oracle/poly_oracle.c) ----------------------------------------

def poly_twist_tables(n):
    tw, un = np.empty(n, np.complex128), np.empty(n, np.complex128)
    lib().orc_poly_twist_tables(n, _ptr(tw), _ptr(un))
    return tw, un


def poly_fold_twist(poly, torus=False):
    """[rows, 2n] int64 -> [rows, n] complex128: fold, convert (torus: x 2^-64), twist."""
    poly = np.ascontiguousarray(poly, dtype=np.int64)
    rows, n = poly.reshape(-1, poly.shape[-1]).shape[0], poly.shape[-1] // 2
    tw, _ = poly_twist_tables(n)
    out = np.empty((rows, n), np.complex128)
    flat = poly.reshape(rows, 2 * n)
    for r in range(rows):
        lib().orc_poly_fold_twist(n, int(torus), _ptr(flat[r]), _ptr(tw), _ptr(out[r]))
    return out.reshape(poly.shape[:-1] + (n,))


def poly_untwist_round(z, torus=False, acc=None):
    """[rows, n] complex128 (output of the unnormalised inverse) -> [rows, 2n] int64."""
    z = np.ascontiguousarray(z, dtype=np.complex128)
    n = z.shape[-1]
    flat = z.reshape(-1, n)
    _, un = poly_twist_tables(n)
    out = np.zeros((flat.shape[0], 2 * n), np.int64) if acc is None else np.ascontiguousarray(acc, dtype=np.int64).reshape(-1, 2 * n).copy()
    for r in range(flat.shape[0]):
        lib().orc_poly_untwist_round(n, int(torus), int(acc is not None), _ptr(flat[r]), _ptr(un), _ptr(out[r]))
    return out


def poly_mul(plan, a, b, torus=False, acc=None, threads=1):
    """The whole product step through the oracle: a [batch, k, 2n] int64, b [k, n] (shared) or [batch, k, n] complex128
    in `plan`'s order (an UnorderedPlan) -> [batch, 2n] int64."""
    a = np.ascontiguousarray(a, dtype=np.int64)
    b = np.ascontiguousarray(b, dtype=np.complex128)
    batch, k, n = a.shape[0], a.shape[1], a.shape[2] // 2
    stride = 0 if b.ndim == 2 else k * n
    out = np.zeros((batch, 2 * n), np.int64) if acc is None else np.ascontiguousarray(acc, dtype=np.int64).copy()
    plan.L.orc_poly_mul_batch(plan.h, n, plan.base_n, _ptr(a), k, _ptr(b), stride, _ptr(out), batch, int(torus), int(acc is not None), threads)
    return out
