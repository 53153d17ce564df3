"""CPU checks of the boundary: the shared library loads, exports every symbol the public header
declares, and rejects what the reference panics on before touching CUDA."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "cfft_b200.h")


@pytest.fixture(scope="module")
def C():
    lib = os.path.join(ROOT, "concrete_fft_b200", "libcfft_b200.so")
    if not os.path.exists(lib):
        import __graft_entry__ as g

        g.build()
    import concrete_fft_b200

    return concrete_fft_b200


def header_functions():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(cfft_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol(C):
    names = header_functions()
    assert len(names) >= 30
    lib = ctypes.CDLL(os.path.join(ROOT, "concrete_fft_b200", "libcfft_b200.so"))
    for name in names:
        assert hasattr(lib, name), "libcfft_b200.so does not export " + name
    # and the Python binding covers the same set
    assert sorted(C._native.EXPORTED_SYMBOLS) == names


def test_no_oracle_in_product():
    """The product must not include, link or call anything under oracle/."""
    for base, _, files in os.walk(os.path.join(ROOT, "concrete_fft_b200")):
        for f in files:
            if f.endswith((".py", ".cc", ".cu", ".cuh", ".h", ".sh")):
                text = open(os.path.join(base, f)).read()
                for needle in ("oracle.h", "oracle_lib", "liboracle", "orc_", "import oracle", "oracle/"):
                    assert needle not in text, (os.path.join(base, f), needle)


def test_argument_checks_match_reference_panics(C):
    U, Od, F = C.unordered, C.ordered, C.fft128
    A = Od.FftAlgo
    with pytest.raises(C.PanicError):
        Od.Plan(48, Od.Method.UserProvided(A.Dif4))  # not a power of two, src/ordered.rs:243
    with pytest.raises(C.PanicError):
        Od.Plan(2048, Od.Method.UserProvided(A.Dif4))  # > 2^10, src/ordered.rs:244
    for n, algo, base_n in [(2048, A.Dif4, 16), (2048, A.Dif4, 2048), (64, A.Dif4, 128), (100, A.Dif4, 32), (64, A.Dif4, 48)]:
        with pytest.raises(C.PanicError):
            U.Plan(n, U.Method.UserProvided(algo, base_n))  # src/unordered.rs:660-669
    for n in [16, 48, 0]:
        with pytest.raises(C.PanicError):
            F.Plan(n)  # src/fft128/mod.rs:1865-1866


def test_fused_product_entry_points_reject_bad_plans_before_cuda(C):
    """cfft_c64_fwd_mul_inv / cfft_f128_fwd_mul_inv: a null plan is CFFT_EINVAL (no CUDA call is made), and the capability
    query answers 0 for it."""
    lib = C._native.lib
    assert lib.cfft_c64_fwd_mul_inv(None, None, 1, None, 0, None, 0, None) == C._native.EINVAL
    assert b"c64 plan" in lib.cfft_last_error()
    assert lib.cfft_f128_fwd_mul_inv(None, None, None, None, None, None, None, None, None, 0, 1.0, 0, None) == C._native.EINVAL
    assert b"fft128 plan" in lib.cfft_last_error()
    assert lib.cfft_c64_fwd_mul_add(None, None, 0, None, 0, None, 0, 0, None) == C._native.EINVAL
    assert lib.cfft_plan_has_fused_mul_kernel(None) == 0


def test_strided_and_replica_entry_points_reject_bad_arguments_before_cuda(C):
    """cfft_*_strided / cfft_*_host_multi / cfft_plan_clone_to_device: null plans and empty replica lists are CFFT_EINVAL
    without any CUDA call."""
    lib, N = C._native.lib, C._native
    assert lib.cfft_c64_fwd_strided(None, None, 4096, 2, None) == N.EINVAL
    assert lib.cfft_c64_inv_strided(None, None, 4096, 2, None) == N.EINVAL
    assert lib.cfft_f128_fwd_strided(None, None, None, None, None, 4096, 2, None) == N.EINVAL
    assert lib.cfft_f128_inv_strided(None, None, None, None, None, 4096, 2, None) == N.EINVAL
    assert lib.cfft_c64_host_multi(None, 0, 0, None, 0, 0) == N.EINVAL
    assert b"replicas" in lib.cfft_last_error()
    assert lib.cfft_f128_host_multi(None, 2, 0, None, None, None, None, 0, 0) == N.EINVAL
    assert lib.cfft_plan_clone_to_device(None, 0, None) == N.EINVAL


def test_no_cpu_fallback(C):
    import torch

    if torch.cuda.is_available():
        pytest.skip("CUDA device present")
    with pytest.raises(C.CfftError) as e:
        C.unordered.Plan(1024, C.unordered.Method.UserProvided(C.ordered.FftAlgo.Dif4, 32))
    assert e.value.status == C._native.ECUDA


def test_method_and_enum_surface(C):
    A = C.ordered.FftAlgo
    assert [a.name for a in A] == ["Dif2", "Dit2", "Dif4", "Dit4", "Dif8", "Dit8", "Dif16", "Dit16"]
    assert C.ordered.Method.UserProvided(A.Dif4) == C.ordered.Method.UserProvided(A.Dif4)
    assert C.ordered.Method.UserProvided(A.Dif4) != C.ordered.Method.UserProvided(A.Dit4)
    m = C.unordered.Method.UserProvided(A.Dif16, 256)
    assert (m.base_algo, m.base_n) == (A.Dif16, 256)
    assert "cfft_b200" in C.version()


def test_header_is_plain_c_and_demo_links(C, tmp_path):
    """include/cfft_b200.h must compile as C99 with warnings as errors, and a C program must link
    against the shared library using nothing but that header."""
    import subprocess

    exe = tmp_path / "c_api_demo"
    cmd = ["gcc", "-std=c99", "-Wall", "-Wextra", "-Werror", "-pedantic", "-I", os.path.join(ROOT, "include"),
           os.path.join(ROOT, "examples", "c_api_demo.c"), "-L", os.path.join(ROOT, "concrete_fft_b200"), "-lcfft_b200",
           "-Wl,-rpath," + os.path.join(ROOT, "concrete_fft_b200"), "-lm", "-o", str(exe)]
    subprocess.run(cmd, check=True)
    rc = subprocess.run([str(exe)], capture_output=True, text=True)
    import torch

    if torch.cuda.is_available():
        assert rc.returncode == 0, rc.stdout + rc.stderr
    else:
        assert rc.returncode == 77, rc.stdout + rc.stderr  # fails loudly without a GPU: no CPU fallback


def test_rust_call_sequences_replayed_in_c(C, tmp_path):
    """The Rust crate (rust/concrete-fft-b200) cannot be compiled here (no rustc in the image), so every call sequence its
    methods make -- panics as statuses, serde invalid_length paths, clone / drop, as_raw() + device entry points, the
    polynomial host entry -- is replayed by a strict-C99 program against the same shared library."""
    import subprocess

    exe = tmp_path / "rust_call_sequences"
    cmd = ["gcc", "-std=c99", "-Wall", "-Wextra", "-Werror", "-pedantic", "-I", os.path.join(ROOT, "include"),
           os.path.join(ROOT, "examples", "rust_call_sequences.c"), "-L", os.path.join(ROOT, "concrete_fft_b200"), "-lcfft_b200",
           "-Wl,-rpath," + os.path.join(ROOT, "concrete_fft_b200"), "-L/usr/local/cuda/lib64", "-lcudart", "-Wl,-rpath,/usr/local/cuda/lib64",
           "-lm", "-o", str(exe)]
    subprocess.run(cmd, check=True)
    rc = subprocess.run([str(exe)], capture_output=True, text=True)
    import torch

    if torch.cuda.is_available():
        assert rc.returncode == 0, rc.stdout + rc.stderr
        assert "all sequences ok" in rc.stdout
    else:
        assert rc.returncode == 77, rc.stdout + rc.stderr


def test_rust_crate_declares_every_symbol_of_the_header():
    """rust/concrete-fft-b200/src/ffi.rs must bind exactly the functions include/cfft_b200.h declares (the crate is source only:
    this is the closest thing to a link check), and the f128 operator surface of the reference must be present by name."""
    import re

    declared = set(header_functions())
    ffi = open(os.path.join(ROOT, "rust", "concrete-fft-b200", "src", "ffi.rs")).read()
    bound = set(re.findall(r"pub fn (cfft_[a-z0-9_]+)\s*\(", ffi))
    assert declared == bound, (sorted(declared - bound), sorted(bound - declared))
    ops = open(os.path.join(ROOT, "rust", "concrete-fft-b200", "src", "fft128", "f128_ops.rs")).read()
    for name in ["add_f64_f64", "add_f128_f64", "add_f64_f128", "add_estimate_f128_f128", "add_f128_f128", "sub_f64_f64", "sub_f128_f64",
                 "sub_f64_f128", "sub_estimate_f128_f128", "sub_f128_f128", "mul_f64_f64", "mul_f128_f64", "mul_f64_f128", "mul_f128_f128",
                 "sqr", "div_f64_f64", "div_f128_f64", "div_f64_f128", "div_estimate_f128_f128", "div_f128_f128", "to_f64", "is_nan", "abs",
                 "sincospi"]:
        assert re.search(r"pub fn %s\b" % name, ops), name
    for tr in ["Neg for f128", "PartialEq<f128> for f128", "PartialEq<f64> for f128", "PartialEq<f128> for f64", "PartialOrd<f128> for f128",
               "PartialOrd<f64> for f128", "PartialOrd<f128> for f64", "From<f64> for f128"]:
        assert tr in ops, tr
    assert "binop!(Add, add, AddAssign" in ops and "binop!(Div, div, DivAssign" in ops
    lib = open(os.path.join(ROOT, "rust", "concrete-fft-b200", "src", "lib.rs")).read()
    assert lib.count("pub fn as_raw(&self)") == 3  # every plan type hands its handle to device::*
