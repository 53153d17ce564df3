#!/usr/bin/env python3
"""Small invocation of every kernel family for compute-sanitizer (memcheck / racecheck / synccheck)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch

import concrete_fft_b200 as C

A = C.ordered.FftAlgo
rng = np.random.default_rng(0)


def run_fused_mul():
    """cfft_c64_fwd_mul_inv: the one-kernel path at every size (1 and 3 terms, ragged tile), the composed path once"""
    for n in [256, 512, 1024, 2048, 4096]:
        plan = C.unordered.Plan(n, C.unordered.Method.UserProvided(A.Dif16, 256))
        for k in (1, 3):
            a = torch.from_numpy(rng.random((5, k, n)) + 1j * rng.random((5, k, n))).cuda()
            b = torch.from_numpy(rng.random((k, n)) + 1j * rng.random((k, n))).cuda()
            plan.fwd_mul_inv(a, b)
    plan = C.unordered.Plan(2048, C.unordered.Method.UserProvided(A.Dif4, 32))
    a = torch.from_numpy(rng.random((3, 2, 2048)) + 1j * rng.random((3, 2, 2048))).cuda()
    plan.fwd_mul_inv(a, a.clone())
    # cfft_f128_fwd_mul_inv: one kernel (ragged multi-row tile, 2048- and 4096-point tiles), then the three launches
    for n, batch in [(64, 70), (2048, 3), (4096, 2), (8192, 2)]:
        p128 = C.fft128.Plan(n)
        lhs = [torch.rand(batch, n, dtype=torch.float64, device="cuda") for _ in range(4)]
        p128.fwd_mul_inv(lhs, [torch.rand(n, dtype=torch.float64, device="cuda") for _ in range(4)], 1.0 / n)
        p128.fwd_mul_inv(lhs, [torch.rand(batch, n, dtype=torch.float64, device="cuda") for _ in range(4)], 1.0 / n)
    torch.cuda.synchronize()


if "--only-fused-mul" in sys.argv:
    run_fused_mul()
    print("sanitize_small (fused mul only) done, launches =", C.launch_count())
    sys.exit(0)


def run_c64(plan, n, batch):
    x = torch.from_numpy(rng.random((batch, n)) + 1j * rng.random((batch, n))).cuda()
    plan.fwd(x)
    plan.inv(x)
    torch.cuda.synchronize()


for n in [256, 512, 1024, 2048, 4096, 8192, 16384, 131072]:
    run_c64(C.unordered.Plan(n, C.unordered.Method.UserProvided(A.Dif16, 256)), n, 3)
for n, algo, base in [(2048, A.Dif4, 32), (1024, A.Dit8, 64), (16384, A.Dit16, 1024), (64, A.Dif2, 64), (8, A.Dif16, 8)]:
    run_c64(C.unordered.Plan(n, C.unordered.Method.UserProvided(algo, base)), n, 5)
for n in [512, 1024]:
    run_c64(C.ordered.Plan(n, C.ordered.Method.UserProvided(A.Dit16)), n, 3)
for n in [2048, 4096, 65536, 131072]:
    run_c64(C.ordered.Plan(n, C.ordered.Method.Measure(), allow_large=True), n, 2)
os.environ["CFFT_B200_FAST_VARIANT"] = "2"
run_c64(C.unordered.Plan(2048, C.unordered.Method.UserProvided(A.Dif16, 256)), 2048, 3)
del os.environ["CFFT_B200_FAST_VARIANT"]
# round-1 additions: whole-transform Dif16 register kernels, fused standard-order kernel, cluster and
# persistent two-phase kernels
for n in [32, 64, 128, 512, 1024]:
    run_c64(C.ordered.Plan(n, C.ordered.Method.UserProvided(A.Dif16)), n, 70)
for n in [2048, 4096, 8192]:
    run_c64(C.ordered.Plan(n, C.ordered.Method.UserProvided(A.Dif16), allow_large=True), n, 3)
for var, sizes in [("4", [8192, 16384]), ("8", [16384, 32768, 65536])]:
    os.environ["CFFT_B200_FAST_VARIANT"] = var
    for n in sizes:
        run_c64(C.unordered.Plan(n, C.unordered.Method.UserProvided(A.Dif16, 256)), n, 5)
    del os.environ["CFFT_B200_FAST_VARIANT"]
for n in [32, 256, 1024, 2048, 4096, 16384]:
    p = C.fft128.Plan(n)
    planes = [torch.rand(3, n, dtype=torch.float64, device="cuda") for _ in range(4)]
    p.fwd(*planes)
    p.inv(*planes)
    torch.cuda.synchronize()
p = C.unordered.Plan(1024, C.unordered.Method.UserProvided(A.Dif4, 32))
buf = torch.zeros(1024, dtype=torch.complex128, device="cuda")
p.fwd_monomial(5, buf)
std = p.serialize_fourier_buffer(buf)
p.deserialize_fourier_buffer(std, buf)
h = rng.random((7, 1024)) + 0j
p.fwd(h)
torch.cuda.synchronize()
run_fused_mul()
# round-2 additions: compile-time schedules, batches longer than one wave of resident CTAs (the L2 prefetch distance), strided
# rows, polynomial entry points, replicas, the fft128 HBM pass with its prefetch
for n, algo, base in [(2048, A.Dif16, 1024), (2048, A.Dif8, 512), (2048, A.Dif4, 32), (1024, A.Dif8, 512), (4096, A.Dif16, 1024)]:
    run_c64(C.unordered.Plan(n, C.unordered.Method.UserProvided(algo, base)), n, 7)
for n, algo in [(1024, A.Dif8), (512, A.Dit8), (1024, A.Dit16)]:
    run_c64(C.ordered.Plan(n, C.ordered.Method.UserProvided(algo)), n, 9)
for n, batch in [(2048, 700), (4096, 350), (8192, 170)]:
    run_c64(C.unordered.Plan(n, C.unordered.Method.UserProvided(A.Dif16, 256)), n, batch)
    run_c64(C.ordered.Plan(n, C.ordered.Method.UserProvided(A.Dif16), allow_large=True), n, batch)
run_c64(C.unordered.Plan(16384, C.unordered.Method.UserProvided(A.Dif16, 256)), 16384, 80)
plan = C.unordered.Plan(2048, C.unordered.Method.UserProvided(A.Dif16, 256))
rec = torch.from_numpy(rng.random((6, 3, 2048)) + 1j * rng.random((6, 3, 2048))).cuda()
plan.fwd_strided(rec[:, 1])
plan.inv_strided(rec[:, 1])
gen = C.unordered.Plan(2048, C.unordered.Method.UserProvided(A.Dif4, 32))
gen.fwd_strided(rec[:, 2])
a = torch.from_numpy(rng.random((700, 4, 2048)) + 1j * rng.random((700, 4, 2048))).cuda()
b = torch.from_numpy(rng.random((4, 2048)) + 1j * rng.random((4, 2048))).cuda()
plan.fwd_mul_inv(a, b)
big = C.unordered.Plan(8192, C.unordered.Method.UserProvided(A.Dif16, 256))
a8 = torch.from_numpy(rng.random((170, 2, 8192)) + 1j * rng.random((170, 2, 8192))).cuda()
big.fwd_mul_inv(a8, torch.from_numpy(rng.random((2, 8192)) + 1j * rng.random((2, 8192))).cuda())
for n2, rows in [(512, 9), (1024, 5), (2048, 310), (4096, 3)]:  # two outputs per row: one kernel (512 .. 2048), output by output (4096)
    p2 = C.unordered.Plan(n2, C.unordered.Method.UserProvided(A.Dif16, 256))
    a2 = torch.from_numpy(rng.random((rows, 3, n2)) + 1j * rng.random((rows, 3, n2))).cuda()
    p2.fwd_mul_inv_multi(a2, torch.from_numpy(rng.random((3, 2, n2)) + 1j * rng.random((3, 2, n2))).cuda())
poly = torch.from_numpy(rng.integers(-1000, 1000, size=(5, 2, 4096))).cuda()
key = plan.fwd_poly(torch.from_numpy(rng.integers(-8, 8, size=(2, 4096))).cuda())
out = plan.poly_mul(poly, key)
plan.inv_poly(plan.fwd_poly(out))
from concrete_fft_b200.sharding import MultiGpu

h = rng.random((9, 2048)) + 0j
MultiGpu(plan, [0, 0]).fwd_inv(h)
p128 = C.fft128.Plan(8192)
planes = [torch.rand(310, 8192, dtype=torch.float64, device="cuda") for _ in range(4)]
p128.fwd(*planes)
p128.inv(*planes)
rec128 = [torch.rand(4, 2, 1024, dtype=torch.float64, device="cuda") for _ in range(4)]
C.fft128.Plan(1024).fwd_strided(*[r[:, 1] for r in rec128])
torch.cuda.synchronize()
print("sanitize_small done, launches =", C.launch_count())
