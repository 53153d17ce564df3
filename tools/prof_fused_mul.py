#!/usr/bin/env python3
"""Two warm-up calls, then one cfft_c64_fwd_mul_inv call, for ncu captures:
    ncu --set full --clock-control none --import-source on -k regex:fwd_mul_inv -s 2 -c 1 -o gpurun_out/x \
        python tools/prof_fused_mul.py <n> <k_terms> <rows> [per-row]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch

import concrete_fft_b200 as C

if sys.argv[1] == "f128":  # python tools/prof_fused_mul.py f128 <n> <rows>   (-k regex:f128_fwd_mul_inv)
    n, rows = int(sys.argv[2]), int(sys.argv[3])
    plan = C.fft128.Plan(n)
    L = [torch.rand(rows, n, dtype=torch.float64, device="cuda") * (1.0 if i % 2 == 0 else 1e-17) for i in range(4)]
    R = [torch.rand(n, dtype=torch.float64, device="cuda") * (1.0 if i % 2 == 0 else 1e-17) for i in range(4)]
    for _ in range(3):
        plan.fwd_mul_inv(L, R, 0.5 / n)
    torch.cuda.synchronize()
    print("fused kernel:", plan.has_fused_mul_kernel())
    sys.exit(0)
n, k, rows = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
per_row = len(sys.argv) > 4
plan = C.unordered.Plan(n, C.unordered.Method.UserProvided(C.ordered.FftAlgo.Dif16, 256))
a = torch.view_as_complex(torch.rand(rows, k, n, 2, dtype=torch.float64, device="cuda") - 0.5)
b = torch.view_as_complex(torch.rand(*((rows, k, n, 2) if per_row else (k, n, 2)), dtype=torch.float64, device="cuda") - 0.5)
out = torch.empty(rows, n, dtype=torch.complex128, device="cuda")
for _ in range(3):
    plan.fwd_mul_inv(a, b, out=out)
torch.cuda.synchronize()
print("fused kernel:", plan.has_fused_mul_kernel())
