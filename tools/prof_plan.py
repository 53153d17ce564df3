#!/usr/bin/env python3
"""Two warm-up rounds then one fwd + inv of an arbitrary plan (for ncu):
    python tools/prof_plan.py ordered <n> <algo> <batch>   |   unordered <n> <algo> <base_n> <batch>"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch

import concrete_fft_b200 as C

A = C.ordered.FftAlgo
kind, n, algo = sys.argv[1], int(sys.argv[2]), A[sys.argv[3]]
if kind == "ordered":
    batch = int(sys.argv[4])
    plan = C.ordered.Plan(n, C.ordered.Method.UserProvided(algo), allow_large=n > 1024)
else:
    base_n, batch = int(sys.argv[4]), int(sys.argv[5])
    plan = C.unordered.Plan(n, C.unordered.Method.UserProvided(algo, base_n))
data = torch.view_as_complex(torch.rand(batch, n, 2, dtype=torch.float64, device="cuda")).contiguous()
for _ in range(2):
    plan.fwd(data)
    plan.inv(data)
    data.mul_(1.0 / n)
torch.cuda.synchronize()
plan.fwd(data)
plan.inv(data)
torch.cuda.synchronize()
print(plan.kernel_name())
