#!/usr/bin/env python3
"""unordered (Dif16, 256) plan with a forced kernel variant, two warm-up rounds then one fwd + inv:
    python tools/prof_variant.py <variant> <n> <batch>"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch

import concrete_fft_b200 as C

var, n, batch = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
os.environ["CFFT_B200_FAST_VARIANT"] = var
plan = C.unordered.Plan(n, C.unordered.Method.UserProvided(C.ordered.FftAlgo.Dif16, 256))
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
