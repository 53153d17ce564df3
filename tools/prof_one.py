#!/usr/bin/env python3
"""Two warm-up rounds, then one fwd + inv launch of a plan, for ncu captures (-s <launches of the warm-up>):
    ncu --set full --clock-control none --import-source on -k regex:<kernel> -o gpurun_out/x python tools/prof_one.py f128 2048 4096
    python tools/prof_one.py c64|ordered|f128 <n> <batch>"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch

import concrete_fft_b200 as C

kind, n, batch = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
dev = torch.device("cuda", 0)
if kind == "f128":
    plan = C.fft128.Plan(n)
    planes = [torch.rand(batch, n, dtype=torch.float64, device=dev), torch.zeros(batch, n, dtype=torch.float64, device=dev),
              torch.rand(batch, n, dtype=torch.float64, device=dev), torch.zeros(batch, n, dtype=torch.float64, device=dev)]
    for _ in range(2):
        plan.fwd(*planes)
        plan.inv(*planes)
        for p in planes:
            p.mul_(1.0 / n)
    torch.cuda.synchronize()
    plan.fwd(*planes)
    plan.inv(*planes)
else:
    if kind == "ordered":
        plan = C.ordered.Plan(n, C.ordered.Method.UserProvided(C.ordered.FftAlgo.Dif16), allow_large=n > 1024)
    else:
        plan = C.unordered.Plan(n, C.unordered.Method.UserProvided(C.ordered.FftAlgo.Dif16, min(n, 256)))
        if len(sys.argv) > 4:
            plan.autotune()
    data = torch.view_as_complex(torch.rand(batch, n, 2, dtype=torch.float64, device=dev)).contiguous()
    for _ in range(2):
        plan.fwd(data)
        plan.inv(data)
        data.mul_(1.0 / n)
    torch.cuda.synchronize()
    plan.fwd(data)
    plan.inv(data)
torch.cuda.synchronize()
print(plan.kernel_name())
