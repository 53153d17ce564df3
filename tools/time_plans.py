#!/usr/bin/env python3
"""Device-resident fwd / inv timing of arbitrary unordered plans, one line per plan:
    python tools/time_plans.py 2048:Dif16:1024 2048:Dif8:512 ...      (2 GiB of rows each; CFFT_B200_* variables apply)"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch

import concrete_fft_b200 as C

A = C.ordered.FftAlgo


def timeit(fn, reps=10):
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(reps + 1)]
    ev[0].record()
    for i in range(reps):
        fn()
        ev[i + 1].record()
    torch.cuda.synchronize()
    ts = sorted(ev[i].elapsed_time(ev[i + 1]) for i in range(reps))
    return ts[len(ts) // 2]


for spec in sys.argv[1:]:
    n, algo, base_n = spec.split(":")  # base_n "ord": the ordered (standard order) plan of that algorithm
    n = int(n)
    batch = int(os.environ.get("CMP_BATCH", 0)) or (1 << 31) // (16 * n)
    data = torch.view_as_complex(torch.rand(batch, n, 2, dtype=torch.float64, device="cuda")).contiguous()
    if base_n == "ord":
        plan = C.ordered.Plan(n, C.ordered.Method.UserProvided(A[algo]), allow_large=n > 1024)
        base_n = n
    else:
        base_n = int(base_n)
        plan = C.unordered.Plan(n, C.unordered.Method.UserProvided(A[algo], base_n))
    for _ in range(3):
        plan.fwd(data); plan.inv(data); data.mul_(1.0 / n)
    f = timeit(lambda: plan.fwd(data)); data.mul_(float(n) ** -10)
    i = timeit(lambda: plan.inv(data)); data.mul_(float(n) ** -10)
    b = 2 * 16 * n * batch
    print("n=%-6d %-6s base %-5d %-20s fwd %.3f ms %5.0f GB/s   inv %.3f ms %5.0f GB/s" % (n, algo, base_n, plan.kernel_name(), f, b / f / 1e6, i, b / i / 1e6), flush=True)
    del data
