#!/usr/bin/env python3
"""Generic register kernel (exact-regs, c64_regs.cu) against the shared-memory tile kernel (exact-tile) on plans
without a specialised kernel: python tools/cmp_exact.py"""
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


cases = [("unordered", 2048, A.Dif16, 1024), ("unordered", 2048, A.Dif4, 32), ("unordered", 2048, A.Dit8, 512),
         ("unordered", 1024, A.Dif8, 1024), ("unordered", 4096, A.Dit16, 512), ("unordered", 16384, A.Dif16, 1024),
         ("unordered", 65536, A.Dif8, 64), ("ordered", 1024, A.Dit16, 0), ("ordered", 512, A.Dif8, 0), ("ordered", 128, A.Dif4, 0),
         ("ordered", 1024, A.Dif2, 0)]
for kind, n, algo, base_n in cases:
    batch = (1 << 30) // (16 * n)
    plan = (C.unordered.Plan(n, C.unordered.Method.UserProvided(algo, base_n)) if kind == "unordered"
            else C.ordered.Plan(n, C.ordered.Method.UserProvided(algo)))
    data = torch.view_as_complex(torch.rand(batch, n, 2, dtype=torch.float64, device="cuda")).contiguous()
    out = []
    for env in (None, "1"):
        if env:
            os.environ["CFFT_B200_EXACT_TILE"] = env
        for _ in range(2):
            plan.fwd(data); plan.inv(data); data.mul_(1.0 / n)
        f = timeit(lambda: plan.fwd(data)); data.mul_(float(n) ** -10)
        i = timeit(lambda: plan.inv(data)); data.mul_(float(n) ** -10)
        os.environ.pop("CFFT_B200_EXACT_TILE", None)
        b = 2 * 16 * n * batch
        out.append("%5.0f / %5.0f" % (b / f / 1e6, b / i / 1e6))
    print("%-9s n=%-6d %-5s base %-5s  regs fwd/inv GB/s %s   tile %s" % (kind, n, algo.name, base_n or n, out[0], out[1]), flush=True)
    del data
    torch.cuda.empty_cache()
