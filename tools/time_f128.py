#!/usr/bin/env python3
"""Device-resident fft128 fwd / inv timing: python tools/time_f128.py <log2 n> [...]   (1 GiB of rows each)"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch

import concrete_fft_b200 as C


def timeit(fn, reps=10):
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(reps + 1)]
    ev[0].record()
    for i in range(reps):
        fn()
        ev[i + 1].record()
    torch.cuda.synchronize()
    ts = sorted(ev[i].elapsed_time(ev[i + 1]) for i in range(reps))
    return ts[len(ts) // 2]


for a in sys.argv[1:]:
    n = 1 << int(a)
    batch = (1 << 30) // (32 * n)
    planes = [torch.rand(batch, n, dtype=torch.float64, device="cuda") for _ in range(4)]
    planes[1].mul_(1e-17); planes[3].mul_(1e-17)
    plan = C.fft128.Plan(n)
    for _ in range(2):
        plan.fwd(*planes); plan.inv(*planes)
        for p in planes: p.mul_(1.0 / n)
    f = timeit(lambda: plan.fwd(*planes))
    for p in planes: p.mul_(float(n) ** -5)
    i = timeit(lambda: plan.inv(*planes))
    instr = 94.0 * (n / 2) * int(a) * batch
    print("fft128 n=2^%-2d batch %-7d fwd %.3f ms %.2f T instr/s   inv %.3f ms %.2f T instr/s" % (int(a), batch, f, instr / f / 1e9, i, instr / i / 1e9), flush=True)
    del planes
