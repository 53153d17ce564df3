#!/usr/bin/env python3
"""Time forced kernel variants of the unordered (Dif16, 256) plan against each other:
    python tools/cmp_variants.py <log2 n> <variant> [<variant> ...]     (variant 'auto' = Method::Measure)"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch

import concrete_fft_b200 as C

logn = int(sys.argv[1])
n = 1 << logn
batch = int(os.environ.get("CMP_BATCH", 0)) or (1 << 31) // (16 * n)
data = torch.view_as_complex(torch.rand(batch, n, 2, dtype=torch.float64, device="cuda")).contiguous()


def timeit(fn, reps=int(os.environ.get("CMP_REPS", 10))):
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(reps + 1)]
    ev[0].record()
    for i in range(reps):
        fn()
        ev[i + 1].record()
    torch.cuda.synchronize()
    ts = sorted(ev[i].elapsed_time(ev[i + 1]) for i in range(reps))
    return ts[len(ts) // 2]


for var in sys.argv[2:]:
    if var == "auto":
        plan = C.unordered.Plan(n, C.unordered.Method.Measure())
    else:
        os.environ["CFFT_B200_FAST_VARIANT"] = var
        plan = C.unordered.Plan(n, C.unordered.Method.UserProvided(C.ordered.FftAlgo.Dif16, 256))
        del os.environ["CFFT_B200_FAST_VARIANT"]
    for _ in range(3):
        plan.fwd(data); plan.inv(data); data.mul_(1.0 / n)
    f = timeit(lambda: plan.fwd(data)); data.mul_(float(n) ** -10)
    i = timeit(lambda: plan.inv(data)); data.mul_(float(n) ** -10)
    b = 2 * 16 * n * batch
    print("n=2^%d %-5s %-42s fwd %.3f ms %5.0f GB/s   inv %.3f ms %5.0f GB/s" % (logn, var, plan.kernel_name(), f, b / f / 1e6, i, b / i / 1e6), flush=True)
