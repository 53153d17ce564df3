#!/usr/bin/env python3
"""Would replaying the multi-launch (L2-chunked, multi-stream) schedule from a CUDA graph help?  Captures
plan.fwd / plan.inv with torch.cuda.graph (the library's fork / join over its auxiliary streams is capturable)
and times replay against direct calls:  python tools/graph_probe.py <log2 n>"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch

import concrete_fft_b200 as C

logn = int(sys.argv[1])
n = 1 << logn
batch = (1 << 31) // (16 * n)
plan = C.unordered.Plan(n, C.unordered.Method.Measure())
data = torch.view_as_complex(torch.rand(batch, n, 2, dtype=torch.float64, device="cuda")).contiguous()
for _ in range(3):
    plan.fwd(data); plan.inv(data); data.mul_(1.0 / n)
torch.cuda.synchronize()


def timeit(fn, reps=10):
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(reps + 1)]
    t0 = time.perf_counter()
    ev[0].record()
    for i in range(reps):
        fn()
        ev[i + 1].record()
    host = (time.perf_counter() - t0) / reps
    torch.cuda.synchronize()
    ts = sorted(ev[i].elapsed_time(ev[i + 1]) for i in range(reps))
    return ts[len(ts) // 2], host * 1e3


s = torch.cuda.Stream()
with torch.cuda.stream(s):
    plan.fwd(data); plan.inv(data); data.mul_(1.0 / n)
    torch.cuda.synchronize()
    gf, gi = torch.cuda.CUDAGraph(), torch.cuda.CUDAGraph()
    with torch.cuda.graph(gf, stream=s):
        plan.fwd(data)
    with torch.cuda.graph(gi, stream=s):
        plan.inv(data)
    torch.cuda.synchronize()
    b = 2 * 16 * n * batch
    d_f, h_f = timeit(lambda: plan.fwd(data)); data.mul_(float(n) ** -10)
    d_i, h_i = timeit(lambda: plan.inv(data)); data.mul_(float(n) ** -10)
    g_f, gh_f = timeit(gf.replay); data.mul_(float(n) ** -10)
    g_i, gh_i = timeit(gi.replay)
print("n=2^%d %s  direct fwd %.3f ms (%.0f GB/s, host %.3f ms) inv %.3f ms (%.0f GB/s) | graph fwd %.3f ms (%.0f GB/s, host %.3f ms) inv %.3f ms (%.0f GB/s)" % (
    logn, plan.kernel_name(), d_f, b / d_f / 1e6, h_f, d_i, b / d_i / 1e6, g_f, b / g_f / 1e6, gh_f, g_i, b / g_i / 1e6))
