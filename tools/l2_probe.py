#!/usr/bin/env python3
"""Does B200's L2 keep freshly written data for a following kernel?  Times repeated in-place
read-modify-write sweeps (x.add_(1)) over working sets of 4 MiB .. 1 GiB: if small sets stay
L2-resident the effective GB/s rises well above the HBM figure."""
import torch

for mb in [4, 8, 16, 32, 48, 64, 96, 128, 192, 256, 1024]:
    n = mb * (1 << 20) // 8
    x = torch.zeros(n, dtype=torch.float64, device="cuda")
    y = torch.zeros(n, dtype=torch.float64, device="cuda")
    reps = max(20, 4096 // mb)
    for _ in range(5):
        y.copy_(x); x.copy_(y)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(reps):
            y.copy_(x)   # pass A: read x, write y
            x.copy_(y)   # pass B: read y (just written), write x
    g.replay(); torch.cuda.synchronize()
    e0.record(); g.replay(); e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    print("working set 2 x %4d MiB: %.0f GB/s (read+write bytes / time)" % (mb, 4 * n * 8 * reps / ms / 1e6), flush=True)
