#!/usr/bin/env python3
"""Host-side issue time vs device time per call (is a multi-launch schedule launch-bound?):
    python tools/issue_time.py <c64|ordered> <log2 n> [bytes]"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch

import concrete_fft_b200 as C

kind, logn = sys.argv[1], int(sys.argv[2])
n = 1 << logn
batch = (int(sys.argv[3]) if len(sys.argv) > 3 else 1 << 31) // (16 * n)
if kind == "ordered":
    plan = C.ordered.Plan(n, C.ordered.Method.Measure(), allow_large=n > 1024)
else:
    plan = C.unordered.Plan(n, C.unordered.Method.Measure())
data = torch.view_as_complex(torch.rand(batch, n, 2, dtype=torch.float64, device="cuda")).contiguous()
for _ in range(3):
    plan.fwd(data); plan.inv(data); data.mul_(1.0 / n)
torch.cuda.synchronize()
reps = 10
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
t0 = time.perf_counter()
e0.record()
for _ in range(reps):
    plan.fwd(data)
e1.record()
t_issue = time.perf_counter() - t0
torch.cuda.synchronize()
print("%s n=2^%d batch %d kernel %s: host issue %.3f ms/call, device %.3f ms/call" % (
    kind, logn, batch, plan.kernel_name(), 1e3 * t_issue / reps, e0.elapsed_time(e1) / reps))
