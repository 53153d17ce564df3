#!/usr/bin/env python3
"""cfft_c64_fwd_mul_inv_multi (two outputs per row, forward transforms shared) against cfft_c64_fwd_mul_inv once per output:
    python tools/fused_mul2_probe.py [n ...]      (a = 1 GiB device resident, b shared by the batch)"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch

import concrete_fft_b200 as C

A = C.ordered.FftAlgo


def timed(fn, reps=10):
    fn()
    torch.cuda.synchronize()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(reps + 1)]
    ev[0].record()
    for i in range(reps):
        fn()
        ev[i + 1].record()
    torch.cuda.synchronize()
    ts = sorted(ev[i].elapsed_time(ev[i + 1]) for i in range(reps))
    return ts[len(ts) // 2]


for n in [int(x) for x in sys.argv[1:]] or [1024, 2048]:
    plan = C.unordered.Plan(n, C.unordered.Method.UserProvided(A.Dif16, 256))
    for k in (2, 4, 6):
        rows = (1 << 30) // (16 * n * k)
        a = torch.view_as_complex(torch.rand(rows, k, n, 2, dtype=torch.float64, device="cuda") - 0.5).contiguous()
        b = torch.view_as_complex(torch.rand(k, 2, n, 2, dtype=torch.float64, device="cuda") - 0.5).contiguous()
        b0, b1 = b[:, 0].contiguous(), b[:, 1].contiguous()
        out = torch.empty((rows, 2, n), dtype=torch.complex128, device="cuda")
        o0 = torch.empty((rows, n), dtype=torch.complex128, device="cuda")
        o1 = torch.empty((rows, n), dtype=torch.complex128, device="cuda")
        tm = timed(lambda: plan.fwd_mul_inv_multi(a, b, out))
        ts = timed(lambda: (plan.fwd_mul_inv(a, b0, o0), plan.fwd_mul_inv(a, b1, o1)))
        print("n=%d k=%d rows=%d: two outputs in one kernel %.3f ms (%.2f M external products/s), one call per output %.3f ms, x %.2f"
              % (n, k, rows, tm, rows / tm / 1e3, ts, ts / tm), flush=True)
        del a, b, out, o0, o1
