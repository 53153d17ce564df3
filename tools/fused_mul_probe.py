#!/usr/bin/env python3
"""cfft_c64_fwd_mul_inv (one kernel) against the same work as separate library launches (fwd per term, point-wise
product / multiply-accumulate, inv), device resident:  python tools/fused_mul_probe.py [n ...]
One JSON line per (n, k_terms, b shared / per row): rows/s of both paths, HBM bytes each moves per row, GB/s."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch

import concrete_fft_b200 as C

A = C.ordered.FftAlgo


def timed(fn, reps):
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


if len(sys.argv) > 1 and sys.argv[1] == "f128":
    # cfft_f128_fwd_mul_inv against cfft_f128_fwd + cfft_f128_cplx_mul_scale + cfft_f128_inv
    for n in [int(s) for s in sys.argv[2:]] or [1024, 2048, 4096]:
        plan = C.fft128.Plan(n)
        rows = (1 << 30) // (32 * n)
        for shared in (True, False):
            L = [torch.rand(rows, n, dtype=torch.float64, device="cuda") * (1.0 if i % 2 == 0 else 1e-17) for i in range(4)]
            R = [torch.rand(*((n,) if shared else (rows, n)), dtype=torch.float64, device="cuda") * (1.0 if i % 2 == 0 else 1e-17) for i in range(4)]
            Rfull = R if not shared else [r.expand(rows, n).contiguous() for r in R]
            f = 0.5 / n  # keeps |lhs| bounded over repetitions (operands in [0, 1): product of n terms < n)

            def separate():
                plan.fwd(*L)
                C.fft128.cplx_mul_scale(L, Rfull, f)
                plan.inv(*L)

            t_fused = timed(lambda: plan.fwd_mul_inv(L, R, f), 10)
            for t in L:
                t.uniform_(0, 1)
            t_sep = timed(separate, 10)
            print(json.dumps({"kind": "f128", "n": n, "rows": rows, "rhs": "shared" if shared else "per-row",
                              "fused_ms": round(t_fused, 4), "separate_ms": round(t_sep, 4), "speedup": round(t_sep / t_fused, 3),
                              "fused_products_per_s": round(rows / t_fused * 1e3),
                              "fp64_issue_frac": round(rows * (n // 2) * (n.bit_length() - 1) * 94 * 2 / (t_fused * 1e-3) / (64 * 148 * 1.965e9), 3)}),
                  flush=True)
            del L, R, Rfull
            torch.cuda.empty_cache()
    sys.exit(0)

sizes = [int(s) for s in sys.argv[1:]] or [512, 1024, 2048, 4096]
for n in sizes:
    plan = C.unordered.Plan(n, C.unordered.Method.UserProvided(A.Dif16, 256))
    for k, shared in [(1, False), (1, True), (2, True), (4, True), (6, True), (4, False)]:
        rows = (1 << 30) // (16 * n * k)  # a = 1 GiB
        a = torch.view_as_complex(torch.rand(rows, k, n, 2, dtype=torch.float64, device="cuda") - 0.5)
        b = torch.view_as_complex(torch.rand(*((k, n, 2) if shared else (rows, k, n, 2)), dtype=torch.float64, device="cuda") - 0.5)
        out = torch.empty(rows, n, dtype=torch.complex128, device="cuda")
        t_fused = timed(lambda: plan.fwd_mul_inv(a, b, out=out), 10)

        # separate launches on buffers laid out for them (term-major so every point-wise call is contiguous)
        fa = a.transpose(0, 1).contiguous()  # [k][rows][n]
        bb = (b.unsqueeze(1).expand(k, rows, n) if shared else b.transpose(0, 1)).contiguous()
        acc = torch.empty(rows, n, dtype=torch.complex128, device="cuda")
        scale = torch.tensor(1.0 / n, dtype=torch.float64, device="cuda")

        def separate():
            plan.fwd(fa)
            acc.copy_(fa[0])
            C.pointwise.mul_assign(acc, bb[0])
            for j in range(1, k):
                C.pointwise.mul_add_assign(acc, fa[j], bb[j])
            plan.inv(acc)
            fa.mul_(scale)  # keep magnitudes bounded across repetitions (not part of either path's count below)

        t_sep = timed(separate, 5)
        t_rescale = timed(lambda: fa.mul_(scale), 5)
        t_copy = timed(lambda: acc.copy_(fa[0]), 5)
        t_sep -= t_rescale + t_copy  # a caller would multiply in place; charge neither helper to the separate path
        bytes_fused = 16 * n * ((k if shared else 2 * k) + 1)  # b shared: L2-resident, not HBM traffic
        bytes_sep = 16 * n * (2 * k + 3 + 4 * (k - 1) + 2)
        print(json.dumps({
            "n": n, "k_terms": k, "b": "shared" if shared else "per-row", "rows": rows,
            "fused_ms": round(t_fused, 4), "separate_ms": round(t_sep, 4), "speedup": round(t_sep / t_fused, 3),
            "fused_rows_per_s": round(rows / t_fused * 1e3), "separate_rows_per_s": round(rows / t_sep * 1e3),
            "fused_hbm_bytes_per_row": bytes_fused, "separate_hbm_bytes_per_row": bytes_sep,
            "fused_GBps": round(bytes_fused * rows / t_fused / 1e6, 1),
            "fused_transforms_per_s": round(rows * (k + 1) / t_fused * 1e3),
        }), flush=True)
        del a, b, out, fa, bb, acc
        torch.cuda.empty_cache()
