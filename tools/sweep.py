#!/usr/bin/env python3
"""Size sweep (BASELINE.json configs[4]): device-resident fwd / inv timing per size, one JSON
line per (workload, n).  Batch = 2 GiB / bytes-per-transform (capped), inputs larger than L2.

    python tools/sweep.py [--workload c64|f128|both] [--min 4] [--max 20] [--out gpurun_out/sweep.jsonl]
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def time_launches(torch, fn, reps):
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(reps + 1)]
    ev[0].record()
    for i in range(reps):
        fn()
        ev[i + 1].record()
    torch.cuda.synchronize()
    ts = sorted(ev[i].elapsed_time(ev[i + 1]) for i in range(reps))
    return ts[len(ts) // 2], ts[0]


def cpu_rate(kind, n, algo_name, base_n, seconds=0.6):
    """forward transforms/s of the oracle port (-O3 build) on all host threads; test infrastructure
    used as the reported CPU baseline only."""
    import time

    import numpy as np

    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle_lib as O

    try:
        threads = len(os.sched_getaffinity(0))
    except Exception:
        threads = os.cpu_count() or 1
    rng = np.random.default_rng(0)
    rows = max(threads * 4, min(4096, (1 << 22) // n))
    if kind == "f128":
        plan = O.F128Plan(n, fast=True)
        rows = max(threads, min(512, (1 << 19) // n))
        planes = [rng.random((rows, n)), np.zeros((rows, n)), rng.random((rows, n)), np.zeros((rows, n))]
        step = lambda: plan.fwd_inplace(planes, O.F128_FMA, threads)
        rescale = lambda: [p.__imul__(1.0 / n) for p in planes]
    else:
        plan = O.UnorderedPlan(n, O.ALGO_NAMES.index(algo_name), base_n, fast=True)
        buf = rng.random((rows, n)) + 1j * rng.random((rows, n))
        step = lambda: plan.fwd_inplace(buf, threads)
        rescale = lambda: buf.__imul__(1.0 / n)
    step(); rescale()
    t0, done = time.perf_counter(), 0
    while time.perf_counter() - t0 < seconds:
        step(); rescale()
        done += 1
    return rows * done / (time.perf_counter() - t0), threads


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="both")
    ap.add_argument("--min", type=int, default=4)
    ap.add_argument("--max", type=int, default=20)
    ap.add_argument("--reps", type=int, default=10)
    ap.add_argument("--bytes", type=int, default=1 << 31)
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "sweep.jsonl"))
    ap.add_argument("--cufft", action="store_true", help="also time torch.fft.fft (cuFFT) as a reference point")
    ap.add_argument("--cpu", action="store_true", help="also time the CPU port of the reference algorithm (oracle/, all host threads)")
    ap.add_argument("--table", default="", help="write a markdown table of the results to this path")
    args = ap.parse_args()
    import torch

    import concrete_fft_b200 as C

    peak = 6650.0
    try:
        peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
    except Exception:
        pass
    os.makedirs(os.path.dirname(args.out), exist_ok=True)
    out = open(args.out, "a")
    rows_out = []
    dev = torch.device("cuda", 0)
    g = torch.Generator(device=dev).manual_seed(1)
    if args.workload in ("c64", "both"):
        for logn in range(args.min, args.max + 1):
            n = 1 << logn
            batch = max(1, args.bytes // (16 * n))
            plan = C.unordered.Plan(n, C.unordered.Method.Measure())
            algo, base_n = plan.algo()
            data = torch.view_as_complex(torch.rand(batch, n, 2, dtype=torch.float64, device=dev, generator=g)).contiguous()
            for _ in range(3):
                plan.fwd(data); plan.inv(data); data.mul_(1.0 / n)
            fwd_ms, fwd_best = time_launches(torch, lambda: plan.fwd(data), args.reps)
            data.mul_(float(n) ** -args.reps)
            inv_ms, inv_best = time_launches(torch, lambda: plan.inv(data), args.reps)
            bytes_ = 2 * 16 * n * batch
            # cuFFT (through torch.fft) as a GPU reference point only -- never on the product path
            cufft_ms = None
            if args.cufft and n >= 2:
                try:
                    outb = torch.fft.fft(data, dim=1)
                    cufft_ms, _ = time_launches(torch, lambda: torch.fft.fft(data, dim=1, out=outb), 5)
                    del outb
                except Exception:
                    cufft_ms = None
            rec = {"workload": "c64-unordered", "n": n, "batch": batch, "plan": "%s/%d" % (algo.name, base_n),
                   "cufft_z2z_out_of_place_ms": cufft_ms, "cufft_gbs": (2 * 16 * n * batch / cufft_ms / 1e6) if cufft_ms else None,
                   "kernel": plan.kernel_name(), "fwd_ms": fwd_ms, "inv_ms": inv_ms,
                   "fwd_gbs": bytes_ / fwd_ms / 1e6, "inv_gbs": bytes_ / inv_ms / 1e6,
                   "frac_of_measured_hbm": bytes_ / fwd_ms / 1e6 / peak,
                   "fwd_transforms_per_s": batch / fwd_ms * 1e3,
                   "fwd_gflops_5nlog2n": 5.0 * n * logn * batch / fwd_ms / 1e6}
            if args.cpu and n >= 2:
                rec["cpu_port_fwd_transforms_per_s"], rec["cpu_threads"] = cpu_rate("c64", n, algo.name, base_n)
            rows_out.append(rec)
            print(json.dumps(rec), flush=True)
            out.write(json.dumps(rec) + "\n")
            del data, plan
            torch.cuda.empty_cache()
    if args.workload in ("ordered", "both"):
        for logn in range(max(args.min, 1), args.max + 1):
            n = 1 << logn
            batch = max(1, args.bytes // (16 * n))
            if n > 1024:
                batch = max(1, batch // 2)  # leave room for the out-of-place workspace
            plan = C.ordered.Plan(n, C.ordered.Method.Measure(), allow_large=n > 1024)
            data = torch.view_as_complex(torch.rand(batch, n, 2, dtype=torch.float64, device=dev, generator=g)).contiguous()
            for _ in range(3):
                plan.fwd(data); plan.inv(data); data.mul_(1.0 / n)
            fwd_ms, _ = time_launches(torch, lambda: plan.fwd(data), args.reps)
            data.mul_(float(n) ** -args.reps)
            inv_ms, _ = time_launches(torch, lambda: plan.inv(data), args.reps)
            bytes_ = 2 * 16 * n * batch
            rec = {"workload": "c64-ordered", "n": n, "batch": batch, "plan": plan.algo().name, "kernel": plan.kernel_name(),
                   "fwd_ms": fwd_ms, "inv_ms": inv_ms, "fwd_gbs": bytes_ / fwd_ms / 1e6, "inv_gbs": bytes_ / inv_ms / 1e6,
                   "frac_of_measured_hbm": bytes_ / fwd_ms / 1e6 / peak, "fwd_transforms_per_s": batch / fwd_ms * 1e3}
            rows_out.append(rec)
            print(json.dumps(rec), flush=True)
            out.write(json.dumps(rec) + "\n")
            del data, plan
            torch.cuda.empty_cache()
    if args.workload in ("f128", "both"):
        for logn in range(max(args.min, 5), min(args.max, 16) + 1):
            n = 1 << logn
            batch = max(1, (args.bytes // 2) // (32 * n))
            plan = C.fft128.Plan(n)
            planes = [torch.rand(batch, n, dtype=torch.float64, device=dev, generator=g), torch.zeros(batch, n, dtype=torch.float64, device=dev),
                      torch.rand(batch, n, dtype=torch.float64, device=dev, generator=g), torch.zeros(batch, n, dtype=torch.float64, device=dev)]
            for _ in range(3):
                plan.fwd(*planes); plan.inv(*planes)
                for p in planes:
                    p.mul_(1.0 / n)
            fwd_ms, _ = time_launches(torch, lambda: plan.fwd(*planes), 5)
            for p in planes:
                p.mul_(float(n) ** -5)
            inv_ms, _ = time_launches(torch, lambda: plan.inv(*planes), 5)
            instr = 94.0 * (n // 2) * logn
            rec = {"workload": "fft128", "n": n, "batch": batch, "kernel": plan.kernel_name(), "fwd_ms": fwd_ms, "inv_ms": inv_ms,
                   "fwd_transforms_per_s": batch / fwd_ms * 1e3, "inv_transforms_per_s": batch / inv_ms * 1e3,
                   "fwd_fp64_pipe_frac_at_1965MHz": instr * batch / (fwd_ms * 1e-3) / (64 * 148 * 1965e6),
                   "inv_fp64_pipe_frac_at_1965MHz": instr * batch / (inv_ms * 1e-3) / (64 * 148 * 1965e6),
                   "fwd_gbs": 2 * 32 * n * batch / fwd_ms / 1e6}
            if args.cpu:
                rec["cpu_port_fwd_transforms_per_s"], rec["cpu_threads"] = cpu_rate("f128", n, "", n)
            rows_out.append(rec)
            print(json.dumps(rec), flush=True)
            out.write(json.dumps(rec) + "\n")
            del planes, plan
            torch.cuda.empty_cache()


    if args.table:
        with open(args.table, "w") as f:
            f.write("| workload | n | plan / kernel | fwd transforms/s | fwd GB/s (algorithmic) | fraction of roofline | inv GB/s | GFLOP/s (5 n log2 n) | cuFFT GB/s | CPU port transforms/s (threads) | GPU / CPU |\n")
            f.write("|---|---|---|---|---|---|---|---|---|---|---|\n")
            for r in rows_out:
                cpu = r.get("cpu_port_fwd_transforms_per_s")
                if r["workload"] == "fft128":
                    frac = "%.2f of FP64 issue" % r["fwd_fp64_pipe_frac_at_1965MHz"]
                    plan = r["kernel"]
                else:
                    frac = "%.2f of measured HBM" % r["frac_of_measured_hbm"]
                    plan = "%s / %s" % (r.get("plan", ""), r["kernel"])
                f.write("| %s | %d | %s | %.3g | %.0f | %s | %s | %s | %s | %s | %s |\n" % (
                    r["workload"], r["n"], plan, r["fwd_transforms_per_s"], r["fwd_gbs"], frac,
                    ("%.0f" % r["inv_gbs"]) if "inv_gbs" in r else "-",
                    ("%.0f" % r["fwd_gflops_5nlog2n"]) if "fwd_gflops_5nlog2n" in r else "-",
                    ("%.0f" % r["cufft_gbs"]) if r.get("cufft_gbs") else "-",
                    ("%.3g (%d)" % (cpu, r["cpu_threads"])) if cpu else "-",
                    ("%.0fx" % (r["fwd_transforms_per_s"] / cpu)) if cpu else "-"))


if __name__ == "__main__":
    main()
