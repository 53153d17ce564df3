#!/usr/bin/env python3
"""Latency of the literal drop-in call (one polynomial per call, host memory) vs the CPU port."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
import concrete_fft_b200 as C
import oracle_lib as O

A = C.ordered.FftAlgo
for n in [1024, 2048, 8192]:
    plan = C.unordered.Plan(n, C.unordered.Method.UserProvided(A.Dif16, 256))
    for kind in ["pageable", "pinned"]:
        x = np.random.default_rng(0).random(n) + 0j
        if kind == "pinned":
            t = torch.from_numpy(x).pin_memory(); x = t.numpy()
        for _ in range(20):
            plan.fwd(x); x *= 1.0 / n
        reps = 2000
        t0 = time.perf_counter()
        for _ in range(reps):
            plan.fwd(x)
        el = (time.perf_counter() - t0) / reps
        print("n=%5d host call, %-8s memory: %.1f us per Plan.fwd (batch 1)" % (n, kind, el * 1e6))
    d = torch.zeros(n, dtype=torch.complex128, device="cuda")
    for _ in range(20):
        plan.fwd(d)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(2000):
        plan.fwd(d)
    torch.cuda.synchronize()
    print("n=%5d device call: %.1f us per Plan.fwd launch (batch 1, async, amortised)" % (n, (time.perf_counter() - t0) / 2000 * 1e6))
    ref = O.UnorderedPlan(n, O.DIF16, 256, fast=True)
    buf = np.random.default_rng(0).random((1, n)) + 0j
    t0 = time.perf_counter()
    for _ in range(2000):
        ref.fwd_inplace(buf, 1)
    print("n=%5d CPU port, 1 thread: %.1f us per transform (incl. ctypes call)" % (n, (time.perf_counter() - t0) / 2000 * 1e6))
