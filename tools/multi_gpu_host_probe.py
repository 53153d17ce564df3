#!/usr/bin/env python3
"""One host call sharded over the GPUs of the box inside the library (cfft_c64_host_multi) against the same call on one GPU:
fwd+inv of N = 2048 x 32768 polynomials (1 GiB each way) from pinned host memory."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch

import concrete_fft_b200 as C
from concrete_fft_b200.sharding import MultiGpu

n, batch = 2048, 32768
plan = C.unordered.Plan(n, C.unordered.Method.UserProvided(C.ordered.FftAlgo.Dif16, 256))
pinned = torch.empty((batch, n), dtype=torch.complex128).pin_memory()
buf = pinned.numpy()
buf[:] = 1.0
for devices in [[0], list(range(torch.cuda.device_count()))]:
    mg = MultiGpu(plan, devices)
    buf[:] = 1.0
    mg.fwd_inv(buf)
    t0 = time.perf_counter()
    reps = 3  # values grow by n per call: 2048^4 is far from overflow
    for _ in range(reps):
        mg.fwd_inv(buf)
    dt = (time.perf_counter() - t0) / reps
    print("devices %s: %.1f ms per fwd+inv of %d x %d, %.2f M transforms/s, %.1f GB/s each way" % (devices, dt * 1e3, batch, n, 2 * batch / dt / 1e6, batch * n * 16 / dt / 1e9), flush=True)
