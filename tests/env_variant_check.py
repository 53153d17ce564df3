"""Helper of test_gpu_env_variants.py (runs in a subprocess because the library reads its tuning variables once per process):
fwd / inv of the (Dif16, 256) plans n = 256 .. 16384 against the oracle, bit for bit, under whatever CFFT_B200_* variables
the parent set.  Prints OK or raises."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import oracle_lib as O

import concrete_fft_b200 as C

rng = np.random.default_rng(31337)
A = C.ordered.FftAlgo
for n in (256, 512, 1024, 2048, 4096, 8192, 16384, 32768):
    variants = [None] + (["4"] if n in (8192, 16384) else []) + (["8"] if n == 32768 else [])
    for var in variants:
        if var:
            os.environ["CFFT_B200_FAST_VARIANT"] = var
        plan = C.unordered.Plan(n, C.unordered.Method.UserProvided(A.Dif16, 256))
        os.environ.pop("CFFT_B200_FAST_VARIANT", None)
        ref = O.UnorderedPlan(n, O.DIF16, 256)
        for batch in (1, 3, 700 if n <= 2048 else 150):  # more rows than one wave of resident CTAs: the prefetch distance is exercised
            x = rng.random((batch, n)) + 1j * rng.random((batch, n))
            d = torch.from_numpy(x.copy()).cuda()
            plan.fwd(d)
            torch.cuda.synchronize()
            want = ref.fwd(x, threads=8)
            assert np.array_equal(d.cpu().numpy().view(np.uint64), want.view(np.uint64)), ("fwd", n, var, batch, plan.kernel_name())
            plan.inv(d)
            torch.cuda.synchronize()
            assert np.array_equal(d.cpu().numpy().view(np.uint64), ref.inv(want, threads=8).view(np.uint64)), ("inv", n, var, batch)
# ordered (standard order) plans above the reference's cap: the DFT definition = the unordered reference plan un-permuted
for n in (2048, 4096, 8192, 16384, 65536):
    plan = C.ordered.Plan(n, C.ordered.Method.UserProvided(A.Dif16), allow_large=True)
    ref = O.UnorderedPlan(n, O.DIF16, 256)
    pi = O.permutation(n, 256)
    for batch in (1, 5) + ((300,) if n <= 8192 else ()):
        x = rng.random((batch, n)) + 1j * rng.random((batch, n))
        d = torch.from_numpy(x.copy()).cuda()
        plan.fwd(d)
        torch.cuda.synchronize()
        want = np.ascontiguousarray(ref.fwd(x, threads=8)[:, pi])
        assert np.array_equal(d.cpu().numpy().view(np.uint64), want.view(np.uint64)), ("ordered fwd", n, batch, plan.kernel_name())
        plan.inv(d)
        torch.cuda.synchronize()
        perm_in = np.empty_like(want)
        perm_in[:, pi] = want
        assert np.array_equal(d.cpu().numpy().view(np.uint64), ref.inv(perm_in, threads=8).view(np.uint64)), ("ordered inv", n, batch)
print("OK")
