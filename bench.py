#!/usr/bin/env python3
"""bench.py -- headline benchmark of the batched c64 FFT hot path (BASELINE.json configs[1]) plus short legs
for the other BASELINE configs.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
                  [--workload c64|ordered|f128] [--n N] [--batch B] [--no-extra]

Step   = one pass of the hot path over one batch: fwd THEN inv of `batch` polynomials of size n, in place
         (2 * batch transforms).
value  = transforms/s, whole job (all ranks), inputs resident in HBM, CUDA-event timed, max over ranks.
         Inputs (2 GiB per GPU at the default size) are far larger than the 126 MB L2, so no flush is
         needed between iterations.  value_sustained = the same over a >= 2 s timed region (the K-step
         region of the default run is a ~0.25 s burst; a B200 settles to a lower clock under power cap).
e2e    = the same step through the public host-memory call (Plan.fwd_inv_host -> cfft_c64_fwd_inv_host):
         pinned host buffers, H2D + fwd + inv + D2H inside the timed region.
roofline    = achieved HBM GB/s of the dominant kernel (algorithmic bytes 2 * 16 * n per transform per
              launch / average launch time from CUDA events on the launch stream) against
              MEASURED_PEAKS.json; fft128: FP64 instructions/s against the rate cfft_probe_fp64_issue_rate
              measures on this very GPU.
cpu_baseline= the oracle port of the reference algorithm (-O3 build, bit-identical results) on the host
              cores, bounded sample of the same workload (rank 0, N = 1 only).
extra  = (default workload only) short legs for BASELINE.json configs[3] (fft128 n = 2048 x 16384) and
         configs[2] (ordered N = 2^16 x 4096 in total), each with value / roofline / e2e / cpu_baseline;
         configs[0] (unordered N = 1024, one polynomial, one host thread, in cache: the `cargo bench` shape);
         poly_mul: the fused integer negacyclic product step, device-resident and from host memory.
parity = every rank pushes the reference's golden vector through its own GPU (bit-exact against
         tests/golden, no oracle involved) and ranks compare a checksum of the headline plan's output on
         identical rows, outside the timed region.
--impl reference times the CPU port with all host threads on the same metric/config (the Rust reference
itself cannot be built in this image: no cargo/rustc).
"""
import argparse
import hashlib
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

FALLBACK_HBM_GBS = 6650.0  # /opt/skills/guides/B200_PROFILING.md fallback


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c64", choices=["c64", "ordered", "f128"])
    ap.add_argument("--n", type=int, default=0)
    ap.add_argument("--batch", type=int, default=0)
    ap.add_argument("--base-n", type=int, default=0)
    ap.add_argument("--algo", default="")
    ap.add_argument("--e2e-steps", type=int, default=3)
    ap.add_argument("--cpu-seconds", type=float, default=12.0)
    ap.add_argument("--sustained-seconds", type=float, default=2.0)
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-extra", action="store_true", help="skip the fft128 / ordered / N=1024 legs of the default run")
    ap.add_argument("--dd-full", action="store_true",
                    help="fft128: full-width double-double inputs (lo ~ U(-1/2, 1/2) ulp(hi)) instead of lo = 0 (SURVEY 8d config 4, second run)")
    return ap.parse_args()


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        try:
            d = json.load(open(path))
            return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs, burst copy)", d
        except Exception:
            pass
    return FALLBACK_HBM_GBS, "fallback (B200_PROFILING.md)", {}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the timed region."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits", "-lms", "50"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons, power = [], [], set(), []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            if len(r) < 7:
                continue
            try:
                sm.append(float(r[0])); mx.append(float(r[1])); power.append(float(r[2]))
            except ValueError:
                continue
            for name, v in zip(names, r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(power) if power else None, "samples": len(sm), "reasons": sorted(reasons)}


def bind_to_gpu_numa_node(index):
    """One process per GPU: run on (and first-touch the pinned host buffers from) the CPUs NVML reports as
    local to this GPU, so the e2e copies do not all cross one socket's memory controller.  No-op when NVML
    or the cpuset does not allow it."""
    try:
        import pynvml

        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(index)
        ncpu = os.cpu_count() or 1
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (ncpu + 63) // 64)
        ideal = {64 * w + b for w, word in enumerate(words) for b in range(64) if (word >> b) & 1}
        allowed = os.sched_getaffinity(0)
        both = ideal & allowed
        if both and both != allowed:
            os.sched_setaffinity(0, both)
            return "bound to %d of %d allowed CPUs (NVML ideal set for GPU %d)" % (len(both), len(allowed), index)
        return "not bound: NVML ideal CPUs %s the allowed set (%d CPUs)" % ("cover" if both else "do not intersect", len(allowed))
    except Exception as e:  # noqa: BLE001
        return "not bound: %s" % type(e).__name__


def workload_defaults(workload, n=0, batch=0, base_n=0, algo=""):
    if workload == "ordered":
        # BASELINE.json configs[2]: standard-order N = 2^16, batch 4096 in total, sharded across the GPUs
        n = n or 65536
        world = int(os.environ.get("WORLD_SIZE", "1"))
        batch = batch or max(1, 4096 // world)
        return n, batch, 256, "Dif16"
    if workload == "c64":
        n = n or 2048
        batch = batch or 65536
        base_n = base_n or min(n, 256)
        algo = algo or "Dif16"
    else:
        n = n or 2048
        batch = batch or 16384
        base_n, algo = n, ""
    return n, batch, base_n, algo


def host_threads():
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


def cpu_workload(workload, n, base_n, algo, rows, threads):
    """(step, rescale) closures of the oracle port (fast build) over `rows` polynomials in host memory:
    step = fwd pass then inv pass over all rows on `threads` threads (the GPU step's shape)."""
    import numpy as np
    import oracle_lib as O

    rng = np.random.default_rng(0)
    if workload in ("c64", "ordered"):
        plan = O.UnorderedPlan(n, O.ALGO_NAMES.index(algo), base_n, fast=True)
        buf = np.empty((rows, n), np.complex128)
        rng.random(out=buf.view(np.float64).reshape(rows, 2 * n))

        def step():
            plan.fwd_inplace(buf, threads)
            plan.inv_inplace(buf, threads)

        def rescale(k):  # fwd+inv multiplies by n: one exact power-of-two rescale every k steps, as in the GPU arm
            np.multiply(buf, float(n) ** -k, out=buf)
    else:
        plan = O.F128Plan(n, fast=True)
        planes = [rng.random((rows, n)), np.zeros((rows, n)), rng.random((rows, n)), np.zeros((rows, n))]

        def step():
            plan.fwd_inplace(planes, O.F128_FMA, threads)
            plan.inv_inplace(planes, O.F128_FMA, threads)

        def rescale(k):
            for p in planes:
                p *= float(n) ** -k
    return step, rescale


def cpu_sample_rows(workload, n, batch, max_bytes):
    """rows of the CPU sample: the whole per-GPU batch when it fits `max_bytes`, else the largest
    power-of-two slice that does -- always far larger than the host's last-level cache, so the CPU
    streams from DRAM like the real workload (a cache-resident sample would flatter it by ~1.4x)."""
    per = n * (32 if workload == "f128" else 16)
    rows = batch
    while rows > 1 and rows * per > max_bytes:
        rows //= 2
    return max(1, rows)


def cpu_port_rate(workload, n, batch, base_n, algo, threads, seconds):
    """transforms/s of the oracle port on `threads` host threads: fwd+inv steps over a DRAM-sized
    sample (<= 512 MiB) of the workload until ~`seconds` of wall time."""
    rows = cpu_sample_rows(workload, n, batch, 512 << 20)
    step, rescale = cpu_workload(workload, n, base_n, algo, rows, threads)
    every = max(1, 900 // max(1, n.bit_length() - 1))
    step()
    rescale(1)
    t0 = time.perf_counter()
    steps = 0
    while True:
        step()
        steps += 1
        if steps % every == 0:
            rescale(every)
        el = time.perf_counter() - t0
        if el >= seconds or steps >= 100000:
            break
    return 2.0 * rows * steps / el, rows, steps, el


def cpu_single_polynomial(seconds=1.0):
    """BASELINE.json configs[0]: unordered fwd+inv, N = 1024, ONE polynomial, ONE host thread, in cache -- the shape
    `cargo bench fft` measures (benches/fft.rs:164-185).  The reference's plan there is Measure(10 ms): base_n in
    {512, 1024} and the fastest of the eight algorithms (src/unordered.rs:568-630), so the port is timed on those
    candidates and on this library's own plan family, and the fastest is reported the way Measure would keep it."""
    import numpy as np
    import oracle_lib as O

    n = 1024
    rng = np.random.default_rng(0)
    x = rng.random(n) + 1j * rng.random(n)
    results = {}
    for algo, base_n in [("Dif16", 1024), ("Dit16", 1024), ("Dif8", 512), ("Dif4", 1024), ("Dif16", 256)]:
        plan = O.UnorderedPlan(n, O.ALGO_NAMES.index(algo), base_n, fast=True)
        buf = x.copy()
        reps, el = 0, 0.0
        t0 = time.perf_counter()
        while el < seconds / 5:
            for _ in range(64):
                plan.fwd_inplace(buf, 1)
                plan.inv_inplace(buf, 1)
                buf *= 1.0 / n
            reps += 64
            el = time.perf_counter() - t0
        results["%s/%d" % (algo, base_n)] = 1e6 * el / reps
    best = min(results, key=results.get)
    return {"workload": "unordered c64 fwd+inv N=1024, one polynomial, one host thread, in cache (BASELINE.json configs[0])",
            "us_per_fwd_inv": results[best], "plan": best, "transforms_per_s": 2e6 / results[best], "cores": 1, "kind": "port",
            "us_per_fwd_inv_by_plan": results,
            "note": "includes one 16 KiB rescale pass and two ctypes calls per iteration (~1 us); the Rust crate cannot be built here"}


METRIC_NAME = {"c64": "c64", "ordered": "standard-order c64", "f128": "fft128"}


def config_dict(workload, n, batch, base_n, algo, gpus):
    if workload == "ordered":
        return {"workload": "ordered (standard order in/out) c64 fwd+inv N=%d, batch %d in total = %d per GPU (BASELINE.json configs[2])" % (n, batch * gpus, batch),
                "n": n, "batch_per_gpu": batch, "plan": "ordered, allow_large (extension: the reference caps ordered plans at 2^10, src/ordered.rs:244); "
                "CPU arm = the reference's unordered plan {Dif16, 256} (no ordered reference exists at this size)",
                "sharding": "batch split across %d GPU(s), no collective" % gpus,
                "l2": "inputs (%.2f GiB per GPU) larger than L2, no flush" % (batch * n * 16 / 2**30)}
    if workload == "c64":
        return {"workload": "unordered c64 fwd+inv N=%d batch %d per GPU (BASELINE.json configs[1], TFHE bootstrapping shape)" % (n, batch),
                "n": n, "batch_per_gpu": batch, "plan": "UserProvided{base_algo: %s, base_n: %d}" % (algo, base_n),
                "sharding": "batch split across %d GPU(s), no collective" % gpus,
                "l2": "inputs (%.2f GiB per GPU) larger than L2, no flush" % (batch * n * 16 / 2**30)}
    return {"workload": "fft128 negacyclic fwd+inv n=%d batch %d per GPU (BASELINE.json configs[3])" % (n, batch),
            "n": n, "batch_per_gpu": batch, "sharding": "batch split across %d GPU(s), no collective" % gpus,
            "l2": "inputs (%.2f GiB per GPU) larger than L2, no flush" % (batch * n * 32 / 2**30)}


def reference_leg(workload, n, batch, base_n, algo, gpus, steps, warmup, budget_s):
    """The CPU port on all host threads over `steps` fwd+inv steps: the JSON line of `--impl reference`."""
    threads = host_threads()
    # Each step = fwd+inv over the WHOLE per-GPU batch when `steps` of them fit the time budget
    # (the default c64 workload does: 200 x 2 GiB); otherwise over the largest power-of-two slice that does.
    probe_rows = cpu_sample_rows(workload, n, batch, 64 << 20)
    step, rescale = cpu_workload(workload, n, base_n, algo, probe_rows, threads)
    step()
    t0 = time.perf_counter()
    step()
    per_row = (time.perf_counter() - t0) / probe_rows
    del step, rescale
    rows = batch
    while rows > probe_rows and rows * per_row * (steps + warmup) > budget_s:
        rows //= 2
    rows = max(rows, probe_rows)
    step, rescale = cpu_workload(workload, n, base_n, algo, rows, threads)
    every = max(1, 900 // max(1, n.bit_length() - 1))
    for _ in range(warmup):
        step()
        rescale(1)
    t0 = time.perf_counter()
    for i in range(steps):
        step()
        if (i + 1) % every == 0:
            rescale(every)
    el = time.perf_counter() - t0
    value = 2.0 * rows * steps / el
    bytes_per = n * (32 if workload == "f128" else 16)
    sample = "%d of %d polynomials per step (%.0f MiB, streamed from host DRAM), fwd+inv, %d host threads" % (
        rows, batch, rows * bytes_per / 2 ** 20, threads)
    return {
        "impl": "reference",
        "metric": "batched %s FFT transforms/s (fwd+inv)" % METRIC_NAME[workload],
        "value": value, "unit": "transforms/s", "n_gpus": gpus, "steps": steps, "warmup": warmup,
        "ms_per_step": 1e3 * el / steps, "higher_is_better": True,
        "scaling": "strong" if workload == "ordered" else "weak", "vs_baseline": None,
        "dtype": "f64x2 (double-double)" if workload == "f128" else "f64", "data": "synthetic",
        "config": config_dict(workload, n, batch, base_n, algo, gpus),
        "cpu_baseline": {"value": value, "unit": "transforms/s", "cores": threads, "kind": "port", "sample": sample,
                         "note": "C restatement of the reference algorithm (oracle/, -O3 AVX2+FMA build: two complex per "
                                 "register like the reference's c64x2 path; bit-identical to the reference's golden "
                                 "vector; persistent thread pool); the Rust crate cannot be built here"},
        "e2e": {"value": value, "unit": "transforms/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import oracle_lib as O

    O.build()
    n, batch, base_n, algo = workload_defaults(args.workload, args.n, args.batch, args.base_n, args.algo)
    line = reference_leg(args.workload, n, batch, base_n, algo, args.gpus, args.steps, args.warmup, 75.0)
    default_run = args.workload == "c64" and not (args.n or args.batch or args.base_n or args.algo)
    if default_run and not args.no_extra:
        extra = {}
        for wl in ("f128", "ordered"):
            en, eb, ebn, ealgo = workload_defaults(wl)
            sub = reference_leg(wl, en, eb, ebn, ealgo, args.gpus, 3, 1, 12.0)
            extra[wl] = {k: sub[k] for k in ("metric", "value", "unit", "ms_per_step", "config", "cpu_baseline", "dtype")}
        import numpy as np

        key = np.random.default_rng(1).integers(-(1 << 10), 1 << 10, size=(4, 4096), dtype=np.int64)
        extra["poly_mul"] = {"metric": "negacyclic integer polynomial products: transforms/s (k fwd + 1 inv per product)",
                             "cpu_baseline": cpu_poly_mul(key, 4.0)}
        extra["poly_mul"]["value"] = extra["poly_mul"]["cpu_baseline"]["value"]
        extra["cpu_n1024_1thread"] = cpu_single_polynomial()
        line["extra"] = extra
    print(json.dumps(line), file=_JSON_OUT, flush=True)


# ------------------------------------------------------------------------------------------------------
# GPU arm
# ------------------------------------------------------------------------------------------------------

class Ctx:
    pass


def parity_check(ctx):
    """Correctness evidence for EVERY rank's device, outside the timed regions and without the oracle:
    (1) the reference's own golden vector (src/unordered.rs:1176-9396, committed under tests/golden) through the
    golden plan (Dif4, 32) on this rank's GPU, bit for bit; (2) the headline plan (Dif16, 256), N = 2048, on 64 rows
    seeded identically on every rank: sha256 of the output bits must agree across ranks (rank 0's device is the one
    the GPU test-suite checks against the oracle)."""
    import numpy as np
    import torch

    C, local, dist, dev = ctx.C, ctx.local, ctx.dist, ctx.dev
    gold = os.path.join(ROOT, "tests", "golden")
    x = np.fromfile(os.path.join(gold, "unordered_n2048_dif4_b32_input.f64"), dtype=np.complex128)
    target = np.fromfile(os.path.join(gold, "unordered_n2048_dif4_b32_target.f64"), dtype=np.complex128)
    A = C.ordered.FftAlgo
    gp = C.unordered.Plan(2048, C.unordered.Method.UserProvided(A.Dif4, 32), device=local)
    d = torch.from_numpy(x.copy()).to(dev)
    gp.fwd(d)
    golden_ok = bool(np.array_equal(d.cpu().numpy().view(np.uint64), target.view(np.uint64)))
    hp = C.unordered.Plan(2048, C.unordered.Method.UserProvided(A.Dif16, 256), device=local)
    rows = torch.view_as_complex(torch.rand(64, 2048, 2, dtype=torch.float64, generator=torch.Generator().manual_seed(7))).contiguous().to(dev)
    hp.fwd(rows)
    digest = hashlib.sha256(rows.cpu().numpy().tobytes()).digest()
    hp.inv(rows)
    digest = hashlib.sha256(digest + rows.cpu().numpy().tobytes()).digest()
    word = int.from_bytes(digest[:7], "little")
    agree, ranks_ok = True, 1
    if dist is not None:
        t = torch.tensor([word, int(golden_ok)], dtype=torch.int64, device=dev)
        allv = [torch.zeros_like(t) for _ in range(ctx.world)]
        dist.all_gather(allv, t)
        agree = all(int(v[0]) == int(allv[0][0]) for v in allv)
        ranks_ok = sum(int(v[1]) for v in allv)
    else:
        ranks_ok = int(golden_ok)
    if not golden_ok:
        raise SystemExit("rank %d: GPU result differs from the reference's golden vector" % ctx.rank)
    if not agree:
        raise SystemExit("ranks disagree on the bits of the headline plan's output")
    return {"golden_vector_bit_exact_ranks": ranks_ok, "of_ranks": ctx.world, "headline_plan_checksum_agrees_across_ranks": agree,
            "checksum": "%014x" % word,
            "how": "per rank, before the timed region: golden vector (tests/golden) through (Dif4, 32) bit for bit; sha256 of fwd and inv "
                   "output bits of (Dif16, 256) on 64 identically seeded rows compared across ranks"}


def gpu_leg(ctx, workload, n, batch, base_n, algo, steps, warmup, e2e_steps, cpu_seconds, sustained_seconds=0.0, dd_full=False):
    """One workload on this rank's GPU: device-resident timing, optional sustained run, e2e through the host API,
    roofline, CPU baseline (rank 0 of a 1-GPU run).  Returns the JSON-line dict on rank 0, None elsewhere."""
    import numpy as np
    import torch

    C, rank, world, local, dist, dev = ctx.C, ctx.rank, ctx.world, ctx.local, ctx.dist, ctx.dev
    from concrete_fft_b200.sharding import max_over_ranks

    g = torch.Generator(device=dev).manual_seed(0x5EED0000 + rank)
    A = C.ordered.FftAlgo
    planes = data = None
    if workload in ("c64", "ordered"):
        if workload == "ordered":
            plan = C.ordered.Plan(n, C.ordered.Method.Measure(), device=local, allow_large=n > 1024)
        else:
            plan = C.unordered.Plan(n, C.unordered.Method.UserProvided(A[algo], base_n), device=local)
            plan.autotune()  # kernel variant only; the plan (order, bits) is fixed by UserProvided
        data = torch.view_as_complex(torch.rand(batch, n, 2, dtype=torch.float64, device=dev, generator=g)).contiguous()
        bytes_per_launch = 2 * 16 * n * batch

        def fwd():
            plan.fwd(data)

        def inv():
            plan.inv(data)
    else:
        plan = C.fft128.Plan(n, device=local)
        planes = [torch.rand(batch, n, dtype=torch.float64, device=dev, generator=g), torch.zeros(batch, n, dtype=torch.float64, device=dev),
                  torch.rand(batch, n, dtype=torch.float64, device=dev, generator=g), torch.zeros(batch, n, dtype=torch.float64, device=dev)]
        if dd_full:  # lo = (u - 1/2) * 2^-53 * hi: below half an ulp of hi, so (hi, lo) is a normalised double-double
            for hi_i in (0, 2):
                planes[hi_i + 1] = (torch.rand(batch, n, dtype=torch.float64, device=dev, generator=g) - 0.5) * planes[hi_i] * 2.0 ** -53
        bytes_per_launch = 2 * 32 * n * batch

        def fwd():
            plan.fwd(*planes)

        def inv():
            plan.inv(*planes)
    inv_scale = 1.0 / n
    # values grow by n per step: rescale before f64 overflows (2^1023), i.e. every ~900 / log2(n) steps
    rescale_every = max(1, 900 // max(1, n.bit_length() - 1))

    def rescale(factor):
        if workload == "f128":
            for p in planes:
                p.mul_(factor)
        else:
            data.mul_(factor)

    def barrier():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
            torch.cuda.synchronize()

    # ---- device-resident timing ------------------------------------------------------------
    W = max(warmup, 3)
    for _ in range(W):
        fwd(); inv()
    rescale(float(n) ** -W)
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    K = steps
    ev = [[torch.cuda.Event(enable_timing=True) for _ in range(3)] for _ in range(K)]
    launches0 = C.launch_count()
    t_begin, t_end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    t_begin.record()
    for i in range(K):
        ev[i][0].record()
        fwd()
        ev[i][1].record()
        inv()
        ev[i][2].record()
        if (i + 1) % rescale_every == 0:  # fwd+inv multiplies by n: one exact power-of-two rescale
            rescale(float(n) ** -rescale_every)  # (a single elementwise launch, < 1 % of the interval)
    t_end.record()
    barrier()
    launches = C.launch_count() - launches0
    clocks = sampler.stop() if rank == 0 else None
    total_ms = t_begin.elapsed_time(t_end)
    fwd_ms = sum(e[0].elapsed_time(e[1]) for e in ev) / K
    inv_ms = sum(e[1].elapsed_time(e[2]) for e in ev) / K
    kb = min(K, 20)  # the first steps run before the power cap pulls the SM clock down: the kernel at full clock
    fwd_ms_first = sum(e[0].elapsed_time(e[1]) for e in ev[:kb]) / kb
    total_ms = max_over_ranks(total_ms, dist, dev)
    value = 2.0 * batch * world * K / (total_ms * 1e-3)
    rescale(float(n) ** -(K % rescale_every))

    # ---- sustained: the same step back to back for >= sustained_seconds -----------------------
    sustained = None
    if sustained_seconds > 0:
        ks = max(K, int(sustained_seconds * 1e3 / (total_ms / K)) + 1)
        s2 = ClockSampler(local)
        if rank == 0:
            s2.start()
        b0, b1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        b0.record()
        for i in range(ks):
            fwd(); inv()
            if (i + 1) % rescale_every == 0:
                rescale(float(n) ** -rescale_every)
        b1.record()
        barrier()
        ms = max_over_ranks(b0.elapsed_time(b1), dist, dev)
        c2 = s2.stop() if rank == 0 else None
        sustained = {"value": 2.0 * batch * world * ks / (ms * 1e-3), "unit": "transforms/s", "steps": ks, "seconds": ms * 1e-3,
                     "ms_per_step": ms / ks, "hbm_gbs_whole_step": 2 * bytes_per_launch * world * ks / (ms * 1e-3) / 1e9, "clocks": c2}
        rescale(float(n) ** -(ks % rescale_every))

    # ---- end to end through the host-memory API ----------------------------------------------
    e2e = None
    if e2e_steps > 0:
        if workload in ("c64", "ordered"):
            host = torch.view_as_complex(torch.rand(batch, n, 2, dtype=torch.float64)).contiguous().pin_memory()
            hnp = host.numpy()
            h2d = d2h = batch * n * 16

            def e2e_step():
                plan.fwd_inv_host(hnp)
                return float(hnp[0, 0].real)
        else:
            hp = [torch.rand(batch, n, dtype=torch.float64).pin_memory() for _ in range(4)]
            hn = [p.numpy() for p in hp]
            h2d = d2h = batch * n * 32

            def e2e_step():
                plan.fwd_inv_host(*hn)
                for p in hn:
                    p[:1] *= inv_scale
                return float(hn[0][0, 0])
        e2e_step()
        if workload in ("c64", "ordered"):
            np.multiply(hnp, inv_scale, out=hnp)
        barrier()
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            e2e_step()
            if workload in ("c64", "ordered"):
                np.multiply(hnp[:1], inv_scale, out=hnp[:1])
        barrier()
        el = time.perf_counter() - t0
        el = max_over_ranks(el, dist, dev)
        e2e = {"value": 2.0 * batch * world * e2e_steps / el, "unit": "transforms/s",
               "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "steps": e2e_steps,
               "ms_per_step": 1e3 * el / e2e_steps,
               "api": "Plan.fwd_inv_host -> cfft_c64_fwd_inv_host (pinned host buffers)" if workload in ("c64", "ordered")
                      else "fft128.Plan.fwd_inv_host on pinned host planes -> cfft_f128_fwd_inv_host"}
        if ctx.numa_note:
            e2e["host_placement"] = ctx.numa_note
        if workload in ("c64", "ordered"):
            del host, hnp
        else:
            del hp, hn

    kernel_name = plan.kernel_name()
    del plan, data, planes
    torch.cuda.empty_cache()
    if rank != 0:
        return None

    peak, peak_src, _ = measured_peaks()
    traffic, traffic_src, ncu_fp64 = None, None, None
    try:
        tj = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
        rec = tj.get("%s:%d:%d" % (workload, n, batch))
        if rec:
            traffic, traffic_src = rec["traffic_bytes"], "profiles/traffic.json (ncu dram__bytes_read+write, %s)" % rec["kernel"]
            if "sm__pipe_fp64_cycles_active_pct" in rec:
                ncu_fp64 = dict(rec["sm__pipe_fp64_cycles_active_pct"], source=rec.get("source", "profiles/traffic.json"))
    except Exception:
        pass
    dom_ms = fwd_ms  # fwd and inv launches are the same kernel family; the fwd launch is reported
    achieved = bytes_per_launch / (dom_ms * 1e-3) / 1e9
    roofline = {"bound": "hbm", "kernel": kernel_name + " (fwd launch)", "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": achieved / peak, "frac_of_8TBps_spec": achieved / 8000.0, "traffic": traffic, "traffic_source": traffic_src, "peak_source": peak_src,
                "algorithmic_bytes_per_launch": bytes_per_launch,
                "fwd_ms": fwd_ms, "inv_ms": inv_ms,
                "frac_first_%d_steps" % kb: bytes_per_launch / (fwd_ms_first * 1e-3) / 1e9 / peak,
                "inv_achieved": bytes_per_launch / (inv_ms * 1e-3) / 1e9,
                "gflops_5nlog2n": 5.0 * n * (n.bit_length() - 1) * value / 1e9}
    if sustained:
        roofline["frac_sustained_whole_step"] = sustained["hbm_gbs_whole_step"] / world / peak
    if workload == "f128":
        instr = 94.0 * (n // 2) * (n.bit_length() - 1)  # FP64 instructions per transform
        rate = instr * batch / (fwd_ms * 1e-3)
        probe = ctx.fp64_probe
        nominal = 64 * 148 * 1965e6
        if probe:
            fp64_peak = max(probe["dfma_per_s"], probe["dadd_per_s"], probe["mix_per_s"])
            src = ("measured: cfft_probe_fp64_issue_rate on this GPU just before the run (DFMA %.2f / DADD %.2f / 1:7 mix %.2f T instr/s at "
                   "%.0f MHz in-kernel = %.1f lanes/clk/SM; the highest is the peak)" % (
                       probe["dfma_per_s"] / 1e12, probe["dadd_per_s"] / 1e12, probe["mix_per_s"] / 1e12, probe["sm_mhz"],
                       fp64_peak / probe["sm_count"] / (probe["sm_mhz"] * 1e6)))
        else:
            fp64_peak, src = nominal, "nominal 64 FP64 lanes x 148 SMs x 1965 MHz (probe unavailable)"
        # fft128 is bound by the FP64 pipe, not by HBM: the roofline is in FP64 instructions (94 per butterfly)
        roofline = {"bound": "fp64", "kernel": roofline["kernel"], "achieved": rate / 1e12, "peak": fp64_peak / 1e12,
                    "unit": "T FP64 instr/s", "frac": rate / fp64_peak, "traffic": traffic, "traffic_source": traffic_src,
                    "peak_source": src, "fp64_instr_per_transform": instr, "frac_of_nominal_64x148x1965MHz": rate / nominal,
                    "inv_frac": instr * batch / (inv_ms * 1e-3) / fp64_peak,
                    "ncu_sm__pipe_fp64_cycles_active_pct": ncu_fp64,
                    "flop_frac_of_37.2TF": 106.0 / 94.0 * rate / (2 * nominal),
                    "algorithmic_bytes_per_launch": bytes_per_launch, "hbm_achieved_gbs": achieved, "hbm_frac": achieved / peak,
                    "fwd_ms": fwd_ms, "inv_ms": inv_ms}

    cpu = None
    if cpu_seconds > 0 and world == 1:
        import oracle_lib as O

        O.build()
        threads = host_threads()
        rate, rows, csteps, el = cpu_port_rate(workload, n, batch, base_n, algo, threads, cpu_seconds)
        cpu = {"value": rate, "unit": "transforms/s", "cores": threads, "kind": "port",
               "sample": "%d polynomials (DRAM-sized sample) x %d fwd+inv steps in %.1f s (same n / plan as the GPU run)" % (rows, csteps, el)}

    line = {
        "metric": "batched %s FFT transforms/s (fwd+inv)" % METRIC_NAME[workload],
        "value": value, "unit": "transforms/s", "n_gpus": world, "steps": K, "warmup": W,
        "ms_per_step": total_ms / K, "higher_is_better": True,
        "scaling": "strong" if workload == "ordered" else "weak", "vs_baseline": None,
        "dtype": "f64x2 (double-double)" if workload == "f128" else "f64",
        "data": "synthetic (full-width double-double: lo ~ U(-1/2, 1/2) ulp(hi))" if workload == "f128" and dd_full else "synthetic",
        "config": config_dict(workload, n, batch, base_n, algo, world),
        "hbm_gbs_whole_step": 2 * bytes_per_launch * world * K / (total_ms * 1e-3) / 1e9,
        "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": int(launches), "clocks": clocks,
    }
    if sustained:
        line["value_sustained"] = sustained["value"]
        line["sustained"] = sustained
    return line


def cpu_poly_mul(key_np, seconds):
    """the same product step through the oracle port on all host threads (N = 4096, k terms, key shared)"""
    import numpy as np
    import oracle_lib as O

    O.build()
    k, npoly = key_np.shape
    n = npoly // 2
    threads = host_threads()
    ref = O.UnorderedPlan(n, O.DIF16, 256, fast=True)
    fb = ref.fwd(O.poly_fold_twist(key_np))
    rows = 1024
    rng = np.random.default_rng(0)
    sample = rng.integers(-(1 << 20), 1 << 20, size=(rows, k, npoly), dtype=np.int64)
    O.poly_mul(ref, sample[:64], fb, threads=threads)
    w0, reps = time.perf_counter(), 0
    while time.perf_counter() - w0 < seconds:
        O.poly_mul(ref, sample, fb, threads=threads)
        reps += 1
    el = time.perf_counter() - w0
    return {"value": rows * (k + 1) * reps / el, "unit": "transforms/s", "products_per_s": rows * reps / el, "cores": threads,
            "kind": "port", "sample": "%d products x %d repetitions in %.1f s: fold + twist, %d reference forward transforms, "
            "element-wise multiply-accumulate, inverse, untwist + round per product" % (rows, reps, el, k)}


def poly_leg(ctx, steps, e2e_steps, cpu_seconds):
    """SURVEY 8f rank 3 as a bench leg: the negacyclic product step on INTEGER polynomials -- out[r] = sum_k a[r][k] * key[k]
    mod X^N + 1, N = 4096 (fft size 2048), k = 4 terms, key shared by the batch and resident in the Fourier domain on the GPU
    (a GGSW row against a batch of decomposed ciphertexts) -- through cfft_c64_poly_mul (device-resident) and
    cfft_c64_poly_mul_host (polynomials in host memory).  Counted in transforms (k forward + 1 inverse per row)."""
    import numpy as np
    import torch

    C, rank, world, local, dist, dev = ctx.C, ctx.rank, ctx.world, ctx.local, ctx.dist, ctx.dev
    from concrete_fft_b200.sharding import max_over_ranks

    npoly, n, k, batch = 4096, 2048, 4, 16384
    A = C.ordered.FftAlgo
    plan = C.unordered.Plan(n, C.unordered.Method.UserProvided(A.Dif16, 256), device=local)
    g = torch.Generator(device=dev).manual_seed(0xB0B0 + rank)
    a = torch.randint(-(1 << 20), 1 << 20, (batch, k, npoly), dtype=torch.int64, device=dev, generator=g)
    key = torch.randint(-(1 << 10), 1 << 10, (k, npoly), dtype=torch.int64, device=dev, generator=g)
    fkey = plan.fwd_poly(key)
    out = torch.empty((batch, npoly), dtype=torch.int64, device=dev)

    def barrier():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
            torch.cuda.synchronize()

    for _ in range(3):
        plan.poly_mul(a, fkey, out=out)
    launches0 = C.launch_count()
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    t0.record()
    for _ in range(steps):
        plan.poly_mul(a, fkey, out=out)
    t1.record()
    barrier()
    launches = C.launch_count() - launches0
    ms = max_over_ranks(t0.elapsed_time(t1), dist, dev) / steps
    tr_per_step = batch * (k + 1)
    value = tr_per_step * world / (ms * 1e-3)
    e2e = None
    if e2e_steps > 0:
        ha = a.cpu().pin_memory()
        ho = torch.empty((batch, npoly), dtype=torch.int64).pin_memory()
        plan.poly_mul(ha.numpy(), fkey, out=ho.numpy())
        assert torch.equal(ho, out.cpu())  # the host entry returns the device call's bits
        barrier()
        w0 = time.perf_counter()
        for _ in range(e2e_steps):
            plan.poly_mul(ha.numpy(), fkey, out=ho.numpy())
        barrier()
        el = max_over_ranks(time.perf_counter() - w0, dist, dev)
        e2e = {"value": tr_per_step * world * e2e_steps / el, "unit": "transforms/s", "h2d_bytes_per_step": batch * k * npoly * 8,
               "d2h_bytes_per_step": batch * npoly * 8, "steps": e2e_steps, "ms_per_step": 1e3 * el / e2e_steps,
               "products_per_s": batch * world * e2e_steps / el,
               "api": "Plan.poly_mul(numpy) -> cfft_c64_poly_mul_host (pinned host polynomials, Fourier-domain key resident on the GPU)"}
        del ha, ho
    fused = plan.has_fused_poly_kernel(k)
    key_np = key.cpu().numpy()
    del plan, a, out, fkey, key
    torch.cuda.empty_cache()
    if rank != 0:
        return None
    peak, peak_src, _ = measured_peaks()
    bytes_per_step = batch * (k + 1) * npoly * 8  # k polynomials in, one out per row; the shared key stays in L2
    line = {"metric": "negacyclic integer polynomial products: transforms/s (k fwd + 1 inv per product)", "value": value, "unit": "transforms/s",
            "products_per_s": batch * world / (ms * 1e-3), "n_gpus": world, "steps": steps, "ms_per_step": ms, "dtype": "i64 polynomials, f64 transforms",
            "config": {"workload": "out[r] = sum_k a[r][k] * key[k] mod X^N + 1, integers in / out (SURVEY 8f rank 3)", "N_poly": npoly, "n": n,
                       "k_terms": k, "batch_per_gpu": batch, "key": "shared by the batch, Fourier domain, device resident", "fused_kernel": bool(fused)},
            "roofline": {"bound": "hbm", "kernel": "c64_fwd_mul_inv_kernel<2048, ..., PIN, POUT>", "achieved": bytes_per_step / (ms * 1e-3) / 1e9, "peak": peak,
                         "unit": "GB/s", "frac": bytes_per_step / (ms * 1e-3) / 1e9 / peak, "peak_source": peak_src,
                         "algorithmic_bytes_per_launch": bytes_per_step, "traffic": None},
            "e2e": e2e, "gpu_launches": int(launches), "cpu_baseline": None}
    if cpu_seconds > 0 and world == 1:
        line["cpu_baseline"] = cpu_poly_mul(key_np, cpu_seconds)
    return line


_JSON_OUT = sys.stdout


def keep_stdout_for_the_json_line():
    """The contract is ONE JSON line on stdout.  Libraries print there too (NCCL's version banner on the first collective), so file
    descriptor 1 is pointed at stderr for the rest of the run and the line goes to a private copy of the original stdout."""
    global _JSON_OUT
    try:
        sys.stdout.flush()
        _JSON_OUT = os.fdopen(os.dup(1), "w")
        os.dup2(2, 1)
    except OSError:
        _JSON_OUT = sys.stdout


def main():
    keep_stdout_for_the_json_line()
    args = parse_args()
    if args.impl == "reference":
        return run_reference(args)

    import torch

    import concrete_fft_b200 as C

    ctx = Ctx()
    ctx.C = C
    ctx.rank = int(os.environ.get("RANK", "0"))
    ctx.world = int(os.environ.get("WORLD_SIZE", "1"))
    ctx.local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        sys.exit("bench.py needs a CUDA device (no CPU fallback); use --impl reference for the CPU arm")
    torch.cuda.set_device(ctx.local)
    ctx.numa_note = bind_to_gpu_numa_node(ctx.local) if ctx.world > 1 else None
    ctx.dist = None
    if ctx.world > 1:
        import torch.distributed as dist

        dist.init_process_group("nccl", device_id=torch.device("cuda", ctx.local))
        ctx.dist = dist
    ctx.dev = torch.device("cuda", ctx.local)
    try:
        ctx.fp64_probe = C.probe_fp64_issue_rate(ctx.local)
    except Exception:  # noqa: BLE001
        ctx.fp64_probe = None

    parity = parity_check(ctx)
    n, batch, base_n, algo = workload_defaults(args.workload, args.n, args.batch, args.base_n, args.algo)
    default_run = args.workload == "c64" and not (args.n or args.batch or args.base_n or args.algo)
    line = gpu_leg(ctx, args.workload, n, batch, base_n, algo, args.steps, args.warmup, 0 if args.no_e2e else args.e2e_steps,
                   0.0 if args.no_cpu else args.cpu_seconds, args.sustained_seconds, args.dd_full)
    extra = {}
    if default_run and not args.no_extra:
        for wl, st in (("f128", 20), ("ordered", 10)):
            en, eb, ebn, ealgo = workload_defaults(wl)
            sub = gpu_leg(ctx, wl, en, eb, ebn, ealgo, st, 3, 0 if args.no_e2e else 2, 0.0 if args.no_cpu else 5.0, 0.0)
            if sub is not None:
                extra[wl] = sub
        sub = poly_leg(ctx, 10, 0 if args.no_e2e else 2, 0.0 if args.no_cpu else 4.0)
        if sub is not None:
            extra["poly_mul"] = sub
        if ctx.rank == 0 and not args.no_cpu:
            import oracle_lib  # noqa: F401

            extra["cpu_n1024_1thread"] = cpu_single_polynomial()
    if ctx.rank == 0:
        line["parity"] = parity
        if ctx.fp64_probe:
            line["fp64_probe"] = ctx.fp64_probe
        if extra:
            line["extra"] = extra
        print(json.dumps(line), file=_JSON_OUT, flush=True)
    if ctx.dist is not None:
        ctx.dist.destroy_process_group()


if __name__ == "__main__":
    main()
