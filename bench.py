#!/usr/bin/env python3
"""bench.py -- headline benchmark of the batched c64 FFT hot path (BASELINE.json config 2).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
                  [--workload c64|f128] [--n N] [--batch B]

Step   = one pass of the hot path over one batch: unordered fwd THEN inv of `batch`
         polynomials of size n, in place (2 * batch transforms).
value  = transforms/s, whole job (all ranks), inputs resident in HBM, CUDA-event timed,
         max over ranks.  Inputs (2 GiB per GPU at the default size) are far larger than the
         126 MB L2, so no flush is needed between iterations.
e2e    = the same step through the public host-memory call (Plan.fwd_inv_host ->
         cfft_c64_fwd_inv_host): pinned host buffers, H2D + fwd + inv + D2H inside the timed
         region.
roofline    = achieved HBM GB/s of the dominant kernel (algorithmic bytes 2 * 16 * n per
              transform per launch / average launch time from CUDA events on the launch
              stream) against MEASURED_PEAKS.json.
cpu_baseline= the oracle port of the reference algorithm (-O3 build, bit-identical results) on
              the host cores, bounded sample of the same workload.
--impl reference times that CPU port with all host threads on the same metric/config (the
Rust reference itself cannot be built in this image: no cargo/rustc).
"""
import argparse
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
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
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


def workload_defaults(args):
    if args.workload == "ordered":
        # BASELINE.json configs[2]: standard-order N = 2^16, batch 4096 in total, sharded across the GPUs
        n = args.n or 65536
        world = int(os.environ.get("WORLD_SIZE", "1"))
        batch = args.batch or max(1, 4096 // world)
        return n, batch, 256, "Dif16"
    if args.workload == "c64":
        n = args.n or 2048
        batch = args.batch or 65536
        base_n = args.base_n or min(n, 256)
        algo = args.algo or "Dif16"
    else:
        n = args.n or 2048
        batch = args.batch or 16384
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


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    n, batch, base_n, algo = workload_defaults(args)
    threads = host_threads()
    import oracle_lib as O

    O.build()
    # Each step = fwd+inv over the WHOLE per-GPU batch when `--steps` of them finish in about a minute
    # (the default c64 workload does: 200 x 2 GiB); otherwise over the largest power-of-two slice that does.
    probe_rows = cpu_sample_rows(args.workload, n, batch, 64 << 20)
    step, rescale = cpu_workload(args.workload, n, base_n, algo, probe_rows, threads)
    step()
    t0 = time.perf_counter()
    step()
    per_row = (time.perf_counter() - t0) / probe_rows
    del step, rescale
    rows = batch
    while rows > probe_rows and rows * per_row * (args.steps + args.warmup) > 75.0:
        rows //= 2
    rows = max(rows, probe_rows)
    step, rescale = cpu_workload(args.workload, n, base_n, algo, rows, threads)
    every = max(1, 900 // max(1, n.bit_length() - 1))
    for _ in range(args.warmup):
        step()
        rescale(1)
    t0 = time.perf_counter()
    for i in range(args.steps):
        step()
        if (i + 1) % every == 0:
            rescale(every)
    el = time.perf_counter() - t0
    value = 2.0 * rows * args.steps / el
    bytes_per = n * (32 if args.workload == "f128" else 16)
    sample = "%d of %d polynomials per step (%.0f MiB, streamed from host DRAM), fwd+inv, %d host threads" % (
        rows, batch, rows * bytes_per / 2 ** 20, threads)
    line = {
        "impl": "reference",
        "metric": "batched %s FFT transforms/s (fwd+inv)" % METRIC_NAME[args.workload],
        "value": value, "unit": "transforms/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * el / args.steps, "higher_is_better": True,
        "scaling": "strong" if args.workload == "ordered" else "weak", "vs_baseline": None,
        "dtype": "f64x2 (double-double)" if args.workload == "f128" else "f64", "data": "synthetic",
        "config": config_dict(args.workload, n, batch, base_n, algo, args.gpus),
        "cpu_baseline": {"value": value, "unit": "transforms/s", "cores": threads, "kind": "port", "sample": sample,
                         "note": "C restatement of the reference algorithm (oracle/, -O3 AVX2+FMA build: two complex per "
                                 "register like the reference's c64x2 path; bit-identical to the reference's golden "
                                 "vector; persistent thread pool); the Rust crate cannot be built here"},
        "e2e": {"value": value, "unit": "transforms/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


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


def main():
    args = parse_args()
    if args.impl == "reference":
        return run_reference(args)

    import numpy as np
    import torch

    import concrete_fft_b200 as C

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        sys.exit("bench.py needs a CUDA device (no CPU fallback); use --impl reference for the CPU arm")
    torch.cuda.set_device(local)
    numa_note = bind_to_gpu_numa_node(local) if world > 1 else None
    dist = None
    if world > 1:
        import torch.distributed as dist

        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    n, batch, base_n, algo = workload_defaults(args)
    dev = torch.device("cuda", local)
    g = torch.Generator(device=dev).manual_seed(0x5EED0000 + rank)
    A = C.ordered.FftAlgo

    if args.workload in ("c64", "ordered"):
        if args.workload == "ordered":
            plan = C.ordered.Plan(n, C.ordered.Method.Measure(), device=local, allow_large=n > 1024)
        else:
            plan = C.unordered.Plan(n, C.unordered.Method.UserProvided(A[algo], base_n), device=local)
            plan.autotune()  # kernel variant only; the plan (order, bits) is fixed by UserProvided
        data = torch.view_as_complex(torch.rand(batch, n, 2, dtype=torch.float64, device=dev, generator=g)).contiguous()
        bytes_per_launch = 2 * 16 * n * batch
        inv_scale = 1.0 / n

        def fwd():
            plan.fwd(data)

        def inv():
            plan.inv(data)

        def renorm():
            data.mul_(inv_scale)
    else:
        plan = C.fft128.Plan(n, device=local)
        planes = [torch.rand(batch, n, dtype=torch.float64, device=dev, generator=g), torch.zeros(batch, n, dtype=torch.float64, device=dev),
                  torch.rand(batch, n, dtype=torch.float64, device=dev, generator=g), torch.zeros(batch, n, dtype=torch.float64, device=dev)]
        if args.dd_full:  # lo = (u - 1/2) * 2^-53 * hi: below half an ulp of hi, so (hi, lo) is a normalised double-double
            for hi_i in (0, 2):
                planes[hi_i + 1] = (torch.rand(batch, n, dtype=torch.float64, device=dev, generator=g) - 0.5) * planes[hi_i] * 2.0 ** -53
        bytes_per_launch = 2 * 32 * n * batch
        inv_scale = 1.0 / n

        def fwd():
            plan.fwd(*planes)

        def inv():
            plan.inv(*planes)

        def renorm():
            for p in planes:
                p.mul_(inv_scale)

    # values grow by n per step: rescale before f64 overflows (2^1023), i.e. every ~900 / log2(n) steps
    rescale_every = max(1, 900 // max(1, n.bit_length() - 1))

    def rescale(factor):
        if args.workload == "f128":
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
    for _ in range(max(args.warmup, 3)):
        fwd(); inv()
    rescale(float(n) ** -max(args.warmup, 3))
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    K = args.steps
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
    from concrete_fft_b200.sharding import max_over_ranks

    total_ms = max_over_ranks(total_ms, dist, dev)
    value = 2.0 * batch * world * K / (total_ms * 1e-3)

    # ---- end to end through the host-memory API ----------------------------------------------
    e2e = None
    if not args.no_e2e:
        if args.workload in ("c64", "ordered"):
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
        if args.workload in ("c64", "ordered"):
            np.multiply(hnp, inv_scale, out=hnp)
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.e2e_steps):
            e2e_step()
            if args.workload in ("c64", "ordered"):
                np.multiply(hnp[:1], inv_scale, out=hnp[:1])
        barrier()
        el = time.perf_counter() - t0
        el = max_over_ranks(el, dist, dev)
        e2e = {"value": 2.0 * batch * world * args.e2e_steps / el, "unit": "transforms/s",
               "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "steps": args.e2e_steps,
               "ms_per_step": 1e3 * el / args.e2e_steps,
               "api": "Plan.fwd_inv_host -> cfft_c64_fwd_inv_host (pinned host buffers)" if args.workload in ("c64", "ordered")
                      else "fft128.Plan.fwd_inv_host on pinned host planes -> cfft_f128_fwd_inv_host"}
        if numa_note:
            e2e["host_placement"] = numa_note

    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return

    peak, peak_src, _ = measured_peaks()
    traffic, traffic_src, ncu_fp64 = None, None, None
    try:
        tj = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
        rec = tj.get("%s:%d:%d" % (args.workload, n, batch))
        if rec:
            traffic, traffic_src = rec["traffic_bytes"], "profiles/traffic.json (ncu dram__bytes_read+write, %s)" % rec["kernel"]
            if "sm__pipe_fp64_cycles_active_pct" in rec:
                ncu_fp64 = dict(rec["sm__pipe_fp64_cycles_active_pct"], source=rec.get("source", "profiles/traffic.json"))
    except Exception:
        pass
    dom_ms = fwd_ms  # fwd and inv launches are the same kernel family; the fwd launch is reported
    achieved = bytes_per_launch / (dom_ms * 1e-3) / 1e9
    roofline = {"bound": "hbm", "kernel": plan.kernel_name() + " (fwd launch)", "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": achieved / peak, "frac_of_8TBps_spec": achieved / 8000.0, "traffic": traffic, "traffic_source": traffic_src, "peak_source": peak_src,
                "algorithmic_bytes_per_launch": bytes_per_launch,
                "fwd_ms": fwd_ms, "inv_ms": inv_ms,
                "inv_achieved": bytes_per_launch / (inv_ms * 1e-3) / 1e9,
                "gflops_5nlog2n": 5.0 * n * (n.bit_length() - 1) * value / 1e9}
    if args.workload == "f128":
        instr = 94.0 * (n // 2) * (n.bit_length() - 1)  # FP64 instructions per transform
        sm_mhz = (clocks or {}).get("sm_mhz") or 1965.0
        fp64_peak = 64 * 148 * sm_mhz * 1e6
        rate = instr * batch / (fwd_ms * 1e-3)
        # fft128 is bound by the FP64 pipe, not by HBM: report the roofline in FP64 instructions
        # (94 per butterfly) against 64 lanes x 148 SMs x the SM clock sampled during the run
        roofline = {"bound": "fp64", "kernel": roofline["kernel"], "achieved": rate / 1e12, "peak": fp64_peak / 1e12,
                    "unit": "T FP64 instr/s", "frac": rate / fp64_peak, "traffic": traffic, "traffic_source": traffic_src,
                    "peak_source": "64 FP64 lanes x 148 SMs x sampled SM clock (%.0f MHz); MEASURED_PEAKS.json has no FP64 entry" % sm_mhz,
                    "fp64_instr_per_transform": instr, "frac_at_max_clock": rate / (64 * 148 * 1965e6),
                    "ncu_sm__pipe_fp64_cycles_active_pct": ncu_fp64,
                    "flop_frac_of_37.2TF": 106.0 / 94.0 * rate / (2 * 64 * 148 * 1965e6),
                    "algorithmic_bytes_per_launch": bytes_per_launch, "hbm_achieved_gbs": achieved, "hbm_frac": achieved / peak,
                    "fwd_ms": fwd_ms, "inv_ms": inv_ms}

    cpu = None
    if not args.no_cpu:
        import oracle_lib as O

        O.build()
        threads = host_threads()
        rate, rows, steps, el = cpu_port_rate(args.workload, n, batch, base_n, algo, threads, args.cpu_seconds)
        cpu = {"value": rate, "unit": "transforms/s", "cores": threads, "kind": "port",
               "sample": "%d polynomials (DRAM-sized sample) x %d fwd+inv steps in %.1f s (same n / plan as the GPU run)" % (rows, steps, el)}

    line = {
        "metric": "batched %s FFT transforms/s (fwd+inv)" % METRIC_NAME[args.workload],
        "value": value, "unit": "transforms/s", "n_gpus": world, "steps": K, "warmup": max(args.warmup, 3),
        "ms_per_step": total_ms / K, "higher_is_better": True,
        "scaling": "strong" if args.workload == "ordered" else "weak", "vs_baseline": None,
        "dtype": "f64x2 (double-double)" if args.workload == "f128" else "f64",
        "data": "synthetic (full-width double-double: lo ~ U(-1/2, 1/2) ulp(hi))" if args.workload == "f128" and args.dd_full else "synthetic",
        "config": config_dict(args.workload, n, batch, base_n, algo, world),
        "hbm_gbs_whole_step": 2 * bytes_per_launch * world * K / (total_ms * 1e-3) / 1e9,
        "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": int(launches), "clocks": clocks,
    }
    print(json.dumps(line), flush=True)
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
