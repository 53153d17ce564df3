#!/usr/bin/env python3
"""PCIe reference points for the e2e number: pinned H2D, D2H and concurrent bidirectional copy
bandwidth through torch (cudaMemcpyAsync), plus the host-API pipeline at several settings."""
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch

    nbytes = 1 << 30
    h1 = torch.empty(nbytes, dtype=torch.uint8).pin_memory()
    h2 = torch.empty(nbytes, dtype=torch.uint8).pin_memory()
    d1 = torch.empty(nbytes, dtype=torch.uint8, device="cuda")
    d2 = torch.empty(nbytes, dtype=torch.uint8, device="cuda")
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()

    def timed(fn, reps=5):
        fn(); torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(reps):
            fn()
        torch.cuda.synchronize()
        return (time.perf_counter() - t0) / reps

    t = timed(lambda: d1.copy_(h1, non_blocking=True))
    print("H2D pinned   : %.1f GB/s" % (nbytes / t / 1e9))
    t = timed(lambda: h2.copy_(d2, non_blocking=True))
    print("D2H pinned   : %.1f GB/s" % (nbytes / t / 1e9))

    def both():
        with torch.cuda.stream(s1):
            d1.copy_(h1, non_blocking=True)
        with torch.cuda.stream(s2):
            h2.copy_(d2, non_blocking=True)
    t = timed(both)
    print("bidirectional: %.1f GB/s each way" % (nbytes / t / 1e9))


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "matrix":
        for chunk in (8, 16, 32, 64, 128):
            for slots in (2, 3, 4):
                env = dict(os.environ, CFFT_B200_PIPE_CHUNK_MB=str(chunk), CFFT_B200_PIPE_SLOTS=str(slots))
                out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "2", "--no-cpu", "--e2e-steps", "4"],
                                     env=env, capture_output=True, text=True).stdout.strip().splitlines()[-1]
                import json

                e = json.loads(out)["e2e"]
                print("chunk %3d MB slots %d: e2e %.2f M transforms/s, %.1f ms/step, %.1f GB/s each way"
                      % (chunk, slots, e["value"] / 1e6, e["ms_per_step"], e["h2d_bytes_per_step"] / e["ms_per_step"] / 1e6), flush=True)
    else:
        main()
