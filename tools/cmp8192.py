import os, sys, json
sys.path.insert(0, os.getcwd())
import torch
import concrete_fft_b200 as C
n, batch = 8192, 16384
dev = torch.device("cuda", 0)
data = torch.view_as_complex(torch.rand(batch, n, 2, dtype=torch.float64, device=dev)).contiguous()
def timeit(fn, reps=10):
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(reps + 1)]
    ev[0].record()
    for i in range(reps):
        fn(); ev[i + 1].record()
    torch.cuda.synchronize()
    ts = sorted(ev[i].elapsed_time(ev[i + 1]) for i in range(reps))
    return ts[len(ts) // 2]
for var in ["1", "4"]:
    os.environ["CFFT_B200_FAST_VARIANT"] = var
    plan = C.unordered.Plan(n, C.unordered.Method.UserProvided(C.ordered.FftAlgo.Dif16, 256))
    del os.environ["CFFT_B200_FAST_VARIANT"]
    for _ in range(3):
        plan.fwd(data); plan.inv(data); data.mul_(1.0 / n)
    f = timeit(lambda: plan.fwd(data)); data.mul_(float(n) ** -10)
    i = timeit(lambda: plan.inv(data)); 
    b = 2 * 16 * n * batch
    print(var, plan.kernel_name(), "fwd %.3f ms %.0f GB/s   inv %.3f ms %.0f GB/s" % (f, b / f / 1e6, i, b / i / 1e6), flush=True)
