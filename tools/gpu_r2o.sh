#!/bin/bash
mkdir -p gpurun_out
o=gpurun_out/r2o_small_n_prefetch.txt
for pf in 0 1; do
  echo "CFFT_B200_FAST_PREFETCH=$pf" >> $o
  for lg in 8 9 10; do CFFT_B200_FAST_PREFETCH=$pf timeout 300 python tools/cmp_variants.py $lg 1 >> $o 2>&1; done
done
cat > /tmp/prof_mul2.py <<'PY'
import sys, os
sys.path.insert(0, os.environ.get("GRAFT_REPO_ROOT", "."))
import torch
import concrete_fft_b200 as C
n, k, rows = 2048, 4, 8192
plan = C.unordered.Plan(n, C.unordered.Method.UserProvided(C.ordered.FftAlgo.Dif16, 256))
a = torch.view_as_complex(torch.rand(rows, k, n, 2, dtype=torch.float64, device="cuda") - 0.5).contiguous()
b = torch.view_as_complex(torch.rand(k, 2, n, 2, dtype=torch.float64, device="cuda") - 0.5).contiguous()
out = torch.empty((rows, 2, n), dtype=torch.complex128, device="cuda")
for _ in range(3):
    plan.fwd_mul_inv_multi(a, b, out)
    plan.fwd_mul_inv(a, b[:, 0].contiguous())
torch.cuda.synchronize()
PY
timeout 600 ncu --set full --import-source on --clock-control none -k regex:fwd_mul_inv -s 4 -c 2 -f -o gpurun_out/r2o_ncu_mul2 python /tmp/prof_mul2.py > gpurun_out/r2o_ncu.log 2>&1
cat $o
