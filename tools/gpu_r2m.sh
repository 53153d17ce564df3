#!/bin/bash
# standard-order fused kernels after the one-exchange rewrite: parity (ordered tests, env variants, sanitizer on the new path) and timing
mkdir -p gpurun_out
python -m pytest tests/test_gpu_c64.py tests/test_gpu_env_variants.py -m gpu -q -k "ordered or env_variant or strided or cuda_graph" > gpurun_out/r2m_pytest.log 2>&1; echo "exit $?" >> gpurun_out/r2m_pytest.log
timeout 300 python tools/time_plans.py 2048:Dif16:ord 4096:Dif16:ord 8192:Dif16:ord > gpurun_out/r2m_std.txt 2>&1
cat > /tmp/san_std.py <<'PY'
import sys, os
sys.path.insert(0, os.environ.get("GRAFT_REPO_ROOT", "."))
import numpy as np, torch
import concrete_fft_b200 as C
rng = np.random.default_rng(0)
for n, batch in [(2048, 700), (4096, 350), (8192, 170), (2048, 3)]:
    p = C.ordered.Plan(n, C.ordered.Method.UserProvided(C.ordered.FftAlgo.Dif16), allow_large=True)
    x = torch.from_numpy(rng.random((batch, n)) + 1j * rng.random((batch, n))).cuda()
    p.fwd(x); p.inv(x)
torch.cuda.synchronize()
print("std sanitize done")
PY
for tool in memcheck racecheck; do CFFT_B200_NO_AUTOTUNE=1 timeout 600 compute-sanitizer --tool $tool --kernel-regex kns=cfft python /tmp/san_std.py 2>&1 | grep -E "std sanitize|ERROR SUMMARY|RACECHECK SUMMARY" >> gpurun_out/r2m_sanitizer.txt; done
tail -3 gpurun_out/r2m_pytest.log; cat gpurun_out/r2m_std.txt gpurun_out/r2m_sanitizer.txt
