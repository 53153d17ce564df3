#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_gpu_c64.py -m gpu -q -k "several_outputs or fwd_mul_inv" > gpurun_out/r2n_pytest.log 2>&1; echo "exit $?" >> gpurun_out/r2n_pytest.log
timeout 600 python tools/fused_mul2_probe.py 512 1024 2048 > gpurun_out/r2n_fused_mul2_probe.txt 2>&1
tail -4 gpurun_out/r2n_pytest.log; cat gpurun_out/r2n_fused_mul2_probe.txt
