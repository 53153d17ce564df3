#!/bin/bash
# session call 1: TMA data-path A/B, variant timings at n = 2^13 .. 2^16, then the full suite
mkdir -p gpurun_out
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o /tmp/ubench_tma tools/ubench_tma.cu && timeout 300 /tmp/ubench_tma > gpurun_out/r2c_ubench_tma.txt 2>&1
for lg in 13 14 16; do
  timeout 300 python tools/cmp_variants.py $lg 2 9 auto >> gpurun_out/r2c_variants.txt 2>&1
done
CFFT_B200_COLPIPE=0 timeout 300 python tools/cmp_variants.py 16 2 9 >> gpurun_out/r2c_variants_nopipe.txt 2>&1
PLANS="2048:Dif16:1024 2048:Dif16:512 2048:Dif8:512 2048:Dif4:32 2048:Dit16:1024 1024:Dif8:512 4096:Dif16:1024 4096:Dif8:512 2048:Dif16:256"
timeout 300 python tools/time_plans.py $PLANS > gpurun_out/r2c_plans_spec.txt 2>&1
CFFT_B200_REGS_NO_SPEC=1 timeout 300 python tools/time_plans.py $PLANS > gpurun_out/r2c_plans_generic.txt 2>&1
bash tools/gpu_suite.sh r2c
