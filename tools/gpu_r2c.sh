#!/bin/bash
# session call 1: TMA data-path A/B, compile-time schedules vs interpreter, L2 prefetch A/B, variant timings, full suite
mkdir -p gpurun_out
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o /tmp/ubench_tma tools/ubench_tma.cu && timeout 300 /tmp/ubench_tma > gpurun_out/r2c_ubench_tma.txt 2>&1
PLANS="2048:Dif16:1024 2048:Dif16:512 2048:Dif8:512 2048:Dif4:32 2048:Dit16:1024 1024:Dif8:512 4096:Dif16:1024 4096:Dif8:512 2048:Dif16:256"
timeout 300 python tools/time_plans.py $PLANS > gpurun_out/r2c_plans_spec.txt 2>&1
CFFT_B200_REGS_NO_SPEC=1 timeout 300 python tools/time_plans.py $PLANS > gpurun_out/r2c_plans_generic.txt 2>&1
for pf in 0 1 2; do
  echo "CFFT_B200_FAST_PREFETCH=$pf" >> gpurun_out/r2c_fast_prefetch.txt
  for lg in 11 12 13; do CFFT_B200_FAST_PREFETCH=$pf timeout 300 python tools/cmp_variants.py $lg 1 >> gpurun_out/r2c_fast_prefetch.txt 2>&1; done
  echo "CFFT_B200_F128_PREFETCH=$pf" >> gpurun_out/r2c_f128_prefetch.txt
  CFFT_B200_F128_PREFETCH=$pf timeout 300 python tools/time_f128.py 10 11 12 13 15 >> gpurun_out/r2c_f128_prefetch.txt 2>&1
done
for lg in 13 14 16; do
  timeout 300 python tools/cmp_variants.py $lg 2 9 auto >> gpurun_out/r2c_variants.txt 2>&1
done
CFFT_B200_COLPIPE=0 timeout 300 python tools/cmp_variants.py 16 2 9 >> gpurun_out/r2c_variants_nopipe.txt 2>&1
bash tools/gpu_suite.sh r2c
# ncu: the compile-time-schedule kernel, and the two passes of n = 2^16 with the batch L2-resident (32 MiB) and HBM-sized
timeout 600 ncu --set full --import-source on --clock-control none -k regex:c64_regs_spec -c 2 -f -o gpurun_out/r2c_ncu_spec_dif16_1024 python tools/prof_plan.py unordered 2048 Dif16 1024 16384 > gpurun_out/r2c_ncu_spec.log 2>&1
CMP_BATCH=32 CMP_REPS=2 CFFT_B200_FAST_VARIANT=9 timeout 600 ncu --set full --import-source on --clock-control none -k regex:colpipe\|fast_b256 -s 8 -c 4 -f -o gpurun_out/r2c_ncu_n65536_l2res python tools/cmp_variants.py 16 9 > gpurun_out/r2c_ncu_l2res.log 2>&1
CMP_BATCH=512 CMP_REPS=2 timeout 600 ncu --set full --import-source on --clock-control none -k regex:colpipe\|fast_b256 -s 8 -c 4 -f -o gpurun_out/r2c_ncu_n65536_hbm python tools/cmp_variants.py 16 9 > gpurun_out/r2c_ncu_hbm.log 2>&1
ls -la gpurun_out
