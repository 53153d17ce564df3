#!/bin/bash
# session call 4: one-exchange ordered rows kernel (A/B + ncu), compile-time schedules for whole-transform plans, env-variant tests
mkdir -p gpurun_out
python -m pytest tests/test_gpu_env_variants.py tests/test_gpu_c64.py -m gpu -x -q -k "env_variant or compile_time or ord16 or ordered_above or generic_register or ordered_fused" > gpurun_out/r2f_pytest_new.log 2>&1; echo "exit $?" >> gpurun_out/r2f_pytest_new.log
o=gpurun_out/r2f_ab.txt
ORD="16384:Dif16:ord 65536:Dif16:ord 262144:Dif16:ord 1048576:Dif16:ord"
for cfg in "CFFT_B200_ROWS_STD_TWO_EXCHANGES=1" "CFFT_B200_ROWS_STD_TWO_EXCHANGES=0" "CFFT_B200_ROWS_STD_TWO_EXCHANGES=1 CFFT_B200_L2_CHUNK_MB=16 CFFT_B200_L2_STREAMS=4" "CFFT_B200_ROWS_STD_TWO_EXCHANGES=0 CFFT_B200_L2_CHUNK_MB=16 CFFT_B200_L2_STREAMS=4"; do
  echo "== ordered n >= 2^14: $cfg" >> $o
  env $cfg timeout 300 python tools/time_plans.py $ORD >> $o 2>&1
done
SP="1024:Dif8:ord 1024:Dit8:ord 1024:Dit16:ord 512:Dif8:ord 512:Dit8:ord 1024:Dif4:ord"
for cfg in "CFFT_B200_REGS_NO_SPEC=1" "CFFT_B200_REGS_NO_SPEC=0"; do
  echo "== whole-transform plans: $cfg" >> $o
  if [ "$cfg" = "CFFT_B200_REGS_NO_SPEC=0" ]; then timeout 300 python tools/time_plans.py $SP >> $o 2>&1; else env $cfg timeout 300 python tools/time_plans.py $SP >> $o 2>&1; fi
done
timeout 600 ncu --set full --import-source on --clock-control none -k regex:rows256_std16 -c 2 -f -o gpurun_out/r2f_ncu_rows_std16 python tools/prof_plan.py ordered 65536 Dif16 512 > gpurun_out/r2f_ncu_rows.log 2>&1
python bench.py --workload ordered --steps 10 --warmup 3 > gpurun_out/r2f_bench_ordered.json 2> gpurun_out/r2f_bench_ordered.err
