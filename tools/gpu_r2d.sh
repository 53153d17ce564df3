#!/bin/bash
# session call 2: data-path A/B (L2-resident sweeps, bulk store), prefetch / bulk-store / twiddle-split A/B on the real kernels,
# large-n variants with the one-shot column tiles, new tests
mkdir -p gpurun_out
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o /tmp/ubench_tma tools/ubench_tma.cu && timeout 300 /tmp/ubench_tma > gpurun_out/r2d_ubench_tma.txt 2>&1
python -m pytest tests/test_gpu_env_variants.py tests/test_gpu_c64.py -m gpu -x -q -k "env_variant or strided or replicas or compile_time" > gpurun_out/r2d_pytest_new.log 2>&1; echo "exit $?" >> gpurun_out/r2d_pytest_new.log
python -m pytest tests/test_gpu_f128.py -m gpu -x -q -k "strided or replicas" >> gpurun_out/r2d_pytest_new.log 2>&1; echo "exit $?" >> gpurun_out/r2d_pytest_new.log
o=gpurun_out/r2d_ab.txt
for cfg in "" "CFFT_B200_BULK_STORE=1" "CFFT_B200_TW_SPLIT_L1=1" "CFFT_B200_BULK_STORE=1 CFFT_B200_TW_SPLIT_L1=1"; do
  echo "== $cfg" >> $o
  for lg in 11 12 13; do env $cfg timeout 300 python tools/cmp_variants.py $lg 1 >> $o 2>&1; done
done
for cfg in "CFFT_B200_CLUSTER_PREFETCH=0" "CFFT_B200_CLUSTER_PREFETCH=1"; do
  echo "== $cfg" >> $o
  for lg in 13 14; do env $cfg timeout 300 python tools/cmp_variants.py $lg 4 >> $o 2>&1; done
done
for cfg in "" "CFFT_B200_BULK_STORE=1" "CFFT_B200_FAST_PREFETCH=0"; do
  echo "== large n: $cfg" >> $o
  for lg in 14 15 16 17 18 20; do env $cfg timeout 300 python tools/cmp_variants.py $lg 2 9 auto >> $o 2>&1; done
done
echo "== spec plans, bulk store n/a; fused product probe" >> $o
timeout 600 python tools/fused_mul_probe.py 8192 > gpurun_out/r2d_fused_mul_probe_8192.jsonl 2>&1
