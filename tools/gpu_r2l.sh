#!/bin/bash
# one-exchange standard-order fused kernels (ordered n = 2048 / 4096 / 8192): parity and A/B
mkdir -p gpurun_out
python -m pytest tests/test_gpu_env_variants.py -m gpu -q -k "ONE_EXCHANGE or defaults" > gpurun_out/r2l_pytest.log 2>&1; echo "exit $?" >> gpurun_out/r2l_pytest.log
o=gpurun_out/r2l_std_ab.txt
ORD="2048:Dif16:ord 4096:Dif16:ord 8192:Dif16:ord"
for cfg in "CFFT_B200_STD_ONE_EXCHANGE=0" "CFFT_B200_STD_ONE_EXCHANGE=1" "CFFT_B200_STD_ONE_EXCHANGE=1 CFFT_B200_FAST_PREFETCH=1" "CFFT_B200_STD_ONE_EXCHANGE=1 CFFT_B200_FAST_PREFETCH=0"; do
  echo "== $cfg" >> $o
  env $cfg timeout 300 python tools/time_plans.py $ORD >> $o 2>&1
done
tail -3 gpurun_out/r2l_pytest.log; cat $o
