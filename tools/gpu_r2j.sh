#!/bin/bash
# session call: the fused (8, 8.2) column group of n = 2^15 -- parity and timing; the C replay of the Rust call sequences on a GPU
mkdir -p gpurun_out
python -m pytest tests/test_abi.py -q > gpurun_out/r2j_pytest.log 2>&1; echo "abi exit $?" >> gpurun_out/r2j_pytest.log
python -m pytest tests/test_gpu_c64.py tests/test_gpu_env_variants.py -m gpu -q -k "large_n_column or ordered_above or persistent or env_variant or autotune or cuda_graph" >> gpurun_out/r2j_pytest.log 2>&1; echo "gpu exit $?" >> gpurun_out/r2j_pytest.log
o=gpurun_out/r2j_n32768.txt
timeout 300 python tools/cmp_variants.py 15 2 8 9 auto >> $o 2>&1
CFFT_B200_L2_CHUNK_MB=16 CFFT_B200_L2_STREAMS=4 timeout 300 python tools/cmp_variants.py 15 2 >> $o 2>&1
timeout 300 python tools/time_plans.py 32768:Dif16:ord >> $o 2>&1
CFFT_B200_L2_CHUNK_MB=8 CFFT_B200_L2_STREAMS=4 timeout 300 python tools/time_plans.py 32768:Dif16:ord >> $o 2>&1
tail -4 gpurun_out/r2j_pytest.log; cat $o
