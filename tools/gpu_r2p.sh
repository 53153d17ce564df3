#!/bin/bash
mkdir -p gpurun_out
o=gpurun_out/r2p_prefetch_ord16_spec.txt
ORD="16:Dif16:ord 32:Dif16:ord 64:Dif16:ord 128:Dif16:ord 512:Dif16:ord 1024:Dif16:ord"
SP="2048:Dif16:1024 2048:Dif8:512 2048:Dif4:32 1024:Dif8:512 4096:Dif16:1024 1024:Dif8:ord 512:Dif8:ord"
for pf in 0 1; do
  echo "== CFFT_B200_ORD16_PREFETCH=$pf CFFT_B200_REGS_PREFETCH=$pf" >> $o
  CFFT_B200_ORD16_PREFETCH=$pf CFFT_B200_REGS_PREFETCH=$pf timeout 600 python tools/time_plans.py $ORD $SP >> $o 2>&1
done
python -m pytest tests/test_gpu_c64.py -m gpu -q -k "ord16 or compile_time or ordered_all_algos or golden" > gpurun_out/r2p_pytest.log 2>&1; echo "exit $?" >> gpurun_out/r2p_pytest.log
tail -3 gpurun_out/r2p_pytest.log; cat $o
