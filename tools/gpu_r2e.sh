#!/bin/bash
# session call 3: prefetch A/B on the persistent two-phase kernel, the one-shot column tiles, the fused product kernels and the
# standard-order kernels; then the full suite
mkdir -p gpurun_out
o=gpurun_out/r2e_ab.txt
for cfg in "CFFT_B200_TWOPASS_PREFETCH=0" "CFFT_B200_TWOPASS_PREFETCH=1"; do
  echo "== persistent two-phase: $cfg" >> $o
  for lg in 14 15 16; do env $cfg timeout 300 python tools/cmp_variants.py $lg 8 >> $o 2>&1; done
done
for cfg in "CFFT_B200_COLUMN_PREFETCH=0" "CFFT_B200_COLUMN_PREFETCH=1" "CFFT_B200_COLUMN_PREFETCH=2"; do
  echo "== one-shot column tiles: $cfg" >> $o
  for lg in 14 16 18 20; do env $cfg timeout 300 python tools/cmp_variants.py $lg 2 9 >> $o 2>&1; done
done
ORD="2048:Dif16:ord 4096:Dif16:ord 8192:Dif16:ord 65536:Dif16:ord"
for cfg in "CFFT_B200_FAST_PREFETCH=0" "CFFT_B200_FAST_PREFETCH=1"; do
  echo "== standard-order kernels: $cfg" >> $o
  env $cfg timeout 300 python tools/time_plans.py $ORD >> $o 2>&1
done
for cfg in "CFFT_B200_FUSED_MUL_PREFETCH=0" "CFFT_B200_FUSED_MUL_PREFETCH=1"; do
  env $cfg timeout 900 python tools/fused_mul_probe.py 2048 4096 8192 > gpurun_out/r2e_fused_mul_probe_pf${cfg: -1}.jsonl 2>&1
done
bash tools/gpu_suite.sh r2e
