#!/bin/bash
# session call 5: fft128 HBM-pass prefetch + tile-size choice for n >= 4096, ordered autotune with smaller chunks, full suite
mkdir -p gpurun_out
o=gpurun_out/r2g_f128.txt
for cfg in "CFFT_B200_F128_TILEMAX=4096" "CFFT_B200_F128_TILEMAX=2048"; do
  echo "== $cfg" >> $o
  env $cfg timeout 600 python tools/time_f128.py 12 13 15 16 18 >> $o 2>&1
done
python bench.py --workload ordered --steps 10 --warmup 3 > gpurun_out/r2g_bench_ordered.json 2> gpurun_out/r2g_bench_ordered.err
python -m pytest tests -m gpu -q > gpurun_out/r2g_pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/r2g_pytest_gpu.log
tail -5 gpurun_out/r2g_pytest_gpu.log
