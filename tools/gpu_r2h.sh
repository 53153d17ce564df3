#!/bin/bash
# session call 6: compute-sanitizer over every kernel family incl. the round-2 paths
mkdir -p gpurun_out
export CFFT_B200_NO_AUTOTUNE=1
for tool in memcheck synccheck racecheck; do
  timeout 1200 compute-sanitizer --tool $tool --kernel-regex kns=cfft python tools/sanitize_small.py > gpurun_out/r2h_sanitizer_$tool.log 2>&1
  echo "== $tool exit $?" >> gpurun_out/r2h_sanitizer_summary.txt
  grep -E "sanitize_small|ERROR SUMMARY|RACECHECK SUMMARY" gpurun_out/r2h_sanitizer_$tool.log | tail -3 >> gpurun_out/r2h_sanitizer_summary.txt
done
CFFT_B200_ROWS_STD_ONE_EXCHANGE=1 CFFT_B200_COLPIPE=1 timeout 900 compute-sanitizer --tool memcheck --kernel-regex kns=cfft python tools/sanitize_small.py > gpurun_out/r2h_sanitizer_memcheck_optin.log 2>&1
echo "== memcheck, opt-in kernels (one-exchange ordered rows, persistent column kernel) exit $?" >> gpurun_out/r2h_sanitizer_summary.txt
grep -E "sanitize_small|ERROR SUMMARY" gpurun_out/r2h_sanitizer_memcheck_optin.log | tail -2 >> gpurun_out/r2h_sanitizer_summary.txt
cat gpurun_out/r2h_sanitizer_summary.txt
