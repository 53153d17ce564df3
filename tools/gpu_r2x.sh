#!/bin/bash
# last check of the session: full GPU suite, smoke, sanitizer over every kernel family
mkdir -p gpurun_out
python -m pytest tests -m gpu -q > gpurun_out/r2x_pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/r2x_pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2x_smoke.log 2>&1
export CFFT_B200_NO_AUTOTUNE=1
for tool in memcheck racecheck; do
  timeout 1200 compute-sanitizer --tool $tool --kernel-regex kns=cfft python tools/sanitize_small.py 2>&1 | grep -E "sanitize_small|ERROR SUMMARY|RACECHECK SUMMARY" | sed "s/^/$tool: /" >> gpurun_out/r2x_sanitizer.txt
done
tail -3 gpurun_out/r2x_pytest_gpu.log; tail -1 gpurun_out/r2x_smoke.log; cat gpurun_out/r2x_sanitizer.txt
