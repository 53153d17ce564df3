#!/bin/bash
# Full GPU suite of one session: tests, smoke, default bench.  Logs under gpurun_out/<tag>_*.
tag=${1:-run}
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/${tag}_pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/${tag}_pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${tag}_smoke.log 2>&1
python bench.py > gpurun_out/${tag}_bench_default.json 2> gpurun_out/${tag}_bench_default.err
tail -3 gpurun_out/${tag}_pytest_gpu.log; tail -2 gpurun_out/${tag}_smoke.log; head -c 600 gpurun_out/${tag}_bench_default.json
