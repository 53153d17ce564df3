#!/bin/bash
# run-to-run spread of the headline line on one box: five default-length runs and five 20-step runs, device legs only
mkdir -p gpurun_out
for i in 1 2 3 4 5; do python bench.py --no-cpu --no-extra --e2e-steps 1 >> gpurun_out/r2r_bench_200steps.jsonl 2>/dev/null; done
for i in 1 2 3 4 5; do python bench.py --no-cpu --no-extra --e2e-steps 1 --steps 20 --sustained-seconds 0 >> gpurun_out/r2r_bench_20steps.jsonl 2>/dev/null; done
python - <<'PY'
import json
for f in ("gpurun_out/r2r_bench_200steps.jsonl", "gpurun_out/r2r_bench_20steps.jsonl"):
    rows = [json.loads(l) for l in open(f) if l.startswith("{")]
    print(f, [round(r["value"] / 1e6, 1) for r in rows], "frac", [round(r["roofline"]["frac"], 3) for r in rows], "sm_mhz", [r["clocks"]["sm_mhz"] for r in rows], "sustained", [round((r.get("value_sustained") or 0) / 1e6, 1) for r in rows])
PY
