#!/bin/bash
# final measurements of a session: full suite + smoke + default bench, reference arm, size sweep with cuFFT / CPU beside it,
# ncu launch list of the bench command and DRAM traffic of the dominant kernels at the bench batch.   usage: gpu_final.sh <tag>
tag=${1:-final}
mkdir -p gpurun_out
bash tools/gpu_suite.sh $tag
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${tag}_bench_reference.json 2> gpurun_out/${tag}_bench_reference.err
python tools/sweep.py --workload both --min 4 --max 20 --cufft --cpu --out gpurun_out/${tag}_sweep.jsonl --table gpurun_out/${tag}_sweep_table.md > gpurun_out/${tag}_sweep.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file gpurun_out/${tag}_launches_c64.csv python bench.py --steps 3 --warmup 3 --no-cpu --no-e2e --no-extra --sustained-seconds 0 > gpurun_out/${tag}_launches_c64.log 2>&1
ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:c64_fast_b256 -s 4 -c 2 --csv --log-file gpurun_out/${tag}_traffic_c64.csv python tools/prof_one.py c64 2048 65536 > /dev/null 2>&1
ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active --clock-control none -k regex:f128_tile -s 4 -c 2 --csv --log-file gpurun_out/${tag}_traffic_f128.csv python tools/prof_one.py f128 2048 16384 > /dev/null 2>&1
ncu --set full --import-source on --clock-control none -k regex:c64_fast_b256 -s 4 -c 2 -f -o gpurun_out/${tag}_ncu_c64_2048 python tools/prof_one.py c64 2048 16384 > /dev/null 2>&1
ncu --set full --import-source on --clock-control none -k regex:c64_fast_b256 -s 4 -c 2 -f -o gpurun_out/${tag}_ncu_c64_8192 python tools/prof_one.py c64 8192 4096 > /dev/null 2>&1
ncu --set full --import-source on --clock-control none -k regex:c64_fast_b256 -s 4 -c 2 -f -o gpurun_out/${tag}_ncu_ordered_2048 python tools/prof_one.py ordered 2048 16384 > /dev/null 2>&1
python tools/time_plans.py 2048:Dif16:1024 2048:Dif8:512 2048:Dif4:32 2048:Dit8:512 2048:Dit16:512 4096:Dit16:1024 1024:Dit8:512 1024:Dif8:ord 1024:Dit16:ord > gpurun_out/${tag}_plans_spec.txt 2>&1
ls -la gpurun_out | tail -20
