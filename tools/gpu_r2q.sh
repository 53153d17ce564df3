#!/bin/bash
# LSU data-pipe wavefronts per element, measured: the model's P for each kernel family against ncu's counters
mkdir -p gpurun_out
M=l1tex__data_pipe_lsu_wavefronts.sum,l1tex__data_pipe_lsu_wavefronts_mem_shared.sum,l1tex__data_pipe_lsu_wavefronts_mem_lgds.sum,l1tex__throughput.avg.pct_of_peak_sustained_active,gpu__time_duration.sum,smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio,smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio,dram__bytes_read.sum,dram__bytes_write.sum
run() { # tag, kernel regex, skip, command...
  tag=$1; k=$2; s=$3; shift 3
  ncu --metrics $M --clock-control none -k regex:$k -s $s -c 2 --csv --log-file gpurun_out/r2q_wavefronts_$tag.csv "$@" > /dev/null 2>&1
}
run c64_2048 c64_fast_b256 4 python tools/prof_one.py c64 2048 16384
run c64_8192 c64_fast_b256 4 python tools/prof_one.py c64 8192 4096
run ordered_2048 c64_fast_b256 4 python tools/prof_one.py ordered 2048 16384
run spec_dif16_1024 c64_regs_spec 4 python tools/prof_plan.py unordered 2048 Dif16 1024 16384
run spec_dif4_32 c64_regs_spec 4 python tools/prof_plan.py unordered 2048 Dif4 32 16384
CFFT_B200_FAST_VARIANT=2 run n65536_column 'c64_column_kernel' 2 python tools/prof_one.py c64 65536 512
CFFT_B200_FAST_VARIANT=2 run n65536_rows 'c64_fast_b256' 2 python tools/prof_one.py c64 65536 512
CFFT_B200_FAST_VARIANT=2 run n32768_column 'c64_column_kernel' 2 python tools/prof_one.py c64 32768 1024
ls gpurun_out | grep r2q
