#!/bin/bash
# usage: tools/prof_round.sh <prefix>   -- launch list + full captures (trace, classify/shade) of the cfg4 wavefront frame
P=${1:-q}
M=gpu__time_duration.sum,smsp__thread_inst_executed_per_inst_executed.ratio,sm__warps_active.avg.pct_of_peak_sustained_active,smsp__issue_active.avg.pct_of_peak_sustained_active,l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed,dram__bytes_read.sum,dram__bytes_write.sum
B="python bench.py --steps 1 --warmup 3 --no-cpu-baseline"
ncu --metrics $M --clock-control none -c 210 --csv --log-file gpurun_out/${P}_launches.csv $B > gpurun_out/${P}_b1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_wf_trace -c 2 -f -o gpurun_out/${P}_full_trace $B > gpurun_out/${P}_b2.log 2>&1
ncu --set full --clock-control none --import-source on -k "regex:k_wf_(classify|shade|logic)" -c 2 -f -o gpurun_out/${P}_full_shade $B > gpurun_out/${P}_b3.log 2>&1
ls -la gpurun_out | grep ${P}_
