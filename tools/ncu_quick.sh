#!/bin/bash
# quick per-launch instruction counts / durations of the hot kernels (run on the GPU box)
#   tools/ncu_quick.sh <tag>
set -u
tag=${1:-x}
M="gpu__time_duration.sum,smsp__inst_executed.sum,sm__inst_executed.avg.per_cycle_elapsed,smsp__thread_inst_executed_per_inst_executed.ratio,sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active,sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active,l1tex__data_pipe_lsu_wavefronts_mem_shared.sum,l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__issue_active.avg.pct_of_peak_sustained_active"
ncu --metrics $M --clock-control none -k regex:'step_kernel|rollout_kernel' -c 40 --csv --log-file gpurun_out/ncuq_${tag}.csv python tools/profile_kernels.py quick > gpurun_out/ncuq_${tag}.log 2>&1
