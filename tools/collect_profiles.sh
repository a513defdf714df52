#!/bin/bash
# Every ncu capture the round's profiles/ are made from (run on the GPU box; ~3 minutes).
set -u
tools/ncu_quick.sh r02
tools/ncu_full.sh rollout rollout_kernel 1 1 rollout --rollout-log2 24
tools/ncu_full.sh greedy rollout_kernel 1 1 greedy --rollout-log2 22
tools/ncu_full.sh traj rollout_kernel 3 1 traj
tools/ncu_full.sh step1m step_kernel 66 2 step
tools/ncu_full.sh step8m step_kernel 73 1 step
tools/ncu_full.sh after afterstates_kernel 1 1 afterstates
tools/ncu_full.sh env env_step_kernel 41 1 env
tools/ncu_full.sh ringappend ring_append_kernel 18 1 ring
tools/ncu_full.sh ringsample ring_sample_kernel 1 1 ring
tools/ncu_full.sh envring env_step_kernel 41 1 ring
python bench.py --steps 2 --warmup 3 --quick --no-cpu-baseline > gpurun_out/bench_quick.json 2> gpurun_out/bench_quick.err && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_bench_launches.csv python bench.py --steps 2 --warmup 3 --quick --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1
ls -la gpurun_out/prof_*.ncu-rep
