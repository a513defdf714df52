#!/bin/bash
# one `ncu --set full` capture per hot kernel (run on the GPU box): tools/ncu_full.sh <tag> <which> <launch-skip> <count>
set -u
tag=$1; which=$2; skip=${3:-0}; cnt=${4:-1}; kre=${5:-.}
ncu --set full --clock-control none --import-source on -k regex:"$kre" -s $skip -c $cnt -f -o gpurun_out/prof_${tag} python tools/profile_kernels.py $which > gpurun_out/prof_${tag}.log 2>&1
ncu -i gpurun_out/prof_${tag}.ncu-rep --page raw --csv > gpurun_out/prof_${tag}_raw.csv 2>/dev/null
