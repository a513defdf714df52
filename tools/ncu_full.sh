#!/bin/bash
# One `ncu --set full` capture of selected launches (run on the GPU box, after the same command has
# exited 0 without ncu):  tools/ncu_full.sh <tag> <kernel-regex> <skip> <count> <profile_kernels.py args...>
# Leaves gpurun_out/prof_<tag>_raw.csv and _source.csv; summarise here with tools/make_profiles.py.
set -u
tag=$1; kre=$2; skip=$3; cnt=$4; shift 4
ncu --set full --clock-control none --import-source on -k regex:"$kre" -s $skip -c $cnt -f -o gpurun_out/prof_${tag} python tools/profile_kernels.py "$@" > gpurun_out/prof_${tag}.log 2>&1
ncu -i gpurun_out/prof_${tag}.ncu-rep --page raw --csv > gpurun_out/prof_${tag}_raw.csv 2>/dev/null
# the per-instruction execution counts (tools/sass_budget.py) as CSV; the report itself is dropped unless
# KEEP_REP=1: gpurun brings back at most 64 MiB and a report is 5-9 MB
ncu -i gpurun_out/prof_${tag}.ncu-rep --page source --csv > gpurun_out/prof_${tag}_source.csv 2>/dev/null
if [ "${KEEP_REP:-0}" != "1" ]; then rm -f gpurun_out/prof_${tag}.ncu-rep; fi
