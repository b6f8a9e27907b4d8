#!/bin/bash
# ncu --set full of one kernel family on one workload.  Usage: bash tools/ncu_one.sh <tag> <workload> <kernel-regex> [skip] [count]
tag=$1; wl=$2; k=$3; skip=${4:-4}; cnt=${5:-1}
RC_GRAPH=0 ncu --set full --clock-control none --import-source on -k "regex:$k" -s $skip -c $cnt -f -o gpurun_out/prof_${tag} \
    python tools/stage_times.py $wl --frames 2 > gpurun_out/ncu_${tag}.log 2>&1
ncu -i gpurun_out/prof_${tag}.ncu-rep --page raw --csv > gpurun_out/raw_${tag}.csv 2>/dev/null
python tools/ncu_summary.py gpurun_out/raw_${tag}.csv > gpurun_out/summary_${tag}.csv
cat gpurun_out/summary_${tag}.csv | cut -c1-600
