#!/bin/bash
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -s 40 -c 40 --csv --log-file gpurun_out/launches_r1_c.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_l.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"k_march|k_gather|k_gbuffer|k_probes|k_link" -s 30 -c 10 -o gpurun_out/prof_r1d python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_d.log 2>&1
ls -la gpurun_out/prof_r1d.ncu-rep gpurun_out/launches_r1_c.csv
