#!/bin/bash
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:"k_march|k_gather|k_gbuffer" -s 24 -c 8 -o gpurun_out/prof_r1b python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_b.log 2>&1
ls -la gpurun_out/prof_r1b.ncu-rep
