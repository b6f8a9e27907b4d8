#!/bin/bash
mkdir -p gpurun_out
RC_MARCH_MAP=PPDDDD ncu --set full --clock-control none --import-source on -k regex:"k_march|k_gather|k_gbuffer" -s 24 -c 8 -o gpurun_out/prof_r1c_lr4k python bench.py --steps 2 --warmup 3 --no-cpu-baseline --workload living_room_4k > gpurun_out/ncu_c.log 2>&1
ls -la gpurun_out/prof_r1c_lr4k.ncu-rep
