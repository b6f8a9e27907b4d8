#!/bin/bash
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
python bench.py --steps 10 --warmup 3 2>&1 | tail -2 | tee gpurun_out/bench_r1_a.json
python bench.py --steps 10 --warmup 3 --separate-merge --no-cpu-baseline 2>&1 | tail -1 | tee gpurun_out/bench_r1_sep.json
python bench.py --steps 5 --warmup 3 --workload living_room_4k --no-cpu-baseline 2>&1 | tail -1 | tee gpurun_out/bench_r1_lr4k.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches_r1.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_b.log 2>&1
tail -25 gpurun_out/launches_r1.csv
