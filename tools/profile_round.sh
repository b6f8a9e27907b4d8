# Round profile capture (one B200): ncu launch lists, ncu --set full of one frame's kernels, bench lines.
# Usage: bash tools/profile_round.sh <tag>      (outputs under gpurun_out/)
tag=${1:-r1f}
K='regex:k_march|k_gather|k_gbuffer|k_probes|k_link_entry|k_need'
for wl in teapot_1080p living_room_4k; do
  ncu --metrics gpu__time_duration.sum --clock-control none -s 60 -c 48 --csv --log-file gpurun_out/launches_${tag}_${wl}.csv \
      python bench.py --steps 2 --warmup 3 --workload $wl --no-cpu-baseline > /dev/null 2>&1
  RC_GRAPH=0 ncu --set full --clock-control none --import-source on -k "$K" -s 42 -c 14 -f -o gpurun_out/prof_${tag}_${wl} \
      python bench.py --steps 2 --warmup 3 --workload $wl --no-cpu-baseline > gpurun_out/ncu_${tag}_${wl}.log 2>&1
  ncu -i gpurun_out/prof_${tag}_${wl}.ncu-rep --page raw --csv > gpurun_out/raw_${tag}_${wl}.csv 2>/dev/null
done
python bench.py --steps 20 --warmup 3 > gpurun_out/bench_${tag}_teapot1080p.json 2> gpurun_out/bench_${tag}_teapot1080p.err
python bench.py --steps 20 --warmup 3 --workload living_room_4k --cpu-sample-div 2 > gpurun_out/bench_${tag}_living_room4k.json 2> gpurun_out/bench_${tag}_living_room4k.err
python bench.py --steps 20 --warmup 3 --workload test_room_1080p --no-cpu-baseline > gpurun_out/bench_${tag}_test_room1080p.json 2>/dev/null
python bench.py --steps 10 --warmup 3 --workload sonic_8k --no-cpu-baseline > gpurun_out/bench_${tag}_sonic8k.json 2>/dev/null
python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/bench_${tag}_reference.json 2>/dev/null
nproc > gpurun_out/nproc.txt
ls -la gpurun_out | tail -20
