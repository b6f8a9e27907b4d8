# Round-2 profile capture (one B200): ncu launch lists, ncu --set full of one frame's kernels, bench lines of every workload.
# Usage: gpurun -- bash tools/profile_round2.sh <tag>      (outputs under gpurun_out/; tools/collect_profiles.sh copies them to profiles/)
tag=${1:-r2p}
K='regex:k_march|k_gather|k_gbuffer|k_probes|k_link_entry|k_need'
for wl in living_room_4k teapot_1080p; do
  ncu --metrics gpu__time_duration.sum --clock-control none -s 56 -c 42 --csv --log-file gpurun_out/launches_${tag}_${wl}.csv \
      python bench.py --steps 2 --warmup 3 --workload $wl --no-cpu-baseline > /dev/null 2>&1
  RC_GRAPH=0 ncu --set full --clock-control none --import-source on -k "$K" -s 42 -c 14 -f -o gpurun_out/prof_${tag}_${wl} \
      python bench.py --steps 2 --warmup 3 --workload $wl --no-cpu-baseline > gpurun_out/ncu_${tag}_${wl}.log 2>&1
  ncu -i gpurun_out/prof_${tag}_${wl}.ncu-rep --page raw --csv > gpurun_out/raw_${tag}_${wl}.csv 2>/dev/null
done
python bench.py --steps 20 --warmup 3 > gpurun_out/bench_${tag}_living_room4k.json 2> gpurun_out/bench_${tag}_living_room4k.err
python bench.py --steps 20 --warmup 3 --workload teapot_1080p > gpurun_out/bench_${tag}_teapot1080p.json 2> gpurun_out/bench_${tag}_teapot1080p.err
python bench.py --steps 20 --warmup 3 --workload test_room_1080p --no-cpu-baseline > gpurun_out/bench_${tag}_test_room1080p.json 2>/dev/null
python bench.py --steps 10 --warmup 3 --workload sonic_8k --no-cpu-baseline > gpurun_out/bench_${tag}_sonic8k.json 2>/dev/null
python bench.py --steps 10 --warmup 3 --workload cube_512 > gpurun_out/bench_${tag}_cube512.json 2>/dev/null
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_${tag}_reference.json 2>/dev/null
python -m pytest tests -m gpu -q 2>&1 | tail -8 > gpurun_out/pytest_gpu_${tag}.log
make -C radiancecascade_b200/csrc > /dev/null 2>&1; ./radiancecascade_b200/rc_headless --help > /dev/null 2>&1
nproc > gpurun_out/nproc.txt
tail -3 gpurun_out/pytest_gpu_${tag}.log
for f in gpurun_out/bench_${tag}_*.json; do python - <<PY
import json
try:
    j=json.loads(open("$f").read().strip().splitlines()[-1]); print("$f", round(j["ms_per_step"],4), round(j["value"],3), round(j.get("e2e",{}).get("ms_per_step",0),4))
except Exception as e: print("$f", "FAILED", e)
PY
done
