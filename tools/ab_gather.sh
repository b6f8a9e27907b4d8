# A/B of the software-pipelined gather (RC_GATHER_TILES) + the full GPU parity suite.  Outputs under gpurun_out/.
tag=${1:-r1i}
mkdir -p gpurun_out
( time timeout 1200 python -m pytest tests -m gpu -q --maxfail=5 ) > gpurun_out/pytest_gpu_${tag}.log 2>&1
echo "pytest exit: $?" >> gpurun_out/pytest_gpu_${tag}.log
tail -6 gpurun_out/pytest_gpu_${tag}.log
for wl in living_room_4k teapot_1080p; do
  for t in 1 2 4 8 16; do
    echo "== $wl gather_tiles=$t"
    RC_GATHER_TILES=$t python bench.py --steps 20 --warmup 3 --workload $wl --no-cpu-baseline 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(round(d['ms_per_step'],4), round(d['e2e']['ms_per_step'],4), {k: round(v,4) for k,v in d['stage_ms'].items()})"
  done
done 2>&1 | tee gpurun_out/ab_gather_${tag}.txt
for occ in 8 12; do
  echo "== living_room_4k march_occ=$occ"
  RC_MARCH_OCC=$occ python bench.py --steps 20 --warmup 3 --workload living_room_4k --no-cpu-baseline 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(round(d['ms_per_step'],4), {k: round(v,4) for k,v in d['stage_ms'].items()})"
done 2>&1 | tee -a gpurun_out/ab_gather_${tag}.txt
