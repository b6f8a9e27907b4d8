# A/B: direction-major ray lists (RC_LIST_DIRMAJOR bit mask of levels), PDL along the k_need chain, auto gather tiles.
tag=${1:-r1j}
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_cpp_facade.py -m gpu -q --maxfail=5 -k "list_order or pipelined or culling or headless or committed_golden" ) > gpurun_out/pytest_gpu_${tag}.log 2>&1
echo "pytest exit: $?" >> gpurun_out/pytest_gpu_${tag}.log
tail -6 gpurun_out/pytest_gpu_${tag}.log
run() { python bench.py --steps 20 --warmup 3 --workload $1 --no-cpu-baseline 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(round(d['ms_per_step'],4), round(d['e2e']['ms_per_step'],4), {k: round(v,4) for k,v in d['stage_ms'].items()}, [round(x,4) for x in d['level_ms']])"; }
for wl in living_room_4k teapot_1080p test_room_1080p; do
  for m in 0 1 3 7 15; do
    echo "== $wl list_dir_major=$m"; RC_LIST_DIRMAJOR=$m run $wl
  done
  echo "== $wl need_pdl=1"; RC_NEED_PDL=1 run $wl
  echo "== $wl need_pdl=1 list_dir_major=7"; RC_NEED_PDL=1 RC_LIST_DIRMAJOR=7 run $wl
done 2>&1 | tee gpurun_out/ab_list_${tag}.txt
