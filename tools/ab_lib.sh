# A/B of an alternative build of the same ABI (RC_B200_LIB=<path to .so>): parity subset + bench lines.
# Usage: bash tools/ab_lib.sh <tag> <lib.so> [more libs...]
tag=$1; shift
mkdir -p gpurun_out
run() { python bench.py --steps 20 --warmup 3 --workload $1 --no-cpu-baseline 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(round(d['ms_per_step'],4), round(d['e2e']['ms_per_step'],4), round(d['all_rays']['ms_per_step'],4), {k: round(v,4) for k,v in d['stage_ms'].items()}, [round(x,4) for x in d['level_ms']])"; }
for lib in "$@"; do
  export RC_B200_LIB=$PWD/$lib
  echo "#### $lib"
  ( timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q --maxfail=5 -k "closest_hit or committed_golden or culling_leaves or cascades_and_irradiance or entry_frontier or gbuffer" 2>&1 | tail -3 )
  for wl in living_room_4k teapot_1080p test_room_1080p; do echo "== $wl"; run $wl; done
done 2>&1 | tee gpurun_out/ab_lib_${tag}.txt
