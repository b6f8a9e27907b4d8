# A/B: triangle pre-splitting budget (RC_BVH_SPLIT x triangles extra references)
for sp in 0 0.3 1.0 2.0; do
  echo "=== RC_BVH_SPLIT=$sp"
  RC_BVH_SPLIT=$sp python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "closest_hit or cascades_and_irradiance or gbuffer" 2>&1 | tail -2
  for wl in teapot_1080p test_room_1080p living_room_4k; do
    echo "== $wl"
    RC_BVH_SPLIT=$sp python bench.py --steps 10 --warmup 3 --workload $wl --no-cpu-baseline 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['stage_ms'], d['level_ms'])"
  done
done
