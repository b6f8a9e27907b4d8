# A/B: per-level fused march (RC_MARCH_BATCH=0) vs all levels in one launch + merges (1), entry frontier off / on
python -m pytest tests -m gpu -x -q 2>&1 | tail -4
for wl in teapot_1080p test_room_1080p living_room_4k; do
for b in 0 1; do for e in 0 3; do
  echo "== $wl RC_MARCH_BATCH=$b RC_MARCH_ENTRY=$e"
  RC_MARCH_BATCH=$b RC_MARCH_ENTRY=$e python bench.py --steps 20 --warmup 3 --workload $wl --no-cpu-baseline 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['e2e']['ms_per_step'], d['stage_ms'])"
done; done; done
