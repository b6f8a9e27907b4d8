python -m pytest tests -m gpu -x -q 2>&1 | tail -3
for wl in teapot_1080p test_room_1080p living_room_4k; do
  echo "== $wl"
  python bench.py --steps 20 --warmup 3 --workload $wl --no-cpu-baseline 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['e2e']['ms_per_step'], d['stage_ms'], d['level_ms'])"
done
