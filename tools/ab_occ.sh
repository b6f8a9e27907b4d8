for wl in teapot_1080p living_room_4k; do for o in 8 10 12 16; do
  echo "== $wl RC_MARCH_OCC=$o"
  RC_MARCH_OCC=$o python bench.py --steps 20 --warmup 3 --workload $wl --no-cpu-baseline 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['stage_ms']['march'])"
done; done
