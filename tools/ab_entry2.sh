for e in 0 -1; do
  echo "== living_room_4k RC_MARCH_ENTRY=$e"
  RC_MARCH_ENTRY=$e python bench.py --steps 20 --warmup 3 --workload living_room_4k --no-cpu-baseline 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['e2e']['ms_per_step'], d['stage_ms'], d['level_ms'])"
done
python bench.py --steps 20 --warmup 3 2>&1 | tail -3
