set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -8
for e in 0 2 3 -1; do
  echo "== RC_MARCH_ENTRY=$e teapot"; RC_MARCH_ENTRY=$e python bench.py --steps 20 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['stage_ms'], d['level_ms'])"
  echo "== RC_MARCH_ENTRY=$e lr4k"; RC_MARCH_ENTRY=$e python bench.py --steps 10 --warmup 3 --workload living_room_4k --no-cpu-baseline 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['stage_ms'], d['level_ms'])"
done
