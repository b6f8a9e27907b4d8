python -m pytest tests -m gpu -x -q 2>&1 | tail -3
for wl in teapot_1080p living_room_4k; do for g in 0 1; do
  echo "== $wl RC_GRAPH=$g"
  RC_GRAPH=$g python bench.py --steps 20 --warmup 3 --workload $wl --no-cpu-baseline 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], 'e2e', d['e2e']['ms_per_step'], 'blocking', d['e2e']['blocking_ms_per_step'], d['stage_ms'])"
  RC_GRAPH=$g python tools/e2e_probe.py $wl 2>&1 | head -2
done; done
