#!/bin/bash
show() { python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('%s ms/frame %.3f  stages %s gather_frac %.3f'%(sys.argv[1], d['ms_per_step'], {k:round(v,3) for k,v in d['stage_ms'].items()}, d['roofline_gather']['frac']))" "$1"; }
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -2
timeout 120 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | show teapot1080
timeout 120 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --workload living_room_4k 2>&1 | tail -1 | show lr4k
timeout 120 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --workload test_room_1080p 2>&1 | tail -1 | show testroom
