#!/bin/bash
show() { python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('%s ms/frame %.3f  gbuf %.3f march %.3f levels %s'%(sys.argv[1], d['ms_per_step'], d['stage_ms']['gbuffer'], d['stage_ms']['march'], [round(x,3) for x in d['level_ms']]))" "$1"; }
for cfg in "4 0" "2 0" "3 0" "6 0" "7 1.0" "7 1.5" "7 2.0" "4 1.0" "4 1.5"; do
  set -- $cfg
  echo "== leaf=$1 node_cost=$2"
  RC_BVH_LEAF=$1 RC_BVH_NODE_COST=$2 timeout 120 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | show teapot1080
  RC_BVH_LEAF=$1 RC_BVH_NODE_COST=$2 timeout 120 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --workload living_room_4k 2>&1 | tail -1 | show lr4k
done
RC_BVH_LEAF=7 RC_BVH_NODE_COST=1.5 timeout 300 python -m pytest tests -m gpu -x -q -k "closest or cascades" 2>&1 | tail -2
