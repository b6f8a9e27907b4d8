#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
show() { python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('%s ms/frame %.3f  stages %s  levels %s'%(sys.argv[1], d['ms_per_step'], {k:round(v,3) for k,v in d['stage_ms'].items()}, [round(x,3) for x in d['level_ms']]))" "$1"; }
for cfg in "1 128 DDDDDD" "0 128 DDDDDD" "1 256 DDDDDD" "1 64 DDDDDD" "1 128 LLLLLL" "1 128 PPDDDD" "1 128 PDDDDD"; do
  set -- $cfg
  echo "== pdl=$1 block=$2 map=$3"
  RC_MARCH_PDL=$1 RC_MARCH_BLOCK=$2 RC_MARCH_MAP=$3 timeout 120 python bench.py --steps 8 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | show teapot1080
done
for cfg in "1 128 DDDDDD" "0 128 DDDDDD" "1 128 PPDDDD" "1 128 PPPDDD" "1 256 PPDDDD"; do
  set -- $cfg
  echo "== 4K pdl=$1 block=$2 map=$3"
  RC_MARCH_PDL=$1 RC_MARCH_BLOCK=$2 RC_MARCH_MAP=$3 timeout 120 python bench.py --steps 4 --warmup 3 --no-cpu-baseline --workload living_room_4k 2>&1 | tail -1 | show lr4k
done
