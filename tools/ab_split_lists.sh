#!/bin/bash
# Split ray lists (k_split): bit-identity test, stage times with / without, and an ncu launch list of one 4K frame.
python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "split" 2>&1 | tail -3
for wl in living_room_4k test_room_1080p teapot_1080p; do
  for s in 1 0; do python tools/stage_times.py $wl --levels --set list_split=$s | cut -c1-420; done
done
RC_GRAPH=0 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/r2s_launches_living_room_4k.csv \
    python tools/stage_times.py living_room_4k --frames 1 > gpurun_out/r2s_launches.log 2>&1
python - <<'PY'
import csv, collections
rows = [r for r in csv.reader(open("gpurun_out/r2s_launches_living_room_4k.csv")) if len(r) > 10 and r[0].isdigit()]
# last frame only: take the last occurrence block (from the last k_gbuffer on)
names = [r[4] for r in rows]
start = max(i for i, n in enumerate(names) if "k_gbuffer" in n)
for r in rows[start:]:
    print(r[4][:60], r[-1])
PY
