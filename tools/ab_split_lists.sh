#!/bin/bash
# Split ray lists (k_split): bit-identity test and stage times with / without.
python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "split or experimental" 2>&1 | tail -3
for wl in living_room_4k test_room_1080p teapot_1080p; do
  python tools/stage_times.py $wl --levels | cut -c1-420
  python tools/stage_times.py $wl --levels --set list_split=0 | cut -c1-420
done
