#!/bin/bash
# k_march_pool (block-local ray pool with refill of finished lanes) against k_march: bit-identity test, then stage times.
python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "experimental" 2>&1 | tail -3
for wl in living_room_4k teapot_1080p; do
  python tools/stage_times.py $wl --levels | cut -c1-420
  for t in 16 20 24 28; do python tools/stage_times.py $wl --levels --set march_pool=1 --set march_pool_thresh=$t | cut -c1-460; done
done
