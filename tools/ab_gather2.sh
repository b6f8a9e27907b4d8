#!/bin/bash
# A/B: tensor-core gather vs scalar pipelined gather (stage_ms.gather), 4K and 1080p
for wl in living_room_4k teapot_1080p; do
  for mma in 0 1; do
    for t in 2 4 8; do
      RC_GATHER_MMA=$mma python tools/stage_times.py $wl --set gather_tiles=$t --tag "mma=$mma tiles=$t"
    done
  done
done
