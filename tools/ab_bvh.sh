#!/bin/bash
# A/B: BVH leaf size and SAH leaf termination (RC_BVH_LEAF, RC_BVH_NODE_COST; any tree gives bit-identical closest hits)
for wl in living_room_4k teapot_1080p; do
  for leaf in 4 2 3 6; do echo -n "leaf=$leaf "; RC_BVH_LEAF=$leaf python tools/stage_times.py $wl --frames 8 | cut -c1-250; done
  for nc in 0.5 1.0 2.0; do echo -n "leaf=4 node_cost=$nc "; RC_BVH_NODE_COST=$nc python tools/stage_times.py $wl --frames 8 | cut -c1-250; done
done
