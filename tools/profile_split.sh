#!/bin/bash
# ncu --set full of the five march launches of one 4K frame with split lists (default) and without, and of k_split.
bash tools/ncu_one.sh r2s_march_split living_room_4k k_march 15 5
RC_LIST_SPLIT=0 bash tools/ncu_one.sh r2s_march_nosplit living_room_4k k_march 15 5
bash tools/ncu_one.sh r2s_ksplit living_room_4k k_split 3 1
rm -f gpurun_out/prof_r2s_march_nosplit.ncu-rep
