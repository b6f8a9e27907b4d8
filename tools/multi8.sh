#!/bin/bash
# 8-GPU box: multi-GPU parity tests (2/4/8 ranks), then the tiled bench at N = 8, 4, 2 with recomputed and exchanged halos.
# Usage: gpurun --gpus 8 -- bash tools/multi8.sh <tag>
tag=${1:-r2m}
timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -x -q 2>&1 | tail -4 | tee gpurun_out/${tag}_pytest_multi.txt
port=29600
for n in 8 4 2; do
  for halo in recompute exchange; do
    port=$((port+1))
    python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $port bench.py --gpus $n --steps 10 --warmup 3 \
        --no-batch-extra --halo $halo > gpurun_out/${tag}_tiled_${n}gpu_${halo}.json 2> gpurun_out/${tag}_tiled_${n}gpu_${halo}.err
    python - <<PY
import json
try:
    j=json.loads(open("gpurun_out/${tag}_tiled_${n}gpu_${halo}.json").read().strip().splitlines()[-1])
    print("$n $halo", "ms", round(j["ms_per_step"],4), "render", round(j["render_ms_per_step"],4), "mean-rank", round(j["mean_rank_ms_per_step"],4), "redundant", round(j["config"]["redundant_rays"],4), "e2e", round(j["e2e"]["ms_per_step"],4))
except Exception as e:
    print("$n $halo FAILED", e); print(open("gpurun_out/${tag}_tiled_${n}gpu_${halo}.err").read()[-1500:])
PY
  done
done
