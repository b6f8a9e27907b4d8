#!/bin/bash
# 2 x B200 (gpurun --gpus 2): the 2-rank parity tests and the default tiled bench line (driver's launch form) of the final code.
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -x -q 2>&1 | tail -4 | tee gpurun_out/r2s_pytest_multi_2gpu.txt
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 3 --no-cpu-baseline 2>/dev/null | tail -1 > gpurun_out/r2s_bench_tiled_2gpu.json
cut -c1-700 gpurun_out/r2s_bench_tiled_2gpu.json
