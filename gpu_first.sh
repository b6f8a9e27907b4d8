#!/bin/bash
# first GPU contact: tests + a timing peek
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv | tee gpurun_out/smi.txt
nproc | tee gpurun_out/nproc.txt
python -m pytest tests -m gpu -x -q 2>&1 | tail -40 | tee gpurun_out/pytest_gpu.txt
