#!/bin/bash
mkdir -p gpurun_out
N=$(nvidia-smi -L | wc -l)
echo "GPUs: $N"
run() { timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $1 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $1 --steps 10 --warmup 3 "${@:2}" 2>&1 | grep '^{' | tail -1; }
run $N --mode tiled | tee gpurun_out/bench_r1_tiled_${N}gpu_strips.json | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('tiled strips', d['n_gpus'], 'ms', round(d['ms_per_step'],3), 'render', round(d['render_ms_per_step'],3), 'G/s', round(d['value'],2), 'redundant', round(d['config']['redundant_rays'],3))"
if [ "$N" = "8" ]; then
run 8 --mode tiled --grid 4x2 | tee gpurun_out/bench_r1_tiled_8gpu_4x2.json | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('tiled 4x2', d['n_gpus'], 'ms', round(d['ms_per_step'],3), 'render', round(d['render_ms_per_step'],3), 'G/s', round(d['value'],2), 'redundant', round(d['config']['redundant_rays'],3))"
run 4 --mode tiled --grid 2x2 | tee gpurun_out/bench_r1_tiled_4gpu_2x2.json | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('tiled 2x2', d['n_gpus'], 'ms', round(d['ms_per_step'],3), 'render', round(d['render_ms_per_step'],3), 'G/s', round(d['value'],2), 'redundant', round(d['config']['redundant_rays'],3))"
run 8 | tee gpurun_out/bench_r1_batch_8gpu.json | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('batch', d['n_gpus'], 'ms', round(d['ms_per_step'],3), 'G/s', round(d['value'],2), 'e2e', round(d['e2e']['value'],2))"
fi
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --workload living_room_4k 2>&1 | grep '^{' | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('single lr4k ms', round(d['ms_per_step'],3), 'G/s', round(d['value'],2))"
