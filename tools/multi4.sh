# 4 x B200 (gpurun --gpus 4): multi-view batch bench and tiled bench (strips, peer-memory exchange) of the final code.
tag=${1:-r1q}
mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 4 --steps 20 --warmup 3 2>/dev/null | tail -1 > gpurun_out/bench_${tag}_batch_4gpu.json
cut -c1-200 gpurun_out/bench_${tag}_batch_4gpu.json
python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 4 --mode tiled --exchange peer --steps 20 --warmup 3 2>/dev/null | tail -1 > gpurun_out/bench_${tag}_tiled_4gpu_peer.json
cut -c1-300 gpurun_out/bench_${tag}_tiled_4gpu_peer.json
