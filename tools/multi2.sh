# 2 x B200 (gpurun --gpus 2): tiled-frame parity test, tiled bench (peer-memory vs NCCL exchange), multi-view batch bench.
tag=${1:-r1m}
mkdir -p gpurun_out
python -m pytest tests/test_gpu_multi.py -m gpu -x -q 2>&1 | tail -5 | tee gpurun_out/pytest_multi2_${tag}.txt
for ex in peer nccl; do
  python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --mode tiled --exchange $ex --steps 20 --warmup 3 2>/dev/null | tail -1 > gpurun_out/bench_${tag}_tiled_2gpu_${ex}.json
  cut -c1-300 gpurun_out/bench_${tag}_tiled_2gpu_${ex}.json
done
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 20 --warmup 3 2>/dev/null | tail -1 > gpurun_out/bench_${tag}_batch_2gpu.json
cut -c1-400 gpurun_out/bench_${tag}_batch_2gpu.json
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --impl reference --gpus 2 --steps 2 --warmup 1 2>/dev/null | tail -1 | cut -c1-200
