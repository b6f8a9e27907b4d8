# Quick check of the bench lines (one B200): default workload, 4K workload, CPU arm under torchrun.
tag=${1:-r1n}
mkdir -p gpurun_out
python bench.py > gpurun_out/bench_${tag}_teapot1080p.json 2> gpurun_out/bench_${tag}_teapot1080p.err; tail -c 300 gpurun_out/bench_${tag}_teapot1080p.err
python bench.py --workload living_room_4k --cpu-sample-div 2 > gpurun_out/bench_${tag}_living_room4k.json 2>/dev/null
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --impl reference --gpus 2 --steps 3 --warmup 1 2>/dev/null | tail -1 > gpurun_out/bench_${tag}_reference_torchrun2.json
for f in gpurun_out/bench_${tag}_*.json; do echo "== $f"; python -c "
import json,sys
d=json.loads(open('$f').read().strip().splitlines()[-1])
print({k: d[k] for k in ('value','ms_per_step','n_gpus') if k in d}, d.get('e2e'), d.get('cpu_baseline'), (d.get('roofline') or {}).get('traffic'))"; done
