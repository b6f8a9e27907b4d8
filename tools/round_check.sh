# One GPU-box call: GPU parity tests, smoke, then the round's profile capture (tools/profile_round.sh).
# Usage (from the repo root): bash tools/round_check.sh <tag>       outputs under gpurun_out/
tag=${1:-r1h}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/gpu_${tag}.txt 2>&1
( time timeout 900 python -m pytest tests -m gpu -q --maxfail=5 ) > gpurun_out/pytest_gpu_${tag}.log 2>&1
echo "pytest exit: $?" >> gpurun_out/pytest_gpu_${tag}.log
tail -5 gpurun_out/pytest_gpu_${tag}.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_${tag}.log 2>&1
echo "smoke exit: $?" >> gpurun_out/smoke_${tag}.log
tail -2 gpurun_out/smoke_${tag}.log
timeout 120 ./radiancecascade_b200/rc_headless --scene scenes/_extracted/teapot/teapot.obj --size 1920x1080 --frames 8 --warmup 3 --quiet > gpurun_out/headless_${tag}.json 2>&1
tail -1 gpurun_out/headless_${tag}.json
bash tools/profile_round.sh ${tag}
for f in gpurun_out/bench_${tag}_*.json; do echo "== $f"; cut -c1-400 $f; done
