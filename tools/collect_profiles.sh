# Copy one round-check capture (gpurun_out/*_<tag>_*) into profiles/ under <prefix> (tracked, judged).
# Usage: bash tools/collect_profiles.sh r1h r1_h
tag=$1; pre=$2
for wl in teapot_1080p living_room_4k; do
  cp gpurun_out/launches_${tag}_${wl}.csv profiles/${pre}_launches_${wl}.csv
  python tools/ncu_summary.py gpurun_out/raw_${tag}_${wl}.csv > profiles/${pre}_ncu_full_summary_${wl}.csv
done
python tools/make_traffic.py ${tag} | sed "s#profiles/${tag}_#profiles/${pre}_#" > profiles/${pre}_traffic.json
for b in teapot1080p living_room4k test_room1080p sonic8k cube512 reference; do
  [ -s gpurun_out/bench_${tag}_${b}.json ] && cp gpurun_out/bench_${tag}_${b}.json profiles/${pre}_bench_${b}.json
done
[ -s gpurun_out/pytest_gpu_${tag}.log ] && tail -8 gpurun_out/pytest_gpu_${tag}.log > profiles/${pre}_pytest_gpu.txt
[ -s gpurun_out/headless_${tag}.json ] && cp gpurun_out/headless_${tag}.json profiles/${pre}_rc_headless_teapot1080p.json
ls profiles | grep "^${pre}_"
