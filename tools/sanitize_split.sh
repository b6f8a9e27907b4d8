# compute-sanitizer over the late-round-2 kernels (k_split, k_march with split lists, k_march_pool) on one B200.
# Usage: gpurun -- bash tools/sanitize_split.sh
out=gpurun_out/r2s_compute_sanitizer.txt
: > $out
run() { echo "\$ $*" >> $out; timeout 600 "$@" 2>&1 | grep -v "^=========$" | tail -4 >> $out; }
run compute-sanitizer --tool memcheck --report-api-errors no --error-exitcode 9 python -c "import __graft_entry__ as g; g.smoke()"
run compute-sanitizer --tool memcheck --report-api-errors no --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "split_ray_lists and (cube or test_room)"
run compute-sanitizer --tool memcheck --report-api-errors no --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "experimental and cube"
run compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "split_ray_lists and cube"
run compute-sanitizer --tool synccheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "experimental and cube"
cat $out
