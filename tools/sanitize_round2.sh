# compute-sanitizer over the round-2 kernels (one B200).  Usage: gpurun -- bash tools/sanitize_round2.sh
out=gpurun_out/r2_compute_sanitizer.txt
: > $out
run() { echo "\$ $*" >> $out; "$@" 2>&1 | grep -v "^=========$" | tail -4 >> $out; }
run compute-sanitizer --tool memcheck --report-api-errors no --error-exitcode 9 python -c "import __graft_entry__ as g; g.smoke()"
run compute-sanitizer --tool memcheck --report-api-errors no --error-exitcode 9 ./radiancecascade_b200/rc_headless --scene scenes/_extracted/living_room/living_room.obj --size 333x205 --frames 3 --read composite --quiet
run compute-sanitizer --tool memcheck --report-api-errors no --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "tensor_core or set_tile or raster_clip or rgb48 or binned and not 3840"
run compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "tensor_core_gather_matches_scalar_gather and cube"
run compute-sanitizer --tool synccheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "tensor_core_gather_matches_scalar_gather and cube"
cat $out
