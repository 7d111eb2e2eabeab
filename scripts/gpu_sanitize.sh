#!/bin/bash
# compute-sanitizer over the smoke workload (one tiny forward + backward of the flagship network through the C ABI) and over
# the per-kernel parity tests that exercise every kernel family once: memcheck (out-of-bounds / misaligned accesses) and
# racecheck (shared-memory hazards).  Summaries go to gpurun_out/ and are copied to profiles/.
#   gpurun --timeout 1500 -- 'bash scripts/gpu_sanitize.sh r2_q'
TAG=${1:-r2_x}
OUT=gpurun_out
mkdir -p $OUT
SMOKE='import __graft_entry__ as g; g.smoke()'
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 --log-file $OUT/sanitizer_memcheck_smoke_$TAG.log python -c "$SMOKE" > $OUT/sanitizer_memcheck_smoke_$TAG.out 2>&1; echo "memcheck smoke exit $?"
tail -3 $OUT/sanitizer_memcheck_smoke_$TAG.log
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 --log-file $OUT/sanitizer_memcheck_kernels_$TAG.log python -m pytest tests/test_gpu_kernels.py -m gpu -q -x -p no:cacheprovider -k "csr or edge_message_plane or pool_unpool or batch_norm or graph_norm or head_tanh or masked_l1 or unpool_concat or instance_norm or linear_fwd" > $OUT/sanitizer_memcheck_kernels_$TAG.out 2>&1; echo "memcheck kernels exit $?"
tail -3 $OUT/sanitizer_memcheck_kernels_$TAG.log; tail -2 $OUT/sanitizer_memcheck_kernels_$TAG.out
timeout 600 compute-sanitizer --tool racecheck --error-exitcode 9 --log-file $OUT/sanitizer_racecheck_smoke_$TAG.log python -c "$SMOKE" > $OUT/sanitizer_racecheck_smoke_$TAG.out 2>&1; echo "racecheck smoke exit $?"
tail -3 $OUT/sanitizer_racecheck_smoke_$TAG.log
