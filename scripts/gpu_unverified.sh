#!/bin/bash
# First GPU call of the next session: run everything that round 1 could only record as non-strict expectations
# (tests marked xfail(strict=False) because the GPU budget ran out) with --runxfail, so they pass or fail for real.
#   gpurun --timeout 300 -- 'bash scripts/gpu_unverified.sh'
OUT=gpurun_out; mkdir -p $OUT
timeout 250 python -m pytest tests/test_zz_gpu_singleconv.py tests/test_hierarchy.py -m gpu --runxfail -q --tb=short \
  -p no:cacheprovider > $OUT/unverified.log 2>&1
echo "exit $?"; tail -25 $OUT/unverified.log
