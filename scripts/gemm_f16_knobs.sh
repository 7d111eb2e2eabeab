#!/bin/bash
# sweep of the planes-GEMM knobs: timing with the product library, wait counters with the debug library
OUT=gpurun_out/gemm_knobs_${1:-r2}.jsonl
: > $OUT
for P in 2 4 8 100000; do for C in 0 1; do
  STINET_TC_PROMOTE16=$P STINET_TC_CORR_ONCE=$C timeout 200 python scripts/gemm_f16_knobs.py >> $OUT 2>> gpurun_out/gemm_knobs.err
done; done
for P in 2 100000; do for C in 0 1; do
  STINET_B200_LIB=$PWD/surface-texture-inpainting-net_b200/stinet_b200/libstinet_b200_dbg.so STINET_TC_PROMOTE16=$P STINET_TC_CORR_ONCE=$C timeout 200 python scripts/gemm_f16_knobs.py >> $OUT 2>> gpurun_out/gemm_knobs.err
done; done
wc -l $OUT; tail -3 gpurun_out/gemm_knobs.err
