#!/bin/bash
# ncu --set full of single dense-layer launches (source-level stall attribution): $1 = tag, rest = launch skips
OUT=gpurun_out; mkdir -p $OUT
TAG=$1; shift
for skip in "$@"; do
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:gemm_tc -s $skip -c 1 \
    -o $OUT/gemm_${TAG}_s$skip -f python scripts/gemm_prof.py fp32 > $OUT/ncu_gemm_${TAG}_s$skip.log 2>&1
  echo "skip $skip exit $?"; ls -la $OUT/gemm_${TAG}_s$skip.ncu-rep
done
