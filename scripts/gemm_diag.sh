#!/bin/bash
# GEMM diagnosis on the box: where the warp roles of the persistent kernel wait (debug build), and time / error per mode
OUT=gpurun_out; mkdir -p $OUT
# the debug library is built here (before gpurun) with: make -C surface-texture-inpainting-net_b200/csrc debug
STINET_B200_LIB=$PWD/surface-texture-inpainting-net_b200/stinet_b200/libstinet_b200_dbg.so timeout 200 python scripts/gemm_waits.py fp32 tf32 bf16 > $OUT/gemm_waits.txt 2>&1
cat $OUT/gemm_waits.txt
timeout 400 python scripts/gemm_check.py --precs fp32,tf32,bf16,bf16x3 --ops fwd,dgrad,wgrad 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    try: d=json.loads(l)
    except Exception: print(l.strip()[:300]); continue
    if 'FAILED' in d: print(d); continue
    if d['M']<=1000: continue
    print(d['prec'],d['op'],d['M'],d['N'],d['K'],'err=%.2e'%d['rel_err'],'ms=%.4f'%d['ms'],'TF=%.1f'%d['TFLOPs'])
" | tee $OUT/gemm_check.txt
