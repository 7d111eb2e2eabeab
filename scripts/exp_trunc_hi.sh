#!/bin/bash
# Experiment for the next GPU session: 3xTF32 with hi = trunc(x) left in place (only lo written by the split warps).
# Build BOTH libraries before gpurun:  make -C surface-texture-inpainting-net_b200/csrc && make -C surface-texture-inpainting-net_b200/csrc trunc
#   gpurun --timeout 400 -- 'bash scripts/exp_trunc_hi.sh'
# Reads: rel_err must stay < 1e-6 (if kind::tf32 ROUNDS its operands instead of truncating them the error jumps to ~1e-4
# and the experiment is over); then compare ms per shape.
OUT=gpurun_out; mkdir -p $OUT
LIBDIR=$PWD/surface-texture-inpainting-net_b200/stinet_b200
for lib in libstinet_b200.so libstinet_b200_trunc.so; do
  echo "== $lib"
  STINET_B200_LIB=$LIBDIR/$lib timeout 180 python scripts/gemm_check.py --precs fp32 --ops fwd,dgrad,wgrad 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    try: d=json.loads(l)
    except Exception: print(l.strip()[:200]); continue
    if 'FAILED' in d: print(d); continue
    if d['M']>1000: print(d['op'],d['M'],d['N'],d['K'],'err=%.2e'%d['rel_err'],'ms=%.4f'%d['ms'],'TF=%.1f'%d['TFLOPs'])
"
done | tee $OUT/exp_trunc_hi.txt
