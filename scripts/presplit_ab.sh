#!/bin/bash
# A/B of the fp32 dense layers with and without pre-split TF32 planes (STINET_TC_PRESPLIT): error vs fp64 and time per shape
OUT=gpurun_out; mkdir -p $OUT
fmt='
import sys,json
for l in sys.stdin:
    try: d=json.loads(l)
    except Exception: print(l.strip()[:300]); continue
    if "FAILED" in d: print(d); continue
    print(d["prec"],d["op"],d["M"],d["N"],d["K"],"err=%.2e"%d["rel_err"],"ms=%.4f"%d["ms"],"TF=%.1f"%d["TFLOPs"],"det=%s"%d["deterministic"])
'
for thr in ${THRS:-0 160}; do
  echo "== STINET_TC_PRESPLIT=$thr"
  STINET_TC_PRESPLIT=$thr timeout 300 python scripts/gemm_check.py --precs fp32 --ops fwd,dgrad,wgrad 2>&1 | python -c "$fmt"
done | tee $OUT/presplit_ab.txt
