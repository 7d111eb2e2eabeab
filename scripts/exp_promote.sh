# experiment: promotion interval x separate correction accumulator for the 3xTF32 GEMM (error vs fp64, time)
for P in 4 2 1; do for S in 0 1; do
echo "== promote=$P splitacc=$S"
STINET_TC_PROMOTE=$P STINET_TC_SPLITACC=$S timeout 300 python scripts/gemm_check.py --precs fp32 --ops fwd,wgrad 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    try: d=json.loads(l)
    except Exception: print(l.strip()[:300]); continue
    if 'FAILED' in d: print(d); continue
    if d['M']<1000 or d['M']==1000: continue
    print(d['op'],d['M'],d['N'],d['K'],'err=%.2e'%d['rel_err'],'ms=%.4f'%d['ms'],'TF=%.1f'%d['TFLOPs'])
"
done; done
