#!/bin/bash
OUT=gpurun_out/gemm_st2_r2_z.jsonl
: > $OUT
for E in "0 0" "1 0" "1 1"; do
  set -- $E
  STINET_TC_SCALED=$1 STINET_TC_STAGING2=$2 timeout 200 python scripts/gemm_f16_knobs.py | sed "s/^{/{\"cfg\": \"s$1_d$2\", /" >> $OUT 2>> gpurun_out/gemm_st2.err
done
tail -3 gpurun_out/gemm_st2.err
python - <<'PY'
import json
rows=[json.loads(l) for l in open("gpurun_out/gemm_st2_r2_z.jsonl")]
key=lambda r:(r["op"],r["M"],r["N"],r["K"])
d={}
for r in rows: d.setdefault(key(r),{})[r["cfg"]]=r
for k,v in d.items():
    print(k, " ".join("%s %.1f us (%.1e)" % (c, 1e3*v[c]["ms"], v[c]["err"]) for c in ("s0_d0","s1_d0","s1_d1") if c in v))
PY
timeout 300 python scripts/gemm_f16_check.py --regimes unit,tails > gpurun_out/gemm_f16_check_r2_z.jsonl 2>&1
python - <<'PY'
import json
worst={}
for l in open("gpurun_out/gemm_f16_check_r2_z.jsonl"):
    try: r=json.loads(l)
    except Exception: print("??", l[:200]); continue
    if "FAILED" in r: print(r); continue
    k=(r["op"],r["passes"],r["regime"]); worst[k]=max(worst.get(k,0), r["err"])
    if not r.get("deterministic", True): print("nondeterministic", r)
print(worst)
PY
timeout 600 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_precision_sizes.py -m gpu -q --tb=short -p no:cacheprovider -x 2>&1 | tail -3
for E in "0 0" "1 1"; do
set -- $E
STINET_TC_SCALED=$1 STINET_TC_STAGING2=$2 timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-profile --no-cached --no-e2e > gpurun_out/bench_st2_$1$2.json 2> gpurun_out/bench_st2_$1$2.err
python -c "
import json; b=json.load(open('gpurun_out/bench_st2_$1$2.json')); print('scaled/staging2 $1 $2', b['ms_per_step'])" || tail -5 gpurun_out/bench_st2_$1$2.err
done
