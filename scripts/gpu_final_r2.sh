#!/bin/bash
# final measurement session of round 2 on one B200: every BASELINE config through bench.py (the lines land in gpurun_out/
# and are copied to profiles/), the reference arm on the CPU-runnable configs, smoke and the parity suite.
TAG=${1:-r2_z}
OUT=gpurun_out
mkdir -p $OUT
python __graft_entry__.py > $OUT/smoke_$TAG.log 2>&1; echo "smoke exit $?"; tail -1 $OUT/smoke_$TAG.log
( timeout 1200 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider ) > $OUT/pytest_$TAG.log 2>&1; echo "pytest exit $?"; grep -E "passed|failed" $OUT/pytest_$TAG.log | tail -1
timeout 400 python bench.py --steps 20 --warmup 5 --kernels-out $OUT/kernels_$TAG.json > $OUT/bench_$TAG.json 2> $OUT/bench_$TAG.err; echo "cfg2 exit $?"
timeout 300 python bench.py --dtype f16 --steps 20 --warmup 5 --no-cpu-baseline > $OUT/bench_cfg2_f16_$TAG.json 2>> $OUT/bench_$TAG.err; echo "cfg2 f16 exit $?"
for W in cfg1 cfg3; do
  timeout 400 python bench.py --workload $W --steps 20 --warmup 5 --kernels-out $OUT/kernels_${W}_$TAG.json > $OUT/bench_${W}_$TAG.json 2>> $OUT/bench_$TAG.err; echo "$W exit $?"
done
timeout 300 python bench.py --workload cfg3 --dtype f16 --steps 20 --warmup 5 --no-cpu-baseline > $OUT/bench_cfg3_f16_$TAG.json 2>> $OUT/bench_$TAG.err; echo "cfg3 f16 exit $?"
for D in fp32 f16; do
  timeout 900 python bench.py --workload cfg5 --dtype $D --steps 5 --no-cached --kernels-out $OUT/kernels_cfg5_${D}_$TAG.json > $OUT/bench_cfg5_${D}_$TAG.json 2>> $OUT/bench_$TAG.err; echo "cfg5 $D exit $?"
done
timeout 400 python bench.py --impl reference --steps 3 --warmup 1 > $OUT/bench_ref_cfg2_$TAG.json 2>> $OUT/bench_$TAG.err; echo "ref cfg2 exit $?"
timeout 300 python bench.py --workload cfg1 --impl reference --steps 5 --warmup 2 > $OUT/bench_ref_cfg1_$TAG.json 2>> $OUT/bench_$TAG.err; echo "ref cfg1 exit $?"
python - <<PY
import json, glob
for f in sorted(glob.glob("$OUT/bench_*$TAG.json")):
    try:
        b = json.load(open(f))
        print(f.split("/")[-1], "value %.4g" % b["value"], "ms %.3f" % b["ms_per_step"], "e2e %.4g" % (b.get("e2e") or {}).get("value", 0),
              "cpu", (b.get("cpu_baseline") or {}).get("value"), (b.get("roofline") or {}).get("kernel", "")[:30], (b.get("roofline") or {}).get("frac"))
    except Exception as e:
        print(f, "unreadable", e)
PY
