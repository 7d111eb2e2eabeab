#!/bin/bash
# A/B of the two side streams (structure build, weight gradients) on cfg2: parity tests first, then one bench line per setting.
#   gpurun --timeout 1200 -- 'bash scripts/gpu_streams.sh r2_u'
TAG=${1:-r2_u}
OUT=gpurun_out
mkdir -p $OUT
( time timeout 600 python -m pytest tests/test_gpu_streams.py tests/test_gpu_engine.py tests/test_zz_gpu_singleconv.py -m gpu -q --tb=short -p no:cacheprovider ) > $OUT/pytest_streams_$TAG.log 2>&1
echo "pytest exit $?"; tail -4 $OUT/pytest_streams_$TAG.log; grep -E "^(FAILED|ERROR)" $OUT/pytest_streams_$TAG.log | head -20
for cfg in "0 0" "1 0" "0 1" "0 2" "1 2" "1 1"; do
  set -- $cfg
  STINET_STRUCT_SIDE_STREAM=$1 STINET_WGRAD_SIDE_STREAM=$2 timeout 300 python bench.py --steps 20 --warmup 5 \
    --no-cpu-baseline --no-profile --no-cached > $OUT/bench_${TAG}_s$1_w$2.json 2> $OUT/bench_${TAG}_s$1_w$2.err
  echo "struct=$1 wgrad=$2 exit $?"
  python - <<PY
import json
try:
    b=json.load(open("$OUT/bench_${TAG}_s$1_w$2.json"))
    print("   ms_per_step", b["ms_per_step"], "e2e", (b.get("e2e") or {}).get("ms_per_step"), "launches", b["gpu_launches"])
except Exception as e:
    print("   parse failed", e); print(open("$OUT/bench_${TAG}_s$1_w$2.err").read()[-1500:])
PY
done
