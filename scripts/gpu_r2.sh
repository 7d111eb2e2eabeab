#!/bin/bash
# round-2 GPU session: smoke, parity suite, bench (+ per-entry-point table), launch list of one graph-replayed step.
#   gpurun --timeout 1500 -- 'bash scripts/gpu_r2.sh r2_b [tests|notests] [ncu|noncu]'
TAG=${1:-r2_x}
MODE=${2:-tests}
NCU=${3:-ncu}
OUT=gpurun_out
mkdir -p $OUT
python __graft_entry__.py > $OUT/smoke_$TAG.log 2>&1; echo "smoke exit $?"; tail -1 $OUT/smoke_$TAG.log
if [ "$MODE" = "tests" ]; then
  ( time timeout 900 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider ${PYTEST_EXTRA:-} ) > $OUT/pytest_$TAG.log 2>&1; echo "pytest exit $?"; tail -4 $OUT/pytest_$TAG.log; grep -E "^(FAILED|ERROR)" $OUT/pytest_$TAG.log | head -40
fi
timeout 400 python bench.py --kernels-out $OUT/kernels_$TAG.json ${BENCH_EXTRA:-} > $OUT/bench_$TAG.json 2> $OUT/bench_$TAG.err; echo "bench exit $?"
python - <<PY
import json
try:
    b=json.load(open("$OUT/bench_$TAG.json"))
    print("ms_per_step", b["ms_per_step"], "e2e", b["e2e"]["ms_per_step"] if b.get("e2e") else None, "launches", b["gpu_launches"])
except Exception as e:
    print("bench parse failed", e); print(open("$OUT/bench_$TAG.err").read()[-2000:])
PY
if [ "$NCU" = "ncu" ]; then
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
  --log-file $OUT/launches_$TAG.csv python bench.py --steps 1 --warmup 3 --no-e2e --no-profile --no-cpu-baseline \
  --profiler-range ${BENCH_EXTRA:-} > $OUT/ncu_launch_$TAG.log 2>&1; echo "ncu launches exit $?"
fi
ls -la $OUT | tail -8
