OUT=gpurun_out
( timeout 900 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider ) > $OUT/pytest_r2_q.log 2>&1; echo "pytest exit $?"; grep -E "passed|failed" $OUT/pytest_r2_q.log | tail -1; grep -E "^(FAILED|ERROR)" $OUT/pytest_r2_q.log | head
timeout 300 python bench.py --no-cpu-baseline --no-cached > $OUT/bench_r2_q.json 2>$OUT/bench_r2_q.err; python -c "import json;b=json.load(open('gpurun_out/bench_r2_q.json'));print('ms', b['ms_per_step'])"
bash scripts/gpu_sanitize.sh r2_q
