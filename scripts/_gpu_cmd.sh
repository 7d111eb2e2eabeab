#!/bin/bash
OUT=gpurun_out
python __graft_entry__.py 2>&1 | tail -1
( timeout 900 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider ) > $OUT/pytest_r2_zz.log 2>&1; echo "pytest exit $?"; grep -E "passed|failed" $OUT/pytest_r2_zz.log | tail -1; grep -E "^(FAILED|ERROR)" $OUT/pytest_r2_zz.log | head
timeout 400 python bench.py --steps 20 --warmup 5 --kernels-out $OUT/kernels_r2_zz.json > $OUT/bench_r2_zz.json 2> $OUT/bench_r2_zz.err; echo "bench exit $?"
python -c "
import json; b=json.load(open('$OUT/bench_r2_zz.json')); print('bench', b['ms_per_step'], b['value'], b['e2e']['value'], b['roofline']['achieved'], b['roofline']['frac'], b['gpu_launches'])"
