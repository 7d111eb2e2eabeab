#!/bin/bash
OUT=gpurun_out
timeout 400 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-cached --kernels-out $OUT/kernels_r2_z0.json > $OUT/bench_r2_z0.json 2> $OUT/bench_r2_z0.err; echo "exit $?"
python - <<'PY'
import json
b=json.load(open("gpurun_out/bench_r2_z0.json")); print(b["ms_per_step"], b["roofline"]["achieved"], b["roofline"]["frac"])
k=json.load(open("gpurun_out/kernels_r2_z0.json")); print(k["total_ms_per_step"])
for n in ['weight_planes_refresh','f16_amax','f16_split','f16_split_colsum','csr_build','edgeconv_hoist_bwd']:
    print(n, k["kernels"][n]["calls_per_step"], round(k["kernels"][n]["ms_per_step"],3))
PY
