OUT=gpurun_out
timeout 600 python scripts/gemm_f16_check.py --regimes unit > $OUT/gemm_f16_r2_m.jsonl 2> $OUT/gemm_f16_r2_m.err; echo "check exit $?"; grep -c FAILED $OUT/gemm_f16_r2_m.jsonl
python - <<'PY'
import json
rows=[json.loads(l) for l in open('gpurun_out/gemm_f16_r2_m.jsonl')]
for r in rows:
    if r.get('passes')==3 and 'err' in r: print(f"{r['op']:6s} {r['M']:7d} {r['N']:5d} {r['K']:5d} err {r['err']:.1e} ms {r['ms']:.4f} TF {r['TFLOPs']:6.1f} det {r['deterministic']}")
PY
( timeout 600 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_model.py -m gpu -q --tb=short -p no:cacheprovider -x ) > $OUT/pytest_r2_m.log 2>&1; echo "pytest exit $?"; tail -3 $OUT/pytest_r2_m.log
timeout 300 python bench.py --no-cpu-baseline --no-cached > $OUT/bench_r2_m.json 2>$OUT/bench_r2_m.err; python -c "import json;b=json.load(open('gpurun_out/bench_r2_m.json'));print(b['ms_per_step'])"
