OUT=gpurun_out
for W in cfg3 cfg1; do
  timeout 400 python bench.py --workload $W --kernels-out $OUT/kernels_${W}_r2_i.json > $OUT/bench_${W}_r2_i.json 2> $OUT/bench_${W}_r2_i.err; echo "$W exit $?"
  python -c "import json;b=json.load(open('$OUT/bench_${W}_r2_i.json'));print('$W', b['metric'], b['value'], b['ms_per_step'], b['e2e']['value'] if b.get('e2e') else None, b['cpu_baseline'])"
done
timeout 600 python bench.py --workload cfg3 --dtype f16 --no-cpu-baseline > $OUT/bench_cfg3_f16_r2_i.json 2> $OUT/bench_cfg3_f16_r2_i.err; python -c "import json;b=json.load(open('$OUT/bench_cfg3_f16_r2_i.json'));print('cfg3 f16', b['value'], b['ms_per_step'])"
for D in fp32 f16; do
  timeout 900 python bench.py --workload cfg5 --dtype $D --steps 5 --no-cached --kernels-out $OUT/kernels_cfg5_${D}_r2_i.json > $OUT/bench_cfg5_${D}_r2_i.json 2> $OUT/bench_cfg5_${D}_r2_i.err; echo "cfg5 $D exit $?"
  python -c "import json;b=json.load(open('$OUT/bench_cfg5_${D}_r2_i.json'));print('cfg5 $D', b['value'], b['ms_per_step'], b['e2e']['value'] if b.get('e2e') else None, b['cpu_baseline'])"
done
timeout 300 python bench.py --workload cfg2 --dtype f16 --no-cpu-baseline --no-cached > $OUT/bench_cfg2_f16_r2_i.json 2> $OUT/bench_cfg2_f16_r2_i.err; python -c "import json;b=json.load(open('$OUT/bench_cfg2_f16_r2_i.json'));print('cfg2 f16', b['value'], b['ms_per_step'])"
timeout 600 python bench.py --workload cfg1 --impl reference --steps 5 --warmup 2 > $OUT/bench_ref_cfg1_r2_i.json 2> $OUT/bench_ref_cfg1_r2_i.err; head -c 400 $OUT/bench_ref_cfg1_r2_i.json
