OUT=gpurun_out
for SMS in 148 140 132; do
STINET_TC_SMS=$SMS timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29$SMS bench.py --gpus 2 --steps 20 --warmup 5 --no-cached --no-profile > $OUT/bench_n2_sms${SMS}_r2_s.json 2> $OUT/bench_n2_sms${SMS}_r2_s.err; echo "sms $SMS exit $?"
python -c "import json;b=json.load(open('$OUT/bench_n2_sms${SMS}_r2_s.json'));print('N=2 sms $SMS', b['value'], b['ms_per_step'], b['e2e']['ms_per_step'])" || tail -5 $OUT/bench_n2_sms${SMS}_r2_s.err
done
NCCL_MAX_CTAS=8 timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29577 bench.py --gpus 2 --steps 20 --warmup 5 --no-cached --no-profile > $OUT/bench_n2_ctas8_r2_s.json 2> $OUT/bench_n2_ctas8_r2_s.err; python -c "import json;b=json.load(open('$OUT/bench_n2_ctas8_r2_s.json'));print('N=2 NCCL_MAX_CTAS=8', b['value'], b['ms_per_step'])"
STINET_TC_SMS=148 timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-cached > $OUT/bench_n1_r2_s.json 2>/dev/null; python -c "import json;b=json.load(open('$OUT/bench_n1_r2_s.json'));print('N=1', b['value'], b['ms_per_step'], b['roofline']['achieved'], b['roofline']['frac'])"
