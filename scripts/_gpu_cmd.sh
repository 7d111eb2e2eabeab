#!/bin/bash
OUT=gpurun_out; mkdir -p $OUT
N=${1:-4}
timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29519 bench.py --gpus $N --steps 20 --warmup 5 --no-cpu-baseline --no-cached > $OUT/bench_r2_zz_n$N.json 2> $OUT/bench_r2_zz_n$N.err
echo "n$N exit $?"; python -c "
import json; b=json.load(open('$OUT/bench_r2_zz_n$N.json')); print(b['n_gpus'], b['ms_per_step'], b['value'], b['e2e']['value'], b['clocks'])" || tail -20 $OUT/bench_r2_zz_n$N.err
