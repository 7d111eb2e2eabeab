#!/bin/bash
# round-2 ncu session: launch list of one graph-replayed step + `--set full` captures of the dominant kernels (eager launches:
# forward pass, then backward pass), exported to raw CSV on the box (the .ncu-rep files exceed the transfer cap).
#   gpurun --timeout 1500 -- 'bash scripts/gpu_ncu_r2.sh r2_p'
TAG=${1:-r2_x}
OUT=gpurun_out
mkdir -p $OUT
timeout 300 python bench.py --no-cpu-baseline --no-cached --kernels-out $OUT/kernels_$TAG.json > $OUT/bench_$TAG.json 2> $OUT/bench_$TAG.err; echo "bench exit $?"
python -c "import json;b=json.load(open('$OUT/bench_$TAG.json'));print('ms_per_step', b['ms_per_step'])"
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
  --log-file $OUT/launches_$TAG.csv python bench.py --steps 1 --warmup 3 --no-e2e --no-profile --no-cpu-baseline \
  --profiler-range > $OUT/ncu_launch_$TAG.log 2>&1; echo "ncu launches exit $?"
timeout 600 ncu --set full --clock-control none --import-source on --profile-from-start off \
  -k regex:'edge_message_fwd_planes|pool_max_fwd|seg_colreduce|segnorm_fused_fwd|segnorm_slice_apply|row_gather|gemm_tc|wplanes|split_f16|csr_' -c ${NCU_COUNT:-40} \
  -o $OUT/prof_fwd_$TAG -f python bench.py --steps 1 --warmup 3 --no-e2e --no-profile --no-cpu-baseline --no-graph \
  --profiler-range > $OUT/ncu_full_fwd_$TAG.log 2>&1; echo "ncu full fwd exit $?"
timeout 600 ncu --set full --clock-control none --import-source on --profile-from-start off \
  --kernel-name-base mangled -k regex:"bwd|split_colsum|cluster_sum|splitk_reduce|gemm_tc_kernelILi128ELb0ELb1|gemm_tc_kernelILi128ELb1ELb1" -c ${NCU_COUNT:-40} \
  -o $OUT/prof_bwd_$TAG -f python bench.py --steps 1 --warmup 3 --no-e2e --no-profile --no-cpu-baseline --no-graph \
  --profiler-range > $OUT/ncu_full_bwd_$TAG.log 2>&1; echo "ncu full bwd exit $?"
for f in fwd bwd; do
  ncu -i $OUT/prof_${f}_$TAG.ncu-rep --page raw --csv > $OUT/prof_${f}_${TAG}_raw.csv 2>/dev/null
  sz=$(stat -c %s $OUT/prof_${f}_$TAG.ncu-rep 2>/dev/null || echo 0)
  if [ "$sz" -gt 15000000 ]; then rm -f $OUT/prof_${f}_$TAG.ncu-rep; echo "dropped prof_${f} rep ($sz bytes), kept raw csv"; fi
done
du -sh $OUT; ls -la $OUT | tail -8
