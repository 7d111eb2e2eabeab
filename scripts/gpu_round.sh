#!/bin/bash
# One consolidated GPU-box session: smoke, [parity suite], bench (default workload), launch list, ncu --set full of the
# top kernels (summarised to CSV on the box: gpurun_out/ is capped at 64 MiB, the .ncu-rep files are dropped if large).
#   gpurun --timeout 1500 -- 'bash scripts/gpu_round.sh r1_f [tests|notests] [ncu|noncu]'
TAG=${1:-r1_x}
MODE=${2:-tests}
NCU=${3:-ncu}
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $OUT/smi_$TAG.txt 2>&1
python __graft_entry__.py > $OUT/smoke_$TAG.log 2>&1; echo "smoke exit $?"; tail -1 $OUT/smoke_$TAG.log
if [ "$MODE" = "tests" ]; then
  ( time timeout 900 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider ${PYTEST_EXTRA:-} ) > $OUT/pytest_$TAG.log 2>&1; echo "pytest exit $?"; tail -5 $OUT/pytest_$TAG.log; grep -E "^(FAILED|ERROR)" $OUT/pytest_$TAG.log | head -40
fi
timeout 400 python bench.py --kernels-out $OUT/kernels_$TAG.json > $OUT/bench_$TAG.json 2> $OUT/bench_$TAG.err; echo "bench exit $?"
head -c 400 $OUT/bench_$TAG.json; echo
[ -n "$SKIP_BF16" ] || timeout 300 python bench.py --dtype bf16 --no-cpu-baseline --kernels-out $OUT/kernels_bf16_$TAG.json > $OUT/bench_bf16_$TAG.json 2>> $OUT/bench_$TAG.err; echo "bench bf16 exit $?"
head -c 400 $OUT/bench_bf16_$TAG.json; echo
if [ "$MODE" = "tests" ]; then
  timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > $OUT/bench_ref_$TAG.json 2>> $OUT/bench_$TAG.err; echo "ref exit $?"
fi
if [ "$NCU" = "ncu" ]; then
# launch list of ONE timed step (graph replay: ncu sees the graph's kernel nodes)
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
  --log-file $OUT/launches_$TAG.csv python bench.py --steps 1 --warmup 3 --no-e2e --no-profile --no-cpu-baseline \
  --profiler-range > $OUT/ncu_launch_$TAG.log 2>&1; echo "ncu launches exit $?"
# full capture of the dominant kernels, eager launches: pass 1 = start of forward (level 0/1 shapes), pass 2 = start of
# backward (output blocks on level 0 come first)
timeout 500 ncu --set full --clock-control none --import-source on --profile-from-start off \
  -k regex:'edge_message_fwd|pool_max_fwd|seg_colreduce|segnorm_apply|segnorm_fused_fwd|segnorm_slice_apply|row_gather|gemm_tc|csr_' -c ${NCU_COUNT:-24} \
  -o $OUT/prof_fwd_$TAG -f python bench.py --steps 1 --warmup 3 --no-e2e --no-profile --no-cpu-baseline --no-graph \
  --profiler-range > $OUT/ncu_full_fwd_$TAG.log 2>&1; echo "ncu full fwd exit $?"
timeout 500 ncu --set full --clock-control none --import-source on --profile-from-start off \
  --kernel-name-base mangled -k regex:"bwd|colsum_partial|cluster_sum|splitk_reduce|gemm_tc_kernelILi128ELb0ELb1|gemm_tc_kernelILi128ELb1ELb1" -c ${NCU_COUNT:-24} \
  -o $OUT/prof_bwd_$TAG -f python bench.py --steps 1 --warmup 3 --no-e2e --no-profile --no-cpu-baseline --no-graph \
  --profiler-range > $OUT/ncu_full_bwd_$TAG.log 2>&1; echo "ncu full bwd exit $?"
for f in fwd bwd; do
  ncu -i $OUT/prof_${f}_$TAG.ncu-rep --page raw --csv > $OUT/prof_${f}_${TAG}_raw.csv 2>/dev/null
  sz=$(stat -c %s $OUT/prof_${f}_$TAG.ncu-rep 2>/dev/null || echo 0)
  if [ "$sz" -gt 20000000 ]; then rm -f $OUT/prof_${f}_$TAG.ncu-rep; echo "dropped prof_${f} rep ($sz bytes), kept raw csv"; fi
done
fi
du -sh $OUT; ls -la $OUT | tail -24
