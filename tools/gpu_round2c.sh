#!/usr/bin/env bash
# Third part of round 2: launch list of the short bench command and ncu --set full of the several-queries scan
# kernels (each ncu pass only after the same command exited 0 without ncu).
#   gpurun --timeout 900 -- 'bash tools/gpu_round2c.sh'
set -u
mkdir -p gpurun_out
if [ "${1:-}" != "traffic" ]; then
BENCH_SHORT="python bench.py --steps 2 --warmup 3 --queries-per-step 4 --batch-iters 1 --no-cpu-baseline --no-c2"
timeout 300 $BENCH_SHORT > gpurun_out/plain_short.log 2>&1 &&
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv \
  -k regex:"scan_|batch_topk|finalize_batch|seed_threshold|merge_topk|prepare_queries|exchange_merge" \
  --log-file gpurun_out/launches_bench_short_third_part.csv $BENCH_SHORT > gpurun_out/ncu_list.log 2>&1
echo "ncu_list rc=$?"
PVDB_CASES=1,3 timeout 300 python tools/bench_exact_batch.py > gpurun_out/plain_exact.log 2>&1 &&
PVDB_CASES=1,3 timeout 600 ncu --set full --clock-control none --import-source on -k regex:"scan_multi_topk|scan_mma_topk" \
  -s 10 -c 3 -f -o gpurun_out/scan_several_queries python tools/bench_exact_batch.py > gpurun_out/ncu_exact.log 2>&1
echo "ncu_full rc=$?"
fi
# final-tree captures of the two single-query scan kernels (profiles/traffic.json is regenerated from these)
if [ "${1:-}" = "traffic" ]; then
  C5="python bench.py --steps 2 --warmup 3 --queries-per-step 4 --no-batch --no-cpu-baseline --no-c2"
  C2="python bench.py --workload c2 --steps 2 --warmup 3 --queries-per-step 4 --no-batch --no-cpu-baseline"
  timeout 300 $C5 > gpurun_out/plain_c5.log 2>&1 &&
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:scan_mma_topk -s 20 -c 2 -f \
    -o gpurun_out/scan_mma_c5_final $C5 > gpurun_out/ncu_c5.log 2>&1; echo "ncu c5 rc=$?"
  timeout 300 $C2 > gpurun_out/plain_c2.log 2>&1 &&
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:scan_topk_kernel -s 20 -c 2 -f \
    -o gpurun_out/scan_f32_c2_final $C2 > gpurun_out/ncu_c2.log 2>&1; echo "ncu c2 rc=$?"
fi
