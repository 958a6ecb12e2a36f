#!/usr/bin/env bash
# Round-2 evidence run on ONE B200:  gpurun --timeout 1800 -- 'bash tools/gpu_round2.sh [stage...]'
# Stages: tests bench peaks ncu_list ncu_scan ncu_batch   (default: all).  Every ncu pass runs only after
# the same command exited 0 without ncu in this call.
set -u
mkdir -p gpurun_out
STAGES=${*:-"tests bench peaks ncu_list ncu_scan ncu_batch"}
SHORT="python bench.py --steps 2 --warmup 3 --queries-per-step 4 --batch-iters 1 --no-cpu-baseline --no-c2"
BATCH="python tools/bench_configs.py --custom 3000000,384,4096,10,bf16 none"
for st in $STAGES; do
  case $st in
    tests)
      timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/summary.txt
      tail -5 gpurun_out/pytest_gpu.log ;;
    bench)
      timeout 900 python bench.py > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo "bench rc=$?" | tee -a gpurun_out/summary.txt
      cat gpurun_out/bench_n1.json
      timeout 300 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "ref rc=$?" | tee -a gpurun_out/summary.txt ;;
    peaks)
      timeout 300 python tools/measure_gemm_peaks.py > gpurun_out/gemm_peaks.json 2>&1; echo "peaks rc=$?" | tee -a gpurun_out/summary.txt
      cat gpurun_out/gemm_peaks.json ;;
    ncu_list)
      timeout 600 $SHORT > gpurun_out/plain_short.log 2>&1 &&
      timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"scan_topk|scan_mma_topk|batch_topk|finalize_batch|seed_threshold|prepare_queries|merge_topk|exchange_merge|fill_empty" -c 3000 --csv \
        --log-file gpurun_out/launches.csv $SHORT > gpurun_out/ncu_list.log 2>&1
      echo "ncu_list rc=$?" | tee -a gpurun_out/summary.txt ;;
    ncu_scan)
      timeout 600 $SHORT > gpurun_out/plain_short2.log 2>&1 &&
      timeout 900 ncu --set full --clock-control none --import-source on -k regex:"scan_topk|scan_mma_topk" -s 100 -c 2 \
        -f -o gpurun_out/scan_bf16_c5 $SHORT > gpurun_out/ncu_scan.log 2>&1
      echo "ncu_scan rc=$?" | tee -a gpurun_out/summary.txt ;;
    ncu_batch)
      # launches of batch_topk per search: seed, sample, main; warm-up 3 + timed: the 9th is a main pass
      timeout 300 $BATCH > gpurun_out/plain_batch.log 2>&1 &&
      timeout 900 ncu --set full --clock-control none --import-source on -k regex:batch_topk -s 8 -c 1 \
        -f -o gpurun_out/batch_bf16_main $BATCH > gpurun_out/ncu_batch.log 2>&1
      echo "ncu_batch rc=$?" | tee -a gpurun_out/summary.txt ;;
  esac
done
cat gpurun_out/summary.txt
