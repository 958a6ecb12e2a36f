#!/usr/bin/env bash
# Run on a GPU box through gpurun:  gpurun --timeout 1800 -- 'bash tools/gpu_check.sh [stage...]'
# Stages: smoke tests bench ref ncu_list ncu_full   (default: all)
set -u
mkdir -p gpurun_out
STAGES=${*:-"smoke tests bench ref ncu_list ncu_full"}
BENCH_SHORT="python bench.py --steps 2 --warmup 3 --queries-per-step 4 --no-cpu-baseline"
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.total --format=csv > gpurun_out/gpu.csv 2>&1
for st in $STAGES; do
  case $st in
    smoke)
      timeout 600 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" | tee -a gpurun_out/summary.txt ;;
    tests)
      timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/summary.txt
      tail -15 gpurun_out/pytest_gpu.log ;;
    bench)
      timeout 900 python bench.py > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo "bench rc=$?" | tee -a gpurun_out/summary.txt
      cat gpurun_out/bench_n1.json ;;
    ref)
      timeout 900 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "ref rc=$?" | tee -a gpurun_out/summary.txt
      cat gpurun_out/bench_ref.json ;;
    ncu_list)
      timeout 600 $BENCH_SHORT > gpurun_out/plain_short.log 2>&1 &&
      timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv \
        --log-file gpurun_out/launches.csv $BENCH_SHORT > gpurun_out/ncu_list.log 2>&1
      echo "ncu_list rc=$?" | tee -a gpurun_out/summary.txt ;;
    ncu_full)
      timeout 600 $BENCH_SHORT > gpurun_out/plain_short2.log 2>&1 &&
      timeout 900 ncu --set full --clock-control none --import-source on -k regex:scan_topk -s 20 -c 3 \
        -f -o gpurun_out/scan_prof $BENCH_SHORT > gpurun_out/ncu_full.log 2>&1
      echo "ncu_full rc=$?" | tee -a gpurun_out/summary.txt ;;
  esac
done
cat gpurun_out/summary.txt
