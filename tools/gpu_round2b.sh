#!/usr/bin/env bash
# Second-half-of-round-2 experiment runs on ONE B200:  gpurun --timeout 900 -- 'bash tools/gpu_round2b.sh <stage>...'
# Every step carries its own short timeout (a kernel that deadlocks must not eat the box's time limit).
#   micro      the three micro-benchmarks under tools/micro/ (build them first: see the header of each .cu)
#   smoke      one bounded batch search (run before anything else after a kernel change)
#   tests      the batch / scan / where parity tests
#   variants   timings of $VARIANTS (space-separated VAR=value settings, e.g. "PVDB_X=0 PVDB_BATCH_PAIR=1";
#              PICOVDB_B200_LIB=picovdb_b200/_variants/libpicovdb_b200_<name>.so selects a variant build)
#   ncu384 / ncu128   ncu --set full of one main-pass launch of the batch kernel (3M rows, 4096 queries)
#   configs    tools/bench_configs.py c1 c3 c4 c5s c5b
set -u
mkdir -p gpurun_out
for st in "$@"; do
  case $st in
    micro)
      for b in tmem_ld_bench tmem_ld_mma_bench hbm_read_sustained; do
        timeout 300 ./tools/micro/$b > gpurun_out/$b.jsonl 2>&1; echo "$b rc=$?"
      done ;;
    smoke)
      timeout 120 python tools/bench_configs.py --custom 300000,128,512,10,bf16 none > gpurun_out/smoke_batch.log 2>&1 ||
        { echo "SMOKE FAILED rc=$?"; tail -5 gpurun_out/smoke_batch.log; exit 1; }
      tail -2 gpurun_out/smoke_batch.log | cut -c1-160 ;;
    tests)
      timeout 500 python -m pytest tests/test_gpu_parity.py tests/test_gpu_batch_large.py tests/test_gpu_where.py -x -q -m gpu \
        --timeout 120 > gpurun_out/pytest_batch.log 2>&1; echo "pytest rc=$?"
      tail -3 gpurun_out/pytest_batch.log ;;
    variants)
      : > gpurun_out/variants.txt
      for env in ${VARIANTS:-"PVDB_X=0"}; do
        echo "== $env" >> gpurun_out/variants.txt
        env $env timeout 200 python tools/bench_configs.py --custom 6000000,384,4096,10,bf16 --custom 3000000,128,4096,10,bf16 \
          --custom 2500000,768,4096,100,tf32 c1 none 2>/dev/null | grep config | cut -c1-150 >> gpurun_out/variants.txt
      done
      cat gpurun_out/variants.txt ;;
    ncu384|ncu128)
      tag=${st#ncu}
      CMD="python tools/bench_configs.py --custom 3000000,$tag,4096,10,bf16 none"
      # launches of batch_topk per search: seed, sample, main; warm-up 3 + timed: the 9th is a main pass
      timeout 200 $CMD > gpurun_out/plain_batch_$tag.log 2>&1 &&
      timeout 600 ncu --set full --clock-control none --import-source on -k regex:batch_topk -s 8 -c 1 \
        -f -o gpurun_out/batch_bf16_main_d$tag $CMD > gpurun_out/ncu_batch_$tag.log 2>&1
      echo "ncu $tag rc=$?"; head -1 gpurun_out/plain_batch_$tag.log | cut -c1-200 ;;
    configs)
      timeout 800 python tools/bench_configs.py c1 c3 c4 c5s c5b > gpurun_out/configs_final.jsonl 2> gpurun_out/configs_final.err
      echo "configs rc=$?"; cut -c1-200 gpurun_out/configs_final.jsonl ;;
  esac
done
