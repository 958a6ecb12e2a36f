#!/usr/bin/env bash
# ncu captures of the batch kernel's main pass (after the plain command exited 0).  gpurun --timeout 1200 -- 'bash tools/gpu_r2c.sh [dims]'
set -u
mkdir -p gpurun_out
for tag in ${*:-384 128}; do
  spec=3000000,$tag,4096,10,bf16
  CMD="python tools/bench_configs.py --custom $spec none"
  timeout 200 $CMD > gpurun_out/plain_batch_$tag.log 2>&1 &&
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:batch_topk -s 8 -c 1 \
    -f -o gpurun_out/batch_bf16_main_d$tag $CMD > gpurun_out/ncu_batch_$tag.log 2>&1
  echo "ncu $tag rc=$?"
  cat gpurun_out/plain_batch_$tag.log | head -2
done
