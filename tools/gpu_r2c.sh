#!/usr/bin/env bash
# ncu captures of the batch kernel's main pass (after the plain command exited 0).  gpurun --timeout 1200 -- 'bash tools/gpu_r2c.sh'
set -u
mkdir -p gpurun_out
for spec in 3000000,384,4096,10,bf16 3000000,128,4096,10,bf16; do
  tag=$(echo $spec | cut -d, -f2)
  CMD="python tools/bench_configs.py --custom $spec none"
  timeout 300 $CMD > gpurun_out/plain_batch_$tag.log 2>&1 &&
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:batch_topk -s 8 -c 1 \
    -f -o gpurun_out/batch_bf16_main_d$tag $CMD > gpurun_out/ncu_batch_$tag.log 2>&1
  echo "ncu $tag rc=$?"
  cat gpurun_out/plain_batch_$tag.log | head -2
done
