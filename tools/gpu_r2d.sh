#!/usr/bin/env bash
# variant timings of the batch kernel after the uniform-issue change
set -u
mkdir -p gpurun_out
: > gpurun_out/variants2.txt
for env in "PVDB_BATCH_ALT=0" "PVDB_BATCH_PAIR=1" "PVDB_BATCH_ALT=1" "PVDB_BATCH_NO_CLUSTER=1" "PVDB_BATCH_CLUSTER=4" "PVDB_BATCH_PAIR=1 PVDB_BATCH_ALT=1"; do
  echo "== $env" >> gpurun_out/variants2.txt
  env $env timeout 300 python tools/bench_configs.py --custom 6000000,384,4096,10,bf16 --custom 3000000,128,4096,10,bf16 none 2>/dev/null | grep config >> gpurun_out/variants2.txt
done
cat gpurun_out/variants2.txt
