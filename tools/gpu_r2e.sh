#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_batch_large.py tests/test_gpu_where.py -x -q -m gpu > gpurun_out/pytest_batch.log 2>&1; echo "pytest rc=$?"
tail -3 gpurun_out/pytest_batch.log
: > gpurun_out/variants3.txt
for env in "PVDB_X=0" "PVDB_BATCH_TILE_BLOCK=1" "PVDB_BATCH_TILE_BLOCK=2" "PVDB_BATCH_TILE_BLOCK=4" "PVDB_BATCH_TILE_BLOCK=8"; do
  echo "== $env" >> gpurun_out/variants3.txt
  env $env timeout 300 python tools/bench_configs.py --custom 6000000,384,4096,10,bf16 --custom 3000000,128,4096,10,bf16 --custom 2500000,768,4096,100,tf32 none 2>/dev/null | grep config | cut -c1-150 >> gpurun_out/variants3.txt
done
cat gpurun_out/variants3.txt
