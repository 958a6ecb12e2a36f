#!/usr/bin/env bash
# quick smoke (bounded) -> batch parity tests -> variant timings.  Every step has its own short timeout.
set -u
mkdir -p gpurun_out
timeout 120 python tools/bench_configs.py --custom 300000,128,512,10,bf16 none > gpurun_out/smoke_batch.log 2>&1 || { echo "SMOKE FAILED rc=$?"; tail -5 gpurun_out/smoke_batch.log; exit 1; }
tail -2 gpurun_out/smoke_batch.log | cut -c1-160
timeout 400 python -m pytest tests/test_gpu_parity.py tests/test_gpu_batch_large.py tests/test_gpu_where.py -x -q -m gpu --timeout 120 > gpurun_out/pytest_batch.log 2>&1; echo "pytest rc=$?"
tail -3 gpurun_out/pytest_batch.log
: > gpurun_out/variants3.txt
for env in ${VARIANTS:-"PVDB_X=0"}; do
  echo "== $env" >> gpurun_out/variants3.txt
  env $env timeout 200 python tools/bench_configs.py --custom 6000000,384,4096,10,bf16 --custom 3000000,128,4096,10,bf16 --custom 2500000,768,4096,100,tf32 c1 none 2>/dev/null | grep config | cut -c1-150 >> gpurun_out/variants3.txt
done
cat gpurun_out/variants3.txt
