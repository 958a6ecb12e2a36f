#!/usr/bin/env bash
# Round-2 (second session) experiment run: warp-uniform MMA/TMA issue.  gpurun --timeout 900 -- 'bash tools/gpu_r2b.sh'
set -u
mkdir -p gpurun_out
timeout 200 ./tools/micro/tmem_ld_mma_bench > gpurun_out/tmem_ld_mma_bench.jsonl 2>&1; echo "micro rc=$?"
timeout 200 ./tools/micro/tmem_ld_bench > gpurun_out/tmem_ld_bench2.jsonl 2>&1; echo "micro0 rc=$?"
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_batch_large.py -x -q -m gpu > gpurun_out/pytest_batch.log 2>&1; echo "pytest rc=$?"
tail -3 gpurun_out/pytest_batch.log
timeout 300 python tools/bench_configs.py --custom 6000000,384,4096,10,bf16 --custom 3000000,128,4096,10,bf16 --custom 2500000,768,4096,100,tf32 none > gpurun_out/cfg_uniform.jsonl 2>gpurun_out/cfg_uniform.err; echo "cfg rc=$?"
cat gpurun_out/cfg_uniform.jsonl; tail -3 gpurun_out/cfg_uniform.err
