#!/usr/bin/env bash
# 8-GPU evidence: smoke on GPU 0, the driver's torchrun bench command at N = 8, 4, 2 (fused exchange), the reference arm at N = 8.
#   gpurun --gpus 8 --timeout 900 -- 'bash tools/gpu_r2_n8.sh'
set -u
mkdir -p gpurun_out
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/smoke.log
for n in ${NS:-8 4 2}; do
  timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2950$n \
    bench.py --gpus $n --steps 20 --warmup 5 > gpurun_out/bench_n$n.json 2> gpurun_out/bench_n$n.err; echo "bench$n rc=$?"
done
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 \
  bench.py --impl reference --gpus 8 --steps 3 --warmup 3 > gpurun_out/bench_ref_n8.json 2> gpurun_out/bench_ref_n8.err; echo "ref8 rc=$?"
