#!/usr/bin/env python
"""C5 batch: 100M x 384 bf16 rows sharded over the ranks, 4096-query batches, top-10.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 \
        --master-port 29571 tools/bench_c5_batch.py [--rows 100000000] [--queries 4096]

Every rank runs the tcgen05 batch kernel over its shard for the whole query batch, then ONE
all-gather of the packed per-rank (4096 x 10) results and the merge kernel.  Timed with CUDA events
on the launching stream, max over ranks; roofline = 2*Q*N*dim flops over the sum of the ranks'
measured bf16 peaks.
"""
from __future__ import annotations

import argparse
import json
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from picovdb_b200.engine import DeviceStore  # noqa: E402
from picovdb_b200.sharded import ShardedSearch, shard_range  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--rows", type=int, default=100_000_000)
    ap.add_argument("--dim", type=int, default=384)
    ap.add_argument("--queries", type=int, default=4096)
    ap.add_argument("--k", type=int, default=10)
    ap.add_argument("--iters", type=int, default=5)
    ap.add_argument("--precision", default="bf16", choices=["bf16", "tf32"],
                    help="bf16: bf16-only store (C5); tf32: fp32 store, TF32 pass + fp32 re-scoring (C3)")
    ap.add_argument("--label", default=None)
    args = ap.parse_args()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    lr = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(lr)
    dev = torch.device("cuda", lr)
    if world > 1:
        os.environ.setdefault("NCCL_DEBUG", "WARN")
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
        dist.init_process_group("nccl", device_id=dev)
    r0, r1 = shard_range(args.rows, world, rank)
    n_local = r1 - r0
    is_bf16 = args.precision == "bf16"
    store = DeviceStore(args.dim, device=lr, reserve_rows=n_local, keep_f32=not is_bf16, bf16_mirror=is_bf16)
    gen = torch.Generator(device=dev).manual_seed(123 + rank)
    stream = torch.cuda.current_stream().cuda_stream
    chunk = 262144
    for c0 in range(0, n_local, chunk):
        m = min(chunk, n_local - c0)
        x = torch.randn(m, args.dim, device=dev, generator=gen)
        store.upsert_range_dev(x.data_ptr(), c0, m, stream=stream)
        torch.cuda.synchronize()
    sh = ShardedSearch(store, r0)
    q = torch.randn(args.queries, args.dim, device=dev, generator=torch.Generator(device=dev).manual_seed(99))
    for _ in range(3):
        sh.search_dev(q, args.k, precision=args.precision)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    times = []
    for _ in range(args.iters):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        s, r = sh.search_dev(q, args.k, precision=args.precision)
        e1.record()
        torch.cuda.synchronize()
        times.append(e0.elapsed_time(e1))
    t = torch.tensor(sorted(times)[len(times) // 2], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t)
    if rank == 0:
        peak = 1615.6
        p = os.path.join(ROOT, "MEASURED_PEAKS.json")
        if os.path.exists(p):
            with open(p) as f:
                peak = float(json.load(f)["bf16_tflops"])
        flops = 2.0 * args.queries * args.rows * args.dim
        if not is_bf16:
            peak = peak / 2  # no measured TF32 peak in MEASURED_PEAKS.json: nominal ratio to bf16
        label = args.label or ("C5 batch" if is_bf16 else "C3 batch")
        print(json.dumps({
            "config": f"{label}: {args.rows} x {args.dim} {args.precision} over {world} GPU(s), {args.queries} queries, top-{args.k}",
            "n_gpus": world, "ms": ms, "qps": args.queries / ms * 1e3, "tflops_aggregate": flops / ms / 1e9,
            "frac_tensor_peak_aggregate": flops / ms / 1e9 / (peak * world), "peak_per_gpu": peak,
            "peak_kind": "measured bf16" if is_bf16 else "measured bf16 / 2 (tf32)",
            "rows_ok": bool((r >= 0).all().item() and (r < args.rows).all().item()),
        }), flush=True)
    store.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
