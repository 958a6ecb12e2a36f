#!/usr/bin/env python
"""Exact batches (no tensor cores): several queries per pass over the matrix vs one pass per query.

    python tools/bench_exact_batch.py            # PVDB_SCAN_NO_MULTI=1 for the one-pass-per-query form

Times `search_dev(..., scan_only=True)` (device-resident queries and results, CUDA events, 3 warm-ups, stores
larger than L2 or flushed) on fp32 stores (4 queries per pass) and a bf16-only store (2 per pass on mma.sync),
and checks the batch against single-query calls bit for bit on 8 queries.  One JSON line per case.
"""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from picovdb_b200.engine import DeviceStore  # noqa: E402

dev = torch.device("cuda", 0)
stream = torch.cuda.current_stream().cuda_stream
l2buf = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
CASES = [  # rows, dim, queries, k, bf16-only
    (100_000, 1024, 1000, 10, False),
    (1_000_000, 1024, 64, 10, False),
    (5_000_000, 384, 64, 10, False),
    (12_500_000, 384, 64, 10, True),
    (3_000_000, 128, 64, 10, True),
    (1_000_000, 1024, 64, 10, True),
]
if os.environ.get("PVDB_CASES"):   # e.g. PVDB_CASES=1,3 (for ncu captures)
    CASES = [CASES[int(i)] for i in os.environ["PVDB_CASES"].split(",")]
for rows, dim, nq, k, b16 in CASES:
    st = DeviceStore(dim, device=0, reserve_rows=rows, **({"keep_f32": False, "bf16_mirror": True} if b16 else {}))
    gen = torch.Generator(device=dev).manual_seed(123)
    chunk = max(1, (256 << 20) // (dim * 4))
    for r0 in range(0, rows, chunk):
        m = min(chunk, rows - r0)
        st.upsert_range_dev(torch.randn(m, dim, device=dev, generator=gen).data_ptr(), r0, m, stream=stream)
        torch.cuda.synchronize()
    q = torch.randn(nq, dim, device=dev, generator=gen)
    out_s = torch.empty((nq, k), dtype=torch.float32, device=dev)
    out_r = torch.empty((nq, k), dtype=torch.int64, device=dev)
    prec = "bf16" if b16 else "f32"

    def run(n=nq, qq=q, s=out_s, r=out_r):
        st.search_dev(qq.data_ptr(), n, k, s.data_ptr(), r.data_ptr(), precision=prec, scan_only=True, stream=stream)

    for _ in range(3):
        run()
    torch.cuda.synchronize()
    times = []
    for _ in range(5):
        l2buf.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        run()
        e1.record()
        torch.cuda.synchronize()
        times.append(e0.elapsed_time(e1))
    ms = float(np.median(times))
    # bit-for-bit against lone queries
    one_s = torch.empty((1, k), dtype=torch.float32, device=dev)
    one_r = torch.empty((1, k), dtype=torch.int64, device=dev)
    same = True
    for i in range(min(8, nq)):
        run(1, q[i:i + 1].contiguous(), one_s, one_r)
        torch.cuda.synchronize()
        same &= bool((one_s[0] == out_s[i]).all() and (one_r[0] == out_r[i]).all())
    nbytes = rows * dim * (2 if b16 else 4)
    print(json.dumps({"rows": rows, "dim": dim, "store": prec, "queries": nq, "k": k, "ms": round(ms, 3),
                      "us_per_query": round(ms / nq * 1e3, 2), "qps": round(nq / ms * 1e3, 1),
                      "matrix_gbs_per_query": round(nbytes / (ms / nq) / 1e6, 1),
                      "one_pass_per_query": os.environ.get("PVDB_SCAN_NO_MULTI") is not None,
                      "equals_single_queries_bitwise": same}), flush=True)
    st.close()
