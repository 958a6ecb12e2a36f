#!/usr/bin/env python
"""Scan-kernel time vs shard size (fixed cost vs streaming part).  python tools/bench_scan_sizes.py"""
import json, os, sys, time
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from picovdb_b200.engine import DeviceStore

dev = torch.device("cuda", 0)
dim, k = 1024, 10
stream = torch.cuda.current_stream().cuda_stream
for rows in [int(x) for x in os.environ.get("PVDB_SIZES", "1024,16384,62500,125000,250000,500000,1000000").split(",")]:
    st = DeviceStore(dim, device=0, reserve_rows=rows)
    gen = torch.Generator(device=dev).manual_seed(1)
    for r0 in range(0, rows, 131072):
        m = min(131072, rows - r0)
        x = torch.randn(m, dim, device=dev, generator=gen)
        st.upsert_range_dev(x.data_ptr(), r0, m, stream=stream)
        torch.cuda.synchronize()
    q = torch.nn.functional.normalize(torch.randn(64, dim, device=dev), dim=1).contiguous()
    out_s = torch.empty(k, dtype=torch.float32, device=dev); out_r = torch.empty(k, dtype=torch.int64, device=dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    for j in range(10):
        st.search_dev(q[j].data_ptr(), 1, k, out_s.data_ptr(), out_r.data_ptr(), precision="f32", normalized=True, stream=stream)
    torch.cuda.synchronize()
    n = 200
    flush.zero_()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    t0 = time.perf_counter()
    for j in range(n):
        st.search_dev(q[j % 64].data_ptr(), 1, k, out_s.data_ptr(), out_r.data_ptr(), precision="f32", normalized=True, stream=stream)
    host_us = (time.perf_counter() - t0) / n * 1e6
    e1.record(); torch.cuda.synchronize()
    us = e0.elapsed_time(e1) / n * 1e3
    print(json.dumps({"rows": rows, "us_per_scan": round(us, 2), "host_enqueue_us": round(host_us, 2), "stream_us_at_7TBs": round(rows * dim * 4 / 7.0e6, 2)}), flush=True)
    st.close()
