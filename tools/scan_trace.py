#!/usr/bin/env python
"""Where a scan launch spends its fixed cost: %globaltimer stamps from a -DPVDB_SCAN_TRACE variant build.

    python -m picovdb_b200.build --variant trace -DPVDB_SCAN_TRACE
    PICOVDB_B200_LIB=picovdb_b200/_variants/libpicovdb_b200_trace.so python tools/scan_trace.py

Block 0 stamps kernel entry / after the prologue / after its walk / after its block merge + ticket; the last
block stamps its start / after folding the per-block lists / before the result write.  Medians over 30 launches.
"""
import ctypes as C
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from picovdb_b200 import _native as N  # noqa: E402
from picovdb_b200.engine import DeviceStore  # noqa: E402

lib = N.load()
lib.pvdb_debug_scan_trace.argtypes = [C.c_void_p, C.c_void_p]
lib.pvdb_debug_scan_trace.restype = C.c_int
dev = torch.device("cuda", 0)
stream = torch.cuda.current_stream().cuda_stream
dim, k = 1024, 10
for rows in (1024, 125_000, 1_000_000):
    st = DeviceStore(dim, device=0, reserve_rows=rows)
    gen = torch.Generator(device=dev).manual_seed(1)
    for r0 in range(0, rows, 131072):
        m = min(131072, rows - r0)
        st.upsert_range_dev(torch.randn(m, dim, device=dev, generator=gen).data_ptr(), r0, m, stream=stream)
        torch.cuda.synchronize()
    q = torch.nn.functional.normalize(torch.randn(64, dim, device=dev), dim=1).contiguous()
    out_s = torch.empty(k, dtype=torch.float32, device=dev)
    out_r = torch.empty(k, dtype=torch.int64, device=dev)
    rec = []
    for j in range(40):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        st.search_dev(q[j % 64].data_ptr(), 1, k, out_s.data_ptr(), out_r.data_ptr(), precision="f32", normalized=True, stream=stream)
        e1.record()
        torch.cuda.synchronize()
        t = np.zeros(24, dtype=np.uint64)
        N.check(lib.pvdb_debug_scan_trace(st.handle, t.ctypes.data_as(C.c_void_p)))
        t = t.astype(np.int64)
        if j >= 10:
            rec.append([t[1] - t[0], t[2] - t[1], t[3] - t[2], t[4] - t[0], t[5] - t[4], t[6] - t[5], t[6] - t[0],
                        e0.elapsed_time(e1) * 1e6, t[7] - t[4], t[8] - t[7], t[9] - t[8], t[5] - t[9]])
    med = np.median(np.array(rec, dtype=np.float64), axis=0) / 1e3
    print(json.dumps({"rows": rows, "dim": dim, "k": k, "us": {
        "block0_prologue": round(med[0], 2), "block0_walk": round(med[1], 2), "block0_merge_ticket": round(med[2], 2),
        "kernel_entry_to_last_block_start": round(med[3], 2), "last_block_fold_block_lists": round(med[4], 2),
        "last_block_fold_parts(fence,first_heads_arrive,inserts+second_round,barrier)": [round(med[8], 2), round(med[9], 2), round(med[10], 2), round(med[11], 2)],
        "last_block_final_merge": round(med[5], 2), "block0_entry_to_result": round(med[6], 2),
        "cuda_events_around_the_call": round(med[7], 2)}}), flush=True)
    st.close()
