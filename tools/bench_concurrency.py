#!/usr/bin/env python
"""Single-query throughput of ONE store under T concurrent reader threads (host buffers in and out,
one `pvdb_search` per query): the store mutex is held only while a call enqueues its work, so the
threads' host-side latency overlaps with each other's scans.

    python tools/bench_concurrency.py [--rows 1000000 --dim 1024]
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from picovdb_b200.engine import DeviceStore  # noqa: E402
from tools.bench_configs import fill  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--rows", type=int, default=1_000_000)
    ap.add_argument("--dim", type=int, default=1024)
    ap.add_argument("--queries", type=int, default=1024)
    args = ap.parse_args()
    dev = torch.device("cuda", 0)
    st = DeviceStore(args.dim, device=0, reserve_rows=args.rows)
    fill(st, args.rows, args.dim, 123, dev)
    q = np.random.default_rng(99).standard_normal((args.queries, args.dim)).astype(np.float32)
    for _ in range(20):
        st.search(q[:1], 10)
    for threads in (1, 2, 4, 8):
        per = args.queries // threads

        def work(t):
            for i in range(t * per, (t + 1) * per):
                st.search(q[i:i + 1], 10)

        ths = [threading.Thread(target=work, args=(t,)) for t in range(threads)]
        t0 = time.perf_counter()
        for th in ths:
            th.start()
        for th in ths:
            th.join()
        dt = time.perf_counter() - t0
        print(json.dumps({"case": f"{args.rows} x {args.dim} fp32, single queries, {threads} reader thread(s)",
                          "qps": per * threads / dt, "us_per_query": dt / (per * threads) * 1e6}), flush=True)
    st.close()


if __name__ == "__main__":
    main()
