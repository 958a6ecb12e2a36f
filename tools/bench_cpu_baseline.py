#!/usr/bin/env python
"""CPU baseline beside the GPU numbers: the oracle port of the reference's NumPy query
path, timed on this machine's host cores for BASELINE config C1 (100k x 1024 fp32: one 1000-query batch
and 100 single queries, mirroring the reference's bench/batch_queries.py and bench/queries.py).
The reference itself cannot travel to the GPU box; the oracle restates its arithmetic line by line
(oracle/picovdb_oracle.py) and is pinned to the reference's outputs by tests/test_oracle_golden.py."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import picovdb_oracle as O  # noqa: E402


def main():
    n, dim, nq, k = 100_000, 1024, 1000, 10
    rng = np.random.default_rng(123)
    store = O.normalize_rows_fast(rng.standard_normal((n, dim)).astype(np.float32))
    q = np.random.default_rng(99).standard_normal((nq, dim)).astype(np.float32)
    threads = None
    try:
        from threadpoolctl import threadpool_info

        threads = [(i.get("internal_api"), i.get("num_threads")) for i in threadpool_info()]
    except Exception:
        pass
    qn, _ = O.prepare_queries(q, dim)
    O.search(store, qn[:10], k)
    t0 = time.perf_counter()
    O.search(store, qn, k)
    t_batch = time.perf_counter() - t0
    t0 = time.perf_counter()
    for i in range(100):
        O.search(store, qn[i: i + 1], k)
    t_single = time.perf_counter() - t0
    print(json.dumps({
        "config": "C1 100k x 1024 fp32, top-10, oracle port of the reference's NumPy path on the host",
        "cpu_count": os.cpu_count(), "blas_threads": threads,
        "batch_1000_queries_ms": t_batch * 1e3, "batch_qps": nq / t_batch,
        "single_query_ms": t_single / 100 * 1e3, "single_qps": 100 / t_single,
    }))


if __name__ == "__main__":
    main()
