#!/usr/bin/env python
"""Write-side measurements (SURVEY.md 8a rows a2/a3 and the persistence rows): how fast vectors get
into and out of the device store, next to the reference's published insert time for the same shape
(100,000 x 1024: 0.5 - 0.7 s on an M3 / i7, 3.4 s measured in the survey container).

    python tools/bench_ingest.py            # needs a B200
"""
from __future__ import annotations

import json
import os
import sys
import tempfile
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from picovdb_b200 import K_ID, K_VECTOR, PicoVectorDB  # noqa: E402
from picovdb_b200.engine import DeviceStore  # noqa: E402


def emit(**kw):
    print(json.dumps(kw), flush=True)


def main():
    n, dim = 100_000, 1024
    rng = np.random.default_rng(0)
    vecs = rng.random((n, dim), dtype=np.float32)  # the reference's bench/upserts.py distribution
    with tempfile.TemporaryDirectory() as tmp:
        db = PicoVectorDB(embedding_dim=dim, storage_file=os.path.join(tmp, "ingest"))
        items = [{K_VECTOR: vecs[i], K_ID: i} for i in range(n)]
        t0 = time.perf_counter()
        db.upsert(items)
        t_dict = time.perf_counter() - t0
        emit(case="PicoVectorDB.upsert(list of dicts), 100k x 1024", seconds=t_dict, vec_per_s=n / t_dict,
             reference="0.5-0.7 s published (README.md:69,81), 3.36 s in the survey container")
        t0 = time.perf_counter()
        db.save()
        t_save = time.perf_counter() - t0
        db.close()
        t0 = time.perf_counter()
        db2 = PicoVectorDB(embedding_dim=dim, storage_file=os.path.join(tmp, "ingest"))
        t_load = time.perf_counter() - t0
        emit(case="save() / reload of 100k x 1024 (410 MB .npy + json)", save_s=t_save, load_s=t_load)
        t0 = time.perf_counter()
        res = db2.query(vecs[:100], top_k=10)
        t_q = time.perf_counter() - t0
        assert [r[0][K_ID] for r in res] == list(range(100))
        emit(case="query(100-query batch) incl. dict assembly", seconds=t_q, qps=100 / t_q)
        t0 = time.perf_counter()
        for i in range(100):
            db2.query(vecs[i], top_k=10, better_than=0.1)
        t_s = time.perf_counter() - t0
        emit(case="100 single query() calls incl. dict assembly (reference: 0.8-1.5 s published)", seconds=t_s,
             qps=100 / t_s)
        db2.close()

        db3 = PicoVectorDB(embedding_dim=dim, storage_file=os.path.join(tmp, "bulk"))
        t0 = time.perf_counter()
        db3.upsert_array(vecs)
        t_bulk = time.perf_counter() - t0
        emit(case="PicoVectorDB.upsert_array(100k x 1024) (pageable host memory)", seconds=t_bulk,
             vec_per_s=n / t_bulk, gb_per_s=vecs.nbytes / t_bulk / 1e9)
        db3.close()

    # larger stores: bulk ingest, streamed save and load -- GB/s on the
    # vector bytes, files on the box's local disk (tmpdir)
    n2, dim2 = 4_000_000, 384
    big = np.random.default_rng(2).standard_normal((n2, dim2), dtype=np.float32)
    with tempfile.TemporaryDirectory() as tmp:
        for label, kw, bytes_per in (("fp32 store", {}, 4), ("bf16-only store", {"keep_f32": False, "bf16_mirror": True}, 2)):
            path = os.path.join(tmp, "big" + str(bytes_per))
            db = PicoVectorDB(embedding_dim=dim2, storage_file=path, **kw)
            db.upsert_array(big[:1000])     # warm the pinned buffers / kernels
            t0 = time.perf_counter()
            db.upsert_array(big[1000:])
            t_in = time.perf_counter() - t0
            t0 = time.perf_counter()
            db.save()
            t_save = time.perf_counter() - t0
            db.close()
            t0 = time.perf_counter()
            db = PicoVectorDB(embedding_dim=dim2, storage_file=path, **kw)
            t_load = time.perf_counter() - t0
            assert len(db) == n2
            file_bytes = n2 * dim2 * bytes_per
            emit(case=f"{label}: upsert_array / save / load of {n2} x {dim2}",
                 ingest_s=t_in, ingest_gb_per_s=big[1000:].nbytes / t_in / 1e9,
                 save_s=t_save, save_gb_per_s=file_bytes / t_save / 1e9,
                 load_s=t_load, load_gb_per_s=file_bytes / t_load / 1e9, file_gb=file_bytes / 1e9)
            db.close()
    del big

    # device-resident source: the fused normalise + scatter kernel alone
    dev = torch.device("cuda", 0)
    store = DeviceStore(dim, device=0, reserve_rows=1_000_000, bf16_mirror=True)
    x = torch.randn(250_000, dim, device=dev)
    stream = torch.cuda.current_stream().cuda_stream
    for r0 in (0, 250_000):
        store.upsert_range_dev(x.data_ptr(), r0, 250_000, stream=stream)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for r0 in (0, 250_000, 500_000, 750_000):
        store.upsert_range_dev(x.data_ptr(), r0, 250_000, stream=stream)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    moved = 1_000_000 * dim * (4 + 4 + 4 + 2)  # two reads of the source row, fp32 write, bf16 write
    emit(case="upsert_normalize_scatter_kernel, 1M x 1024 from HBM (fp32 + bf16 mirror)", ms=ms,
         vec_per_s=1e6 / ms * 1e3, hbm_gbs=moved / ms / 1e6)
    # delete 30 % + vacuum-style compaction
    dead = np.random.default_rng(1).choice(1_000_000, 300_000, replace=False)
    t0 = time.perf_counter()
    store.delete_rows(dead)
    t_del = time.perf_counter() - t0
    keep = np.setdiff1d(np.arange(1_000_000), dead)
    t0 = time.perf_counter()
    store.compact(keep)
    t_cmp = time.perf_counter() - t0
    emit(case="delete 300k rows / compact 700k rows of 1M x 1024", delete_s=t_del, compact_s=t_cmp)
    store.close()


if __name__ == "__main__":
    main()
