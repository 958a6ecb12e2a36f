#!/usr/bin/env python
"""Where the batch kernel's warps spend their cycles (needs a library built with
PVDB_NVCC_EXTRA=-DPVDB_BATCH_STATS).  CAUTION: every counter is two clock64() reads and one read costs ~100
cycles on the B200; with ~14 reads per visit the instrumented kernel is 25-40 % slower and small intervals are
mostly the reads themselves (profiles/round2/README.md, second half, 1.).  Use ncu's source-level sampling
(--set full --import-source on; tools/gpu_round2b.sh ncu384) for anything finer than per-visit totals.  Usage: python tools/batch_stats.py rows,dim,nq,k,prec[,mirror] ..."""
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
from tools.bench_configs import fill, time_search  # noqa: E402

NAMES = ["epi_loop_cyc", "epi_wait_tfull_cyc", "epi_prune_cyc", "prunes", "appended", "epi_final_prune_cyc",
         "mma_wait_tempty_cyc", "mma_wait_full_cyc", "mma_total_cyc", "epi_warp_visits", "launches",
         "epi_wait_tmem_ld_cyc", "epi_scan_chunk_cyc", "tma_wait_empty_cyc", "tma_total_cyc"]


def main():
    lib = N.load()
    fn = lib.pvdb_debug_batch_stats
    fn.argtypes = [C.POINTER(C.c_ulonglong), C.c_int, C.c_int]
    dev = torch.device("cuda", 0)
    for spec in sys.argv[1:]:
        parts = spec.split(",")
        rows, dim, nq, k = (int(x) for x in parts[:4])
        prec = parts[4]
        mirror = prec == "bf16" or (len(parts) > 5 and parts[5] == "mirror")
        st = DeviceStore(dim, device=0, reserve_rows=rows, bf16_mirror=mirror)
        fill(st, rows, dim, 123, dev)
        q = torch.randn(nq, dim, device=dev, generator=torch.Generator(device=dev).manual_seed(99))
        ms, _, _ = time_search(st, q, k, prec, iters=3)
        buf = (C.c_ulonglong * 16)()
        fn(buf, 16, 1)
        ms1, _, _ = time_search(st, q, k, prec, iters=1, warm=0)
        fn(buf, 16, 0)
        v = dict(zip(NAMES, [int(x) for x in buf[:15]]))
        launches = max(v["launches"], 1)
        ew = 148 * 8  # epilogue warps per launch (upper bound: idle units count as zero time)
        out = {"config": spec, "ms": ms, "launches_in_sample": launches}
        for name in ("epi_loop_cyc", "epi_wait_tfull_cyc", "epi_prune_cyc", "epi_final_prune_cyc",
                     "epi_wait_tmem_ld_cyc", "epi_scan_chunk_cyc"):
            out[name + "_per_warp"] = v[name] / ew
        out["prunes_per_warp"] = v["prunes"] / ew
        out["appended_per_query_state"] = v["appended"] / (148 * 128)
        out["visits_per_warp"] = v["epi_warp_visits"] / ew
        for name in ("mma_wait_tempty_cyc", "mma_wait_full_cyc", "mma_total_cyc", "tma_wait_empty_cyc", "tma_total_cyc"):
            out[name + "_per_cta"] = v[name] / 148
        if v["prunes"]:
            out["cyc_per_prune"] = v["epi_prune_cyc"] / v["prunes"]
        print(json.dumps(out), flush=True)
        st.close()


if __name__ == "__main__":
    main()
