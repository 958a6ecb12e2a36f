"""world_size-2 test of the row-sharded search flow on CPU (gloo): partition arithmetic, result
packing, the all-gather and the merge call.  Local scans and the merge use the test engine."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import picovdb_oracle as O
from picovdb_b200.sharded import ShardedSearch, owner_of, shard_range

from _host_engine import HostEngine


def _free_port() -> int:
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank: int, world: int, port: int, out_dir: str) -> None:
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        n, dim, k = 1000, 24, 10
        rng = np.random.default_rng(42)
        full = O.normalize_rows_fast(rng.standard_normal((n, dim)).astype(np.float32))
        dead = rng.choice(n, 150, replace=False)
        queries = rng.standard_normal((5, dim)).astype(np.float32)
        r0, r1 = shard_range(n, world, rank)
        eng = HostEngine(dim)
        eng.upload(full[r0:r1], 0, None)
        local_dead = dead[(dead >= r0) & (dead < r1)] - r0
        eng.delete_rows(local_dead)
        sh = ShardedSearch(eng, r0, merge=O.merge_topk)
        pf_global = (np.arange(n) % 3) != 0
        s, r = sh.search(queries, k)
        s2, r2 = sh.search(queries, k, prefilter=pf_global[r0:r1])
        np.savez(os.path.join(out_dir, f"rank{rank}.npz"), s=s, r=r, s2=s2, r2=r2)
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(120)
def test_two_rank_sharded_search_matches_unsharded(tmp_path):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    n, dim, k = 1000, 24, 10
    rng = np.random.default_rng(42)
    full = O.normalize_rows_fast(rng.standard_normal((n, dim)).astype(np.float32))
    dead = rng.choice(n, 150, replace=False)
    queries = rng.standard_normal((5, dim)).astype(np.float32)
    active = np.ones(n, bool)
    active[dead] = False
    full[dead] = 0
    qn, _ = O.prepare_queries(queries, dim)
    ref_s, ref_r = O.search(full, qn, k, active)
    ref_s2, ref_r2 = O.search(full, qn, k, active, (np.arange(n) % 3) != 0)
    outs = [np.load(str(tmp_path / f"rank{r}.npz")) for r in range(world)]
    for o in outs:  # every rank ends with the same, correct answer
        np.testing.assert_array_equal(o["r"], ref_r)
        np.testing.assert_allclose(o["s"], ref_s, rtol=2e-6)
        np.testing.assert_array_equal(o["r2"], ref_r2)
        np.testing.assert_allclose(o["s2"], ref_s2, rtol=2e-6)


def test_shard_ranges_cover_and_align():
    for total in (0, 1, 31, 32, 33, 1000, 1_000_000, 100_000_000):
        for world in (1, 2, 3, 4, 8):
            edges = [shard_range(total, world, r) for r in range(world)]
            assert edges[0][0] == 0 and edges[-1][1] == total
            for (a0, a1), (b0, b1) in zip(edges, edges[1:]):
                assert a1 == b0 and a0 <= a1
            assert all(e[0] % 32 == 0 for e in edges if e[1] > e[0])  # non-empty shards
            for row in (0, total // 2, max(0, total - 1)):
                if total:
                    r = owner_of(row, total, world)
                    assert edges[r][0] <= row < edges[r][1]
