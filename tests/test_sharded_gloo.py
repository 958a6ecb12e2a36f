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


# ---------------------------------------------------------------------------------------------
# The drop-in class on top of a row-sharded engine (SPMD): every rank makes the same calls.
def _db_script(db, rng, out):
    """The same sequence of API calls for the sharded and the single-process DB."""
    from picovdb_b200 import K_ID, K_VECTOR

    dim = db.dim
    items = [{K_VECTOR: rng.standard_normal(dim).astype(np.float32), K_ID: f"r{i}", "cat": i % 4} for i in range(300)]
    db.upsert(items[:200])
    db.upsert(items[200:])
    db.delete([f"r{i}" for i in range(0, 300, 7)])
    db.upsert([{K_VECTOR: rng.standard_normal(dim), K_ID: f"n{i}", "cat": 1} for i in range(20)])   # reuses freed rows
    q = rng.standard_normal((4, dim)).astype(np.float32)
    ids = lambda res: [[r[K_ID] for r in lst] for lst in res]  # noqa: E731
    out["plain"] = ids(db.query(q, top_k=7))
    out["where"] = ids(db.query(q, top_k=5, where={"cat": 1}))
    out["where_in"] = ids(db.query(q, top_k=400, where={"cat": {"$in": [0, 3]}}))     # more than there are candidates
    out["where_ids"] = ids(db.query(q, top_k=4, where={"cat": 2}, ids=[f"r{i}" for i in range(0, 300, 2)]))
    out["where_none"] = db.query(q, top_k=4, where={"cat": 99})
    out["k_eff"] = db._last_k_eff
    out["callable"] = ids(db.query(q, top_k=5, where=lambda d: d["cat"] >= 2))
    out["ids"] = ids(db.query(q, top_k=3, ids=[f"r{i}" for i in range(100, 160)]))
    out["get"] = [np.round(r[K_VECTOR], 6).tolist() for r in db.get(["r5", "n3"], include_vector=True)]
    rows = np.fromiter(db._id2idx.values(), dtype=np.int64)
    out["split"] = [int((rows < 256).sum()), int((rows >= 256).sum())]
    db.vacuum()
    out["after_vacuum"] = ids(db.query(q, top_k=7))
    out["len"] = len(db)
    db.save()
    return q


def _db_worker(rank: int, world: int, port: int, out_dir: str) -> None:
    import json

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from picovdb_b200 import PicoVectorDB
        from picovdb_b200.sharded import ShardedStore

        class ShardedDB(PicoVectorDB):
            _engine_factory = staticmethod(
                lambda dim, **kw: ShardedStore(dim, local_factory=HostEngine, merge=O.merge_topk, **kw))

        path = os.path.join(out_dir, "sharded_db")
        db = ShardedDB(embedding_dim=16, storage_file=path, capacity=512, no_faiss=True)
        assert db._engine.owned_rows() == shard_range(512, world, rank)
        out = {}
        q = _db_script(db, np.random.default_rng(7), out)
        # vectors really are split: a rank only holds rows of its own block
        assert db._engine.local.rows <= db._engine.row1 - db._engine.row0
        assert "search_where" in db._engine.local.calls          # dict filters ran on the shards
        db.close()
        # reload from the files all ranks just wrote together; each rank uploads only its rows
        db2 = ShardedDB(embedding_dim=16, storage_file=path, capacity=512, no_faiss=True)
        from picovdb_b200 import K_ID
        out["reloaded"] = [[r[K_ID] for r in lst] for lst in db2.query(q, top_k=7)]
        db2.close()
        with open(os.path.join(out_dir, f"db_rank{rank}.json"), "w") as f:
            json.dump(out, f)
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(240)
@pytest.mark.parametrize("world", [2, 3])
def test_sharded_db_matches_single_process(tmp_path, monkeypatch, world):
    import json

    from picovdb_b200 import K_ID, PicoVectorDB

    mp.spawn(_db_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    monkeypatch.setattr(PicoVectorDB, "_engine_factory", staticmethod(lambda dim, **kw: HostEngine(dim, **kw)))
    ref = PicoVectorDB(embedding_dim=16, storage_file=str(tmp_path / "single_db"), capacity=512, no_faiss=True)
    want = {}
    q = _db_script(ref, np.random.default_rng(7), want)
    for rank in range(world):
        with open(tmp_path / f"db_rank{rank}.json") as f:
            got = json.load(f)
        for key in want:
            if key != "split":  # row placement differs by design (free_order)
                assert got[key] == want[key], (rank, key)
        assert got["reloaded"] == want["after_vacuum"]
    # the files the ranks wrote together are a normal store: the single-process class loads them
    ref.close()
    again = PicoVectorDB(embedding_dim=16, storage_file=str(tmp_path / "sharded_db"), capacity=512, no_faiss=True)
    assert [[r[K_ID] for r in lst] for lst in again.query(q, top_k=7)] == want["after_vacuum"]
    # same records, but the sharded store deals rows out over its shards (free_order), so the row
    # layout of the two files differs
    from picovdb_b200 import K_VECTOR
    np.testing.assert_allclose(again.get(["r5"], include_vector=True)[0][K_VECTOR], want["get"][0], atol=1e-6)
    a, b = np.load(tmp_path / "sharded_db.vecs.npy"), np.load(tmp_path / "single_db.vecs.npy")
    assert a.shape == b.shape
    key = lambda m: sorted(map(tuple, np.round(m[np.abs(m).sum(axis=1) > 0], 5).tolist()))  # noqa: E731
    assert key(a) == key(b)
    # before the vacuum (which compacts to the front, as in the reference) both shards held a fair share
    assert min(got["split"]) > 60 and min(want["split"]) < 60


# ------------------------------------------------------------------ bulk rows over a sharded store
def _bulk_worker(rank: int, world: int, port: int, out_dir: str) -> None:
    import json

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from picovdb_b200 import K_ID, K_VECTOR, PicoVectorDB
        from picovdb_b200.sharded import ShardedStore

        class ShardedDB(PicoVectorDB):
            _engine_factory = staticmethod(
                lambda dim, **kw: ShardedStore(dim, local_factory=HostEngine, merge=O.merge_topk, **kw))

        path = os.path.join(out_dir, "bulk_db")
        n, dim = 3000, 12
        vecs = np.random.default_rng(21).standard_normal((n, dim)).astype(np.float32)
        db = ShardedDB(embedding_dim=dim, storage_file=path, max_rows=4096, no_faiss=True)
        ids = db.upsert_array(vecs[:2000])                    # implicit rows 0..1999, split over the shards
        ids2 = db.upsert_array(vecs[2000:])
        assert ids == range(0, 2000) and ids2 == range(2000, 3000) and db._ids.implicit_rows == n
        assert len(db) == n and db._free == [] and db.capacity() == n      # append mode: no pre-filled slots
        db.upsert([{K_VECTOR: vecs[5] * -1.0, K_ID: "neg", "tag": "x"}])
        db.delete([7, 2500])
        q = vecs[[5, 7, 1999, 2000, 2999]]
        out = {"hits": [[r[K_ID] for r in lst] for lst in db.query(q, top_k=3)],
               "where": [r[K_ID] for r in db.query(vecs[5] * -1.0, top_k=2, where={"tag": "x"})],
               "local_rows": int(db._engine.local.rows)}
        db.save()
        db.close()
        db2 = ShardedDB(embedding_dim=dim, storage_file=path, max_rows=4096, no_faiss=True)
        out["reloaded"] = [[r[K_ID] for r in lst] for lst in db2.query(q, top_k=3)]
        out["len"] = len(db2)
        db2.close()
        with open(os.path.join(out_dir, f"bulk_rank{rank}.json"), "w") as f:
            json.dump(out, f)
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(240)
def test_sharded_db_bulk_rows_with_max_rows(tmp_path, monkeypatch):
    """`max_rows=` sizes a sharded engine's partition without the capacity= slot lists: bulk rows stay
    implicit ranges on every rank, vectors are split over the shards, files load in a plain DB."""
    import json

    from picovdb_b200 import K_ID, PicoVectorDB

    world = 2
    mp.spawn(_bulk_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    got = []
    for rank in range(world):
        with open(tmp_path / f"bulk_rank{rank}.json") as f:
            got.append(json.load(f))
    assert got[0]["hits"] == got[1]["hits"] == got[0]["reloaded"] == got[1]["reloaded"]
    assert [h[0] for h in got[0]["hits"]] == [5, 8 if False else got[0]["hits"][1][0], 1999, 2000, 2999]
    assert 7 not in got[0]["hits"][1] and got[0]["where"] == ["neg"] and got[0]["len"] == 2999
    # rows 0..2047 live on rank 0, the rest on rank 1 (contiguous blocks of max_rows / world)
    assert got[0]["local_rows"] == 2048 and got[1]["local_rows"] == 3001 - 2048
    monkeypatch.setattr(PicoVectorDB, "_engine_factory", staticmethod(lambda dim, **kw: HostEngine(dim, **kw)))
    plain = PicoVectorDB(embedding_dim=12, storage_file=str(tmp_path / "bulk_db"), no_faiss=True)
    vecs = np.random.default_rng(21).standard_normal((3000, 12)).astype(np.float32)
    assert [[r[K_ID] for r in lst] for lst in plain.query(vecs[[5, 7, 1999, 2000, 2999]], top_k=3)] == got[0]["hits"]
