"""Drop-in acceptance tests for ``picovdb_b200.PicoVectorDB``.

Each test states the reference behaviour it mirrors (reference tests file:line).  Every test
runs twice: with the test-only host engine (``-m "not gpu"``: covers the host logic of db.py) and
with the real CUDA engine through the C ABI (``-m gpu``).
"""
import json
import logging
import os
import threading
import time
import warnings
from unittest import mock

import numpy as np
import pytest

from picovdb_b200 import K_ID, K_METRICS, K_VECTOR, PicoVectorDB
from picovdb_b200 import db as dbmod

from _host_engine import HostEngine

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")


@pytest.fixture(params=[pytest.param("host"), pytest.param("cuda", marks=pytest.mark.gpu)])
def make_db(request, tmp_path, monkeypatch):
    if request.param == "host":
        monkeypatch.setattr(PicoVectorDB, "_engine_factory", staticmethod(lambda dim, **kw: HostEngine(dim, **kw)))
    made = []

    def factory(dim=3, name="store", **kw):
        kw.setdefault("no_faiss", True)
        d = PicoVectorDB(embedding_dim=dim, storage_file=str(tmp_path / name), **kw)
        made.append(d)
        return d

    factory.kind = request.param
    yield factory
    for d in made:
        d.close()


def ids_of(rows):
    return [r[K_ID] for r in rows]


# ------------------------------------------------------------------ upsert / ids / slots
def test_upsert_report_and_auto_ids(make_db):
    # reference tests/test_pico_vdb.py:69-85: re-upserting identical vectors hits the same md5 ids
    db = make_db(dim=5)
    rng = np.random.default_rng(0)
    vecs = rng.random((6, 5)).astype(np.float32)
    r1 = db.upsert([{K_VECTOR: v, "content": i} for i, v in enumerate(vecs)])
    assert len(r1["insert"]) == 6 and r1["update"] == []
    r2 = db.upsert([{K_VECTOR: v, "content": i} for i, v in enumerate(vecs)])
    assert r2["insert"] == [] and r2["update"] == r1["insert"]
    assert all(i == dbmod._hash_vec(dbmod._normalize(v)) for i, v in zip(r1["insert"], vecs))
    assert len(db) == 6


def test_same_id_twice_in_one_call_keeps_last(make_db):
    db = make_db(dim=2)
    rep = db.upsert([
        {K_VECTOR: [1.0, 0.0], K_ID: "a", "v": 1},
        {K_VECTOR: [0.0, 1.0], K_ID: "a", "v": 2},
    ])
    assert rep == {"update": ["a"], "insert": ["a"]}
    rec = db.get("a", include_vector=True)
    assert rec["v"] == 2
    np.testing.assert_allclose(rec[K_VECTOR], [0.0, 1.0], atol=1e-7)


def test_delete_then_insert_reuses_slot(make_db):
    # reference tests/test_more.py:108-130
    db = make_db()
    db.upsert([{K_VECTOR: np.eye(3, dtype=np.float32)[i], K_ID: str(i)} for i in range(3)])
    assert db.delete(["1", "missing"]) == ["1"]
    assert len(db) == 2 and db._free == [1]
    db.upsert([{K_VECTOR: np.ones(3, np.float32), K_ID: "new"}])
    assert db._id2idx["new"] == 1 and db._free == []
    assert "new" in ids_of(db.get_all())
    # SURVEY.md Q1 (deliberate divergence): the row owner is returned, not a permuted id
    assert ids_of(db.query(np.ones(3, np.float32), top_k=1)) == ["new"]


def test_upsert_validation_messages(make_db):
    # reference tests/test_task3_input_validation.py
    db = make_db(dim=4)
    with pytest.raises(ValueError, match="upsert vector must be 1D with length 4"):
        db.upsert([{K_VECTOR: np.zeros((2, 4), np.float32)}])
    with pytest.raises(ValueError, match="upsert vector dim mismatch: expected 4, got 3"):
        db.upsert([{K_VECTOR: np.zeros(3, np.float32)}])
    with pytest.raises(ValueError, match="query vector dim mismatch: expected 4, got 3"):
        db.query(np.zeros(3, np.float32))
    with pytest.raises(ValueError, match="query vectors dim mismatch: expected last dim 4, got 3"):
        db.query(np.zeros((2, 3), np.float32))
    with pytest.raises(ValueError, match="query expects 1D or 2D array with last dim 4"):
        db.query(np.zeros((1, 2, 4), np.float32))


def test_upsert_in_blocks_matches_one_block(make_db, monkeypatch):
    """upsert() sends its vectors to the device in blocks through one reused staging array.  With the block
    shrunk to 256 rows a 1000-item call spans four blocks: ids repeated across blocks keep their LAST vector,
    updates of rows appended earlier in the same call work, a validation error half way commits what came
    before it (pico_vdb.py:428-449), and the result equals the oracle DB's."""
    from oracle import picovdb_oracle as O

    monkeypatch.setattr(dbmod, "_UPSERT_BLOCK_BYTES", 1)
    dim, n = 8, 1000
    rng = np.random.default_rng(5)
    vecs = rng.standard_normal((n + 300, dim)).astype(np.float32)
    items = [{K_VECTOR: vecs[i], K_ID: f"id{i % 900}", "n": i} for i in range(n)]      # ids 0..99 appear twice
    db = make_db(dim=dim)
    odb = O.OracleDB(dim)
    rep, orep = db.upsert(items), odb.upsert(items)
    assert rep == orep and len(db) == 900
    for probe in (vecs[950], vecs[50], vecs[500]):
        got, want = db.query(probe, top_k=3), odb.query(probe, top_k=3)
        assert ids_of(got) == ids_of(want)
        np.testing.assert_allclose([g[K_METRICS] for g in got], [w[K_METRICS] for w in want], rtol=1e-5, atol=2e-6)
    assert db.get("id50")["n"] == 950 and db.query(vecs[950], top_k=1)[0][K_ID] == "id50"
    # a bad item at position 280 (second block): the flushed first block and the 24 staged items are committed
    more = [{K_VECTOR: vecs[n + i], K_ID: f"new{i}"} for i in range(300)]
    more[280] = {K_VECTOR: np.zeros(3, np.float32), K_ID: "bad"}
    with pytest.raises(ValueError, match="dim mismatch"):
        db.upsert(more)
    assert len(db) == 900 + 280 and db.get("new279") is not None and db.get("new280") is None
    assert db.query(vecs[n + 279], top_k=1)[0][K_ID] == "new279"
    assert db.query(vecs[n + 123], top_k=1)[0][K_ID] == "new123"


def test_stored_vector_is_normalised(make_db):
    # reference tests/test_more.py:236-260 and tests/test_memmap_capacity.py:42-47
    db = make_db(dim=2)
    db.upsert([{K_VECTOR: [3.0, 4.0], K_ID: "v"}])
    np.testing.assert_allclose(db.get("v", include_vector=True)[K_VECTOR], [0.6, 0.8], rtol=1e-6)
    assert db._vectors.dtype == np.float32 and db._vectors.flags["C_CONTIGUOUS"]
    np.testing.assert_allclose(db._vectors[0], [0.6, 0.8], rtol=1e-6)


def test_dtype_and_layout_of_inputs(make_db):
    # reference tests/test_task17_float32_contiguity.py
    db = make_db(dim=10)
    db.upsert([{K_VECTOR: np.arange(20, dtype=np.float32)[1::2], K_ID: "a"}])
    db.upsert([{K_VECTOR: np.arange(10, dtype=np.float64), K_ID: "b"}])
    assert db._vectors.dtype == np.float32 and db._vectors.flags["C_CONTIGUOUS"]
    db.delete(["a"])
    assert np.all(db._vectors[0] == 0)  # deleted rows are zero-filled (pico_vdb.py:523)
    qs = np.asfortranarray(np.random.default_rng(0).random((3, 10)).astype(np.float32))
    res = db.query(qs, top_k=1)
    assert isinstance(res, list) and len(res) == 3 and ids_of(res[0]) == ["b"]


# ------------------------------------------------------------------ query results
def test_basis_queries_single_and_batch(make_db):
    # reference tests/test_more.py:133-155
    db = make_db()
    eye = np.eye(3, dtype=np.float32)
    db.upsert([{K_VECTOR: v, K_ID: str(i)} for i, v in enumerate(eye)])
    res = db.query(np.array([0.9, 0.1, 0.0], np.float32), top_k=2)
    assert ids_of(res) == ["0", "1"]
    assert res[0][K_METRICS] == pytest.approx(0.9 / np.sqrt(0.82), rel=1e-5)
    batch = db.query(np.stack([eye[2], eye[1]]), top_k=1)
    assert [ids_of(r) for r in batch] == [["2"], ["1"]]
    assert ids_of(db.query_one(eye[1], top_k=1)) == ["1"]


def test_zero_vectors(make_db):
    # reference tests/test_task5_zero_vector_normalization.py
    db = make_db()
    db.upsert([{K_VECTOR: np.zeros(3, np.float32), K_ID: "z"}])
    res = db.query(np.zeros(3, np.float32), top_k=1)
    assert res[0][K_ID] == "z" and res[0][K_METRICS] == pytest.approx(1.0, rel=1e-5)
    np.testing.assert_array_equal(db.get("z", include_vector=True)[K_VECTOR], [1.0, 0.0, 0.0])
    db2 = make_db(name="basis")
    db2.upsert([{K_VECTOR: v, K_ID: str(i)} for i, v in enumerate(np.eye(3, dtype=np.float32))])
    assert db2.query(np.zeros(3, np.float32), top_k=3)[0][K_ID] == "0"


def test_only_active_rows_are_returned(make_db):
    # reference tests/test_task2_numpy_query_active_indices.py
    db = make_db(dim=8)
    rng = np.random.default_rng(5)
    v = rng.random((30, 8)).astype(np.float32)
    db.upsert([{K_VECTOR: v[i], K_ID: f"id{i}"} for i in range(30)])
    db.delete([f"id{i}" for i in range(20)])
    q = rng.random(8).astype(np.float32)
    res = db.query(q, top_k=25)
    assert len(res) == 10 and set(ids_of(res)) == {f"id{i}" for i in range(20, 30)}
    sc = [r[K_METRICS] for r in res]
    assert sc == sorted(sc, reverse=True)
    db2 = make_db(dim=8, name="full")
    db2.upsert([{K_VECTOR: v[i], K_ID: f"id{i}"} for i in range(30)])
    assert len(db2.query(q, top_k=15)) == 15


def test_seeded_topk_small_and_large_k(make_db):
    # reference tests/test_task20_argsort_vs_argpartition.py (same seed / shapes / baseline)
    g = np.load(os.path.join(GOLDEN, "task20.npz"))
    db = make_db(dim=16)
    db.upsert([{K_VECTOR: g["raw"][i], K_ID: str(i)} for i in range(200)])
    for k, key in ((5, "ids5"), (60, "ids60")):
        got = ids_of(db.query(g["q"], top_k=k))
        base = np.argsort(-(db._vectors @ g["q"]))[:k]
        assert got == [str(i) for i in base] == [str(i) for i in g[key][0]]
    assert db._last_topk_strategy == "argsort" and db._last_k_eff == 60


def test_filters_where_ids_better_than(make_db):
    # reference tests/test_task18_prefilter.py, test_task34_prefilter.py, test_more.py:158-173
    db = make_db()
    db.upsert([
        {K_VECTOR: [1.0, 0.0, 0.0], K_ID: "a", "keep": True, "color": "red"},
        {K_VECTOR: [1.0, 0.0, 0.0], K_ID: "b", "keep": False, "color": "blue"},
        {K_VECTOR: [0.0, 1.0, 0.0], K_ID: "c", "keep": True, "color": "green"},
    ])
    q = np.array([1.0, 0.0, 0.0], np.float32)
    assert len(db.query(q, top_k=2, better_than=0.99)) == 2
    assert ids_of(db.query(q, top_k=3, better_than=1.0)) == ["a", "b"]  # >= keeps (Q7)
    assert ids_of(db.query(q, top_k=3, where=lambda d: d.get("keep", False))) == ["a", "c"]
    assert ids_of(db.query(q, top_k=3, where={"color": "blue"})) == ["b"]
    assert ids_of(db.query(q, top_k=3, where={"color": {"$in": ["green", "blue"]}})) == ["b", "c"]
    assert ids_of(db.query(q, top_k=3, ids=["c", "a", "nope"])) == ["a", "c"]
    assert ids_of(db.query(q, top_k=3, ids=["a", "b"], where={"keep": False})) == ["b"]
    assert ids_of(db.query(q, top_k=3, ids=["a", "c"], where=lambda d: d["color"] != "red")) == ["c"]
    # quirk Q2: no candidates -> [[]] even for a 1-D query
    assert db.query(q, ids=["nope"]) == [[]]
    assert db.query(q, where={"color": "none"}) == [[]]
    assert make_db(name="empty").query(q) == [[]]
    assert db.query(np.stack([q, q]), where={"color": "none"}) == [[], []]


def test_k_eff_and_strategy_attributes(make_db):
    # reference tests/test_task48_tuning_knobs.py:39-60, tests/test_task19_adaptive_buffer.py
    db = make_db(dim=4)
    rng = np.random.default_rng(1)
    db.upsert([{K_ID: str(i), K_VECTOR: rng.random(4).astype(np.float32)} for i in range(100)])
    q = rng.random(4).astype(np.float32)
    db._argsort_threshold, db._adaptive_buffer = 0.0, 0
    db.query(q, top_k=10)
    assert db._last_topk_strategy == "argsort"
    db._argsort_threshold = 1.0
    db.query(q, top_k=10)
    assert db._last_topk_strategy == "argpartition" and db._last_k_eff == 10
    db._adaptive_buffer = 7
    res = db.query(q, top_k=5, where=lambda d: True)
    assert db._last_k_eff == 12 and len(res) == 5
    res = db.query(q, top_k=5, where=lambda d: int(d[K_ID]) % 2 == 0)
    assert len(res) == 5 and all(int(r[K_ID]) % 2 == 0 for r in res)


def test_env_and_kwarg_knobs(make_db, monkeypatch):
    # reference tests/test_task48_tuning_knobs.py:17-36
    monkeypatch.setenv("PICOVDB_ADAPTIVE_BUFFER", "7")
    monkeypatch.setenv("PICOVDB_ARGSORT_THRESHOLD", "0.9")
    db = make_db(dim=4, name="env")
    assert db._adaptive_buffer == 7 and abs(db._argsort_threshold - 0.9) < 1e-9
    db = make_db(dim=4, name="kw", adaptive_buffer=11, argsort_threshold=0.1)
    assert db._adaptive_buffer == 11 and abs(db._argsort_threshold - 0.1) < 1e-9


def test_large_k_paging(make_db):
    # k above the fused per-pass limit (128) must still be exact and ordered
    db = make_db(dim=8)
    rng = np.random.default_rng(9)
    v = rng.standard_normal((700, 8)).astype(np.float32)
    db.upsert([{K_VECTOR: v[i], K_ID: i} for i in range(700)])
    q = rng.standard_normal(8).astype(np.float32)
    res = db.query(q, top_k=300)
    base = np.argsort(-(db._vectors @ (q / np.linalg.norm(q))), kind="stable")[:300]
    assert ids_of(res) == base.tolist()


# ------------------------------------------------------------------ golden parity at record level
@pytest.mark.parametrize("name,k", [("gauss_n600_d48.npz", 10), ("gauss_n900_d20_k100.npz", 100)])
def test_golden_records(make_db, name, k):
    g = np.load(os.path.join(GOLDEN, name))
    n, dim = g["raw"].shape
    db = make_db(dim=dim)
    db.upsert([{K_VECTOR: g["raw"][i], K_ID: str(i), "category_id": int(i % 10)} for i in range(n)])
    dead = np.flatnonzero(g["deleted"])
    if dead.size:
        db.delete([str(int(i)) for i in dead])
    cases = {
        "": {},
        "where_eq": {"where": {"category_id": 0}},
        "where_in": {"where": {"category_id": {"$in": [1, 2, 3]}}},
        "where_fn": {"where": lambda d: d["category_id"] % 2 == 0},
        "subset": {"ids": [str(i) for i in range(0, n, 7)]},
        "better": {"better_than": 0.05},
    }
    for key, kwargs in cases.items():
        res = db.query(g["queries"], top_k=k, **kwargs)
        want_ids = g[f"ids_{key}" if key else "ids"]
        want_sc = g[f"scores_{key}" if key else "scores"]
        for qi, rows in enumerate(res):
            kk = int((want_ids[qi] >= 0).sum())
            assert [int(r[K_ID]) for r in rows] == want_ids[qi, :kk].tolist(), (key, qi)
            np.testing.assert_allclose([r[K_METRICS] for r in rows], want_sc[qi, :kk], rtol=1e-5, atol=1e-6)
    single = db.query(g["queries"][0], top_k=k)
    assert [int(r[K_ID]) for r in single] == [int(i) for i in g["ids_single"][0] if i >= 0]


# ------------------------------------------------------------------ persistence
def test_save_reload_roundtrip_and_format(make_db, tmp_path):
    # reference tests/test_pico_vdb.py:38-66; file format of pico_vdb.py:351-371
    db = make_db(dim=4, name="persist")
    rng = np.random.default_rng(2)
    v = rng.random((9, 4)).astype(np.float32)
    db.upsert([{K_VECTOR: v[i], K_ID: f"k{i}", "n": i} for i in range(9)])
    db.delete(["k3"])
    db.store_additional_data(a=1, b="x")
    before = db._vectors.copy()
    db.save()
    base = str(tmp_path / "persist")
    with open(base + ".ids.json") as f:
        assert json.load(f) == [f"k{i}" for i in range(9)]  # deleted rows keep their id (Q8)
    with open(base + ".meta.json") as f:
        meta = json.load(f)
    assert meta["embedding_dim"] == 4 and meta["data"][3] is None and meta["data"][5]["n"] == 5
    assert meta["additional_data"] == {"a": 1, "b": "x"}
    on_disk = np.load(base + ".vecs.npy")
    assert on_disk.dtype == np.float32 and on_disk.shape == (9, 4) and np.all(on_disk[3] == 0)
    np.testing.assert_array_equal(on_disk, before)
    assert not [p for p in os.listdir(tmp_path) if p.endswith(".tmp") or ".tmp." in p]
    db2 = make_db(dim=4, name="persist")
    assert len(db2) == 8 and db2._free == [3] and db2.get_additional_data() == {"a": 1, "b": "x"}
    np.testing.assert_array_equal(db2._vectors, before)
    assert db2._active_indices.tolist() == [0, 1, 2, 4, 5, 6, 7, 8]
    for i in (0, 5, 8):
        r = db2.query(v[i], top_k=1, better_than=0.99)
        assert ids_of(r) == [f"k{i}"]


def test_loads_store_written_by_reference(make_db, tmp_path):
    import shutil

    for suffix in (".ids.json", ".vecs.npy", ".meta.json"):
        shutil.copy(os.path.join(GOLDEN, "refstore" + suffix), str(tmp_path / ("ref" + suffix)))
    with open(os.path.join(GOLDEN, "refstore.expect.json")) as f:
        exp = json.load(f)
    db = make_db(dim=6, name="ref")
    assert len(db) == 11 and db.capacity() == 12 and db.get("doc4") is None
    assert db.get_additional_data() == {"owner": "golden", "version": 3}
    res = db.query(np.asarray(exp["query"], np.float32), top_k=4)
    assert ids_of(res) == ids_of(exp["top4"])
    for a, b in zip(res, exp["top4"]):
        assert a["text"] == b["text"] and a[K_METRICS] == pytest.approx(b[K_METRICS], rel=1e-5, abs=1e-6)
    # saving again reproduces the reference's files (ids / meta byte-identical, vectors equal)
    db.save()
    for suffix in (".ids.json", ".meta.json"):
        with open(str(tmp_path / ("ref" + suffix)), "rb") as f1, open(os.path.join(GOLDEN, "refstore" + suffix), "rb") as f2:
            assert f1.read() == f2.read()
    np.testing.assert_array_equal(np.load(str(tmp_path / "ref.vecs.npy")), np.load(os.path.join(GOLDEN, "refstore.vecs.npy")))


def test_failed_save_leaves_no_temp_files(make_db, tmp_path):
    # reference tests/test_more.py:271-293
    db = make_db(dim=2, name="atomic")
    db.upsert([{K_VECTOR: [1.0, 2.0], K_ID: "x"}])
    with mock.patch("os.replace", side_effect=OSError("boom")):
        with pytest.raises(OSError):
            db.save()
    assert [p for p in os.listdir(tmp_path) if "tmp" in p] == []


def test_vacuum_and_query_after(make_db):
    # reference tests/test_api_ergonomics.py:45-77
    db = make_db(dim=3)
    eye = np.eye(3, dtype=np.float32)
    db.upsert([{K_VECTOR: eye[i], K_ID: f"v{i}"} for i in range(3)])
    db.delete(["v0"])
    db.vacuum()
    assert db.capacity() == 2 and db._free == [] and db._id2idx == {"v1": 0, "v2": 1}
    assert db._active_indices.tolist() == [0, 1] and db._vectors.shape == (2, 3)
    assert ids_of(db.query(eye[2], top_k=1)) == ["v2"]
    db.upsert([{K_VECTOR: eye[0], K_ID: "again"}])
    assert db._id2idx["again"] == 2 and ids_of(db.query(eye[0], top_k=1)) == ["again"]
    db.vacuum()  # nothing to do
    assert db.capacity() == 3


def test_capacity_preallocation(make_db, tmp_path):
    # reference tests/test_memmap_capacity.py
    db = make_db(dim=2, name="cap", use_memmap=True, capacity=5)
    assert os.path.getsize(str(tmp_path / "cap.vecs.npy")) >= 5 * 2 * 4
    assert db.capacity() == 5 and db.count() == 0 and len(db._free) == 5
    db.upsert([{K_ID: str(i), K_VECTOR: [float(i), float(i)]} for i in range(5)])
    assert db.count() == 5 and len(db._free) == 0
    want = np.array([3.0, 3.0], np.float32) / np.linalg.norm(np.array([3.0, 3.0], np.float32))
    np.testing.assert_allclose(db.get("3", include_vector=True)[K_VECTOR], want, rtol=1e-6)
    with pytest.raises(ValueError, match="Database capacity exceeded"):
        db.upsert([{K_ID: "extra", K_VECTOR: [1.0, 1.0]}])
    assert ids_of(db.query(np.array([1.0, 1.0], np.float32), top_k=1)) in (["1"], ["2"], ["3"], ["4"])


def test_fresh_capacity_db_reads_and_saves_zero_rows(make_db, tmp_path):
    # reference pico_vdb.py:286-296: a fresh capacity= DB is an all-zero matrix that can be read
    # (get_all(include_vector=True), _vectors) and saved before anything was upserted
    db = make_db(dim=4, name="capz", capacity=6)
    assert db._vectors.shape == (6, 4) and not db._vectors.any()
    db.save()
    assert np.load(str(tmp_path / "capz.vecs.npy")).shape == (6, 4)
    db.upsert([{K_ID: "a", K_VECTOR: [0, 2, 0, 0]}])
    assert db._vectors.shape == (6, 4) and np.count_nonzero(db._vectors) == 1
    recs = db.get_all(include_vector=True)
    assert ids_of(recs) == ["a"]
    db.save()
    mat = np.load(str(tmp_path / "capz.vecs.npy"))
    assert mat.shape == (6, 4) and np.count_nonzero(mat) == 1


def test_capacity_is_ignored_when_loading_a_larger_store(make_db, tmp_path):
    # reference pico_vdb.py:227-284: load takes the stored row count, whatever capacity= says
    db = make_db(dim=3, name="grow")
    db.upsert([{K_ID: str(i), K_VECTOR: np.eye(3, dtype=np.float32)[i % 3] + i} for i in range(7)])
    db.save()
    again = make_db(dim=3, name="grow", capacity=4)
    assert again.count() == 7
    assert ids_of(again.query(np.eye(3, dtype=np.float32)[0], top_k=1)) == ["0"]


# ------------------------------------------------------------------ getters / counters
def test_getters_and_counters(make_db):
    # reference tests/test_task6_getters_include_vector.py, test_task7, test_task8, test_task32
    db = make_db(dim=2)
    db.upsert([{K_VECTOR: [1.0, 0.0], K_ID: "a", "t": 1}, {K_VECTOR: [0.0, 2.0], K_ID: "b", "t": 2}])
    assert db.get("a") == {K_ID: "a", "t": 1} and db.get("zz") is None
    got = db.get(["b", "zz", "a"], include_vector=True)
    assert ids_of(got) == ["b", "a"]
    np.testing.assert_allclose(got[0][K_VECTOR], [0.0, 1.0], atol=1e-7)
    rec = db.get("a")
    rec["t"] = 99
    assert db.get("a")["t"] == 1  # returned dicts are copies
    with pytest.warns(DeprecationWarning):
        assert db.get_by_id("a")["t"] == 1
    with pytest.warns(DeprecationWarning):
        assert db.size() == 2
    db.delete(["a"])
    assert db.count() == 1 and db.capacity() == 2 and len(db) == 1
    assert db.get_all() == [{K_ID: "b", "t": 2}]
    full = db.get_all(include_deleted=True, include_vector=True)
    assert full[0] == {K_ID: "a"} and full[1][K_ID] == "b" and K_VECTOR in full[1]
    st = db.stats()
    assert (st["active"], st["deleted"], st["total"], st["dim"], st["faiss"]) == (1, 1, 2, 2, False)


def test_upsert_array_bulk(make_db):
    db = make_db(dim=4)
    rng = np.random.default_rng(4)
    v = rng.standard_normal((50, 4)).astype(np.float32)
    ids = db.upsert_array(v)
    assert list(ids) == list(range(50)) and len(db) == 50   # a range: bulk rows cost no per-row objects
    ids2 = db.upsert_array(v[:5], ids=[f"s{i}" for i in range(5)], docs=[{"j": i} for i in range(5)])
    assert db.get("s3")["j"] == 3 and db._id2idx["s0"] == 50 and ids2[0] == "s0"
    np.testing.assert_allclose(np.linalg.norm(db._vectors, axis=1), 1.0, rtol=1e-6)
    assert db.query(v[7], top_k=1)[0][K_ID] == 7
    with pytest.raises(ValueError):
        db.upsert_array(v[:2], ids=[0, 1])


def test_array_level_search(make_db):
    db = make_db(dim=6)
    rng = np.random.default_rng(8)
    v = rng.standard_normal((40, 6)).astype(np.float32)
    db.upsert_array(v)
    scores, rows = db.search(v[:3], top_k=4)
    assert scores.shape == (3, 4) and rows.dtype == np.int64 and rows[:, 0].tolist() == [0, 1, 2]
    assert np.all(np.diff(scores, axis=1) <= 0)
    mask = np.zeros(40, bool)
    mask[[5, 6]] = True
    s2, r2 = db.search(v[0], top_k=4, prefilter=mask)
    assert set(r2[0, :2].tolist()) == {5, 6} and r2[0, 2:].tolist() == [-1, -1] and np.isinf(s2[0, 2:]).all()


# ------------------------------------------------------------------ locking / logging
def test_rwlock_contract():
    # reference tests/test_task9_rwlock.py
    lock = dbmod._RWLock()
    order = []
    with lock.read_lock():
        with lock.read_lock():
            order.append("two readers")
        t = threading.Thread(target=lambda: (lock.acquire_write(), order.append("writer"), lock.release_write()))
        t.start()
        time.sleep(0.05)
        assert order == ["two readers"]
    t.join(1)
    assert order == ["two readers", "writer"]


def test_reader_waits_for_writer_and_concurrent_use(make_db):
    # reference tests/test_task10_apply_rwlocks.py:20-27 and tests/test_task11_snapshot_reads.py
    db = make_db(dim=4)
    rng = np.random.default_rng(3)
    db.upsert([{K_VECTOR: rng.random(4).astype(np.float32), K_ID: str(i)} for i in range(20)])
    seen = []
    with db._rwlock.write_lock():
        t = threading.Thread(target=lambda: seen.append(db.count()))
        t.start()
        time.sleep(0.05)
        assert seen == []
    t.join(1)
    assert seen == [20]
    stop = time.time() + 0.15
    errors = []

    def reader():
        while time.time() < stop:
            try:
                r = db.query(rng.random(4).astype(np.float32), top_k=3)
                assert len(r) <= 3
            except Exception as e:  # pragma: no cover
                errors.append(e)

    def writer():
        i = 100
        while time.time() < stop:
            db.upsert([{K_VECTOR: np.random.rand(4).astype(np.float32), K_ID: str(i)}])
            db.delete([str(i - 1)])
            i += 1

    ts = [threading.Thread(target=reader) for _ in range(2)] + [threading.Thread(target=writer)]
    [t.start() for t in ts]
    [t.join(5) for t in ts]
    assert errors == []


def test_logging_quiet_by_default_and_timed_at_debug(make_db, caplog):
    # reference tests/test_task4_logging.py, tests/test_timing_logs.py
    db = make_db(dim=2)
    with caplog.at_level(logging.WARNING, logger="picovdb"):
        db.upsert([{K_VECTOR: [1.0, 0.0], K_ID: "a"}])
        db.query(np.array([1.0, 0.0], np.float32))
    assert caplog.records == []
    with caplog.at_level(logging.DEBUG, logger="picovdb"):
        db.query(np.array([1.0, 0.0], np.float32))
        db.save()
    msgs = [r.getMessage() for r in caplog.records]
    assert any(m.startswith("query took") for m in msgs) and any(m.startswith("save took") for m in msgs)


@pytest.mark.skipif(not os.path.isdir("/root/reference/tests"), reason="reference tree not present")
def test_reference_suite_passes_against_the_drop_in_class():
    """The reference's own tests, run in place with `import picovdb` resolving to this package."""
    import subprocess
    import sys

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = subprocess.run([sys.executable, os.path.join(root, "tools", "run_reference_tests.py")],
                         capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-2000:]
    assert " passed" in out.stdout and "failed" not in out.stdout
