"""Pin oracle/picovdb_oracle.py against outputs of the unmodified reference (tests/golden/)."""
import json
import os

import numpy as np
import pytest

from oracle import picovdb_oracle as O

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")


def _load(name):
    return dict(np.load(os.path.join(GOLDEN, name)))


def test_normalize_bit_exact_with_reference():
    g = _load("normalize.npz")
    for key in [k for k in g if k.startswith("in_")]:
        dim = key[3:]
        got = np.stack([O.normalize(v) for v in g[key]])
        assert got.dtype == np.float32
        assert np.array_equal(got, g[f"out_{dim}"]), f"dim {dim}"
    assert np.allclose(g["out_34"], [[0.6, 0.8]], rtol=1e-6)  # tests/test_more.py:258-260


def test_task20_seeded_golden():
    # reference tests/test_task20_argsort_vs_argpartition.py:12-36: both top-k strategies must
    # equal argsort(-(V @ q))[:k]
    g = _load("task20.npz")
    store = O.normalize_rows(g["raw"])
    assert np.array_equal(store, g["store"])
    qn, single = O.prepare_queries(g["q"], 16)
    assert single
    for k, ids_key, sc_key in ((5, "ids5", "sc5"), (60, "ids60", "sc60")):
        s, r = O.search(store, qn, k)
        assert np.array_equal(r, g[ids_key])
        assert np.allclose(s, g[sc_key], rtol=1e-6)
        base = np.argsort(-(g["store"] @ g["q"]))[:k]
        assert np.array_equal(r[0], base)


@pytest.mark.parametrize(
    "name,k",
    [("gauss_n600_d48.npz", 10), ("gauss_n400_d384_del30.npz", 10), ("gauss_n900_d20_k100.npz", 100)],
)
def test_gauss_fixtures(name, k):
    g = _load(name)
    n, dim = g["raw"].shape
    store = O.normalize_rows(g["raw"])
    store[g["deleted"]] = 0  # delete zero-fills the row (pico_vdb.py:523)
    assert np.array_equal(store, g["store"])
    active = ~g["deleted"]
    qn, _ = O.prepare_queries(g["queries"], dim)
    cat = np.arange(n) % 10

    def check(prefix, prefilter, better_than=None):
        s, r = O.search(store, qn, k, active, prefilter)
        if better_than is not None:
            keep = s >= better_than
            s = np.where(keep, s, -np.inf)
            r = np.where(keep, r, -1)
        assert np.array_equal(r, g[f"ids_{prefix}" if prefix else "ids"]), prefix
        ref = g[f"scores_{prefix}" if prefix else "scores"]
        fin = np.isfinite(ref)
        assert np.array_equal(fin, np.isfinite(s))
        assert np.allclose(s[fin], ref[fin], rtol=2e-6, atol=1e-7), prefix

    check("", None)
    check("where_eq", cat == 0)
    check("where_in", np.isin(cat, [1, 2, 3]))
    check("where_fn", cat % 2 == 0)
    subset = np.zeros(n, dtype=bool)
    subset[::7] = True
    check("subset", subset)
    check("better", None, better_than=0.05)
    # 1-D query form
    q1, single = O.prepare_queries(g["queries"][0], dim)
    assert single
    s, r = O.search(store, q1, k, active)
    assert np.array_equal(r, g["ids_single"])
    # chunked variant gives the same answer
    s2, r2 = O.search_chunked(store, qn, k, active, None, chunk_rows=97)
    s1, r1 = O.search(store, qn, k, active)
    assert np.array_equal(r1, r2)
    assert np.allclose(s1, s2, rtol=2e-6, atol=1e-7)


def test_record_level_golden():
    with open(os.path.join(GOLDEN, "records.json")) as f:
        g = json.load(f)

    def ids_of(rows):
        return [r["_id_"] for r in rows]

    db = O.OracleDB(3)
    eye = np.eye(3, dtype=np.float32)
    db.upsert([{"_vector_": v, "_id_": str(i)} for i, v in enumerate(eye)])
    res = db.query(np.array([0.9, 0.1, 0], dtype=np.float32), top_k=2)
    assert ids_of(res) == ids_of(g["basis_single"]) == ["0", "1"]
    for a, b in zip(res, g["basis_single"]):
        assert a["_metrics_"] == pytest.approx(b["_metrics_"], rel=1e-6)
    res = db.query(np.stack([eye[2], eye[1]]), top_k=1)
    assert [ids_of(r) for r in res] == [ids_of(r) for r in g["basis_batch"]]
    assert ids_of(db.query(np.zeros(3, np.float32), top_k=3))[0] == ids_of(g["basis_zero_query"])[0] == "0"
    # quirk Q2
    assert O.OracleDB(3).query(np.ones(3, np.float32)) == g["empty_db_single"] == [[]]
    assert db.query(np.ones(3, np.float32), ids=["nope"]) == g["missing_ids_single"] == [[]]
    assert db.query(np.ones(3, np.float32), where={"x": 1}) == g["where_nomatch_single"] == [[]]
    # Q7: better_than keeps >=
    assert ids_of(db.query(eye[0], top_k=3, better_than=1.0)) == ids_of(g["better_than_1"]) == ["0"]
    dbz = O.OracleDB(3)
    dbz.upsert([{"_vector_": np.zeros(3, np.float32), "_id_": "z"}])
    rz = dbz.query(np.zeros(3, np.float32), top_k=1)
    assert rz[0]["_id_"] == "z" and rz[0]["_metrics_"] == pytest.approx(1.0, rel=1e-5)
    assert g["zero_upsert_zero_query"][0]["_metrics_"] == pytest.approx(1.0, rel=1e-5)
    # active-only (reference tests/test_task2_numpy_query_active_indices.py)
    v30 = np.asarray(g["task2_vectors"], dtype=np.float32)
    db2 = O.OracleDB(8)
    db2.upsert([{"_vector_": v30[i], "_id_": f"id{i}"} for i in range(30)])
    db2.delete([f"id{i}" for i in range(20)])
    q = np.asarray(g["task2_query"], dtype=np.float32)
    got = db2.query(q, top_k=25)
    assert ids_of(got) == ids_of(g["task2_top25"]) and len(got) == 10
    db2.query(q, top_k=3, where=lambda d: True)
    assert db2.last_k_eff == g["task48_k_eff_filtered"]
    db2.query(q, top_k=3)
    assert db2.last_k_eff == g["task48_k_eff_plain"]
    assert db2.last_strategy == g["task48_strategy_small"]


def test_merge_topk_matches_unsharded():
    rng = np.random.default_rng(3)
    v = O.normalize_rows_fast(rng.standard_normal((1000, 32)).astype(np.float32))
    qn, _ = O.prepare_queries(rng.standard_normal((6, 32)).astype(np.float32), 32)
    s_all, r_all = O.search(v, qn, 10)
    parts_s, parts_r = [], []
    for r0 in range(0, 1000, 250):
        s, r = O.search(v[r0 : r0 + 250], qn, 10)
        parts_s.append(s)
        parts_r.append(r + r0)
    s_m, r_m = O.merge_topk(parts_s, parts_r, 10)
    assert np.array_equal(r_m, r_all)
    assert np.allclose(s_m, s_all, rtol=2e-6)


def test_compare_topk_rule():
    ref_s = np.array([[0.9, 0.8, 0.8000001, 0.5]], dtype=np.float32)
    ref_r = np.array([[1, 2, 3, 4]], dtype=np.int64)
    swapped = np.array([[1, 3, 2, 4]], dtype=np.int64)
    stats = O.compare_topk(ref_s, swapped, ref_s, ref_r, rtol=1e-5)
    assert stats["rank_swaps"] == 2
    with pytest.raises(AssertionError):
        O.compare_topk(ref_s, np.array([[2, 1, 3, 4]]), ref_s, ref_r, rtol=1e-5)
    with pytest.raises(AssertionError):
        O.compare_topk(ref_s * 1.001, ref_r, ref_s, ref_r, rtol=1e-5)
