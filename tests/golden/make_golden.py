"""Generate golden input/output fixtures by running the UNMODIFIED reference class.

Run in the build container only (``/root/reference`` does not exist on the GPU box):

    python tests/golden/make_golden.py

Imports ``picovdb`` from /root/reference (``no_faiss=True`` -> the NumPy path,
picovdb/pico_vdb.py:670-714), feeds it seeded inputs and stores inputs + outputs as small
``.npz`` / ``.json`` files next to this script.  The fixtures pin ``oracle/picovdb_oracle.py``
(tests/test_oracle_golden.py) and are the expected values of the ``-m gpu`` parity tests.
Every store used here has a sorted ``_active_indices`` (fresh in-order inserts), so reference
quirk Q1 does not affect the recorded ids.
"""
import json
import os
import sys
import tempfile

import numpy as np

sys.path.insert(0, "/root/reference")
from picovdb import PicoVectorDB, K_ID, K_VECTOR, K_METRICS  # noqa: E402
from picovdb.pico_vdb import _normalize  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def _results_to_arrays(res, k):
    """list[list[dict]] -> (ids as int array padded with -1, scores padded with -inf)."""
    nq = len(res)
    ids = np.full((nq, k), -1, dtype=np.int64)
    sc = np.full((nq, k), -np.inf, dtype=np.float32)
    for qi, rows in enumerate(res):
        for j, r in enumerate(rows):
            ids[qi, j] = int(r[K_ID])
            sc[qi, j] = np.float32(r[K_METRICS])
    return ids, sc


def case_task20(tmp):
    # tests/test_task20_argsort_vs_argpartition.py:12-36 with its own seed and shapes
    dim, n = 16, 200
    db = PicoVectorDB(embedding_dim=dim, storage_file=os.path.join(tmp, "ap"), no_faiss=True)
    rng = np.random.default_rng(0)
    vecs = rng.random((n, dim), dtype=np.float32)
    db.upsert([{K_VECTOR: vecs[i], K_ID: str(i)} for i in range(n)])
    q = rng.random(dim, dtype=np.float32)
    small = db.query(q, top_k=5)
    large = db.query(q, top_k=60)
    ids5, sc5 = _results_to_arrays([small], 5)
    ids60, sc60 = _results_to_arrays([large], 60)
    return dict(raw=vecs, q=q, store=np.asarray(db._vectors), ids5=ids5, sc5=sc5,
                ids60=ids60, sc60=sc60)


def case_gauss(tmp, name, n, dim, nq, k, seed, delete_frac=0.0):
    rng = np.random.default_rng(seed)
    raw = rng.standard_normal((n, dim)).astype(np.float32)
    raw[n // 3] = 0.0  # a zero row: must be stored as e0
    db = PicoVectorDB(embedding_dim=dim, storage_file=os.path.join(tmp, name), no_faiss=True)
    db.upsert([{K_VECTOR: raw[i], K_ID: str(i), "category_id": int(i % 10)} for i in range(n)])
    deleted = np.zeros(n, dtype=bool)
    if delete_frac > 0:
        sel = np.random.default_rng(1).choice(n, size=int(n * delete_frac), replace=False)
        db.delete([str(int(i)) for i in sel])
        deleted[sel] = True
    qs = np.random.default_rng(seed + 1000).standard_normal((nq, dim)).astype(np.float32)
    if nq > 2:
        qs[1] = 0.0  # zero query -> e0
    out = dict(raw=raw, deleted=deleted, queries=qs, store=np.asarray(db._vectors))
    res = db.query(qs, top_k=k)
    out["ids"], out["scores"] = _results_to_arrays(res, k)
    # single-query form (1-D input)
    res1 = db.query(qs[0], top_k=k)
    out["ids_single"], out["scores_single"] = _results_to_arrays([res1], k)
    # dict prefilter (10 %), $in prefilter (~30 %), callable (50 %), ids subset, better_than
    res = db.query(qs, top_k=k, where={"category_id": 0})
    out["ids_where_eq"], out["scores_where_eq"] = _results_to_arrays(res, k)
    res = db.query(qs, top_k=k, where={"category_id": {"$in": [1, 2, 3]}})
    out["ids_where_in"], out["scores_where_in"] = _results_to_arrays(res, k)
    res = db.query(qs, top_k=k, where=lambda d: d["category_id"] % 2 == 0)
    out["ids_where_fn"], out["scores_where_fn"] = _results_to_arrays(res, k)
    subset = [str(i) for i in range(0, n, 7)]
    res = db.query(qs, top_k=k, ids=subset)
    out["ids_subset"], out["scores_subset"] = _results_to_arrays(res, k)
    res = db.query(qs, top_k=k, better_than=0.05)
    out["ids_better"], out["scores_better"] = _results_to_arrays(res, k)
    return out


def case_normalize():
    rng = np.random.default_rng(7)
    vecs = {}
    for dim in (1, 2, 3, 5, 16, 384, 1024):
        m = (rng.standard_normal((6, dim)) * rng.uniform(1e-3, 1e3)).astype(np.float32)
        m[2] = 0.0
        vecs[f"in_{dim}"] = m
        vecs[f"out_{dim}"] = np.stack([_normalize(v) for v in m])
    vecs["in_34"] = np.array([[3.0, 4.0]], dtype=np.float32)
    vecs["out_34"] = np.stack([_normalize(v) for v in vecs["in_34"]])
    return vecs


def case_records(tmp):
    """Record-level behaviours pinned by the reference tests, captured as JSON."""
    out = {}
    # tests/test_more.py:133-155 orthonormal basis
    db = PicoVectorDB(embedding_dim=3, storage_file=os.path.join(tmp, "basis"), no_faiss=True)
    vs = np.eye(3, dtype=np.float32)
    db.upsert([{K_VECTOR: v, K_ID: str(i)} for i, v in enumerate(vs)])
    out["basis_single"] = db.query(np.array([0.9, 0.1, 0], dtype=np.float32), top_k=2)
    out["basis_batch"] = db.query(np.stack([vs[2], vs[1]]), top_k=1)
    # tests/test_task5_zero_vector_normalization.py:17-41
    out["basis_zero_query"] = db.query(np.zeros(3, dtype=np.float32), top_k=3)
    dbz = PicoVectorDB(embedding_dim=3, storage_file=os.path.join(tmp, "z"), no_faiss=True)
    dbz.upsert([{K_VECTOR: np.zeros(3, dtype=np.float32), K_ID: "z"}])
    out["zero_upsert_zero_query"] = dbz.query(np.zeros(3, dtype=np.float32), top_k=1)
    # quirk Q2: no candidates -> [[]] even for a single query
    out["empty_db_single"] = PicoVectorDB(
        embedding_dim=3, storage_file=os.path.join(tmp, "e"), no_faiss=True
    ).query(np.ones(3, dtype=np.float32))
    out["missing_ids_single"] = db.query(np.ones(3, dtype=np.float32), ids=["nope"])
    out["where_nomatch_single"] = db.query(np.ones(3, dtype=np.float32), where={"x": 1})
    # better_than keeps score >= threshold (Q7)
    out["better_than_1"] = db.query(vs[0], top_k=3, better_than=1.0)
    # tests/test_task2_numpy_query_active_indices.py:6-41 shape, seeded
    rng = np.random.default_rng(5)
    v30 = rng.random((30, 8), dtype=np.float32)
    db2 = PicoVectorDB(embedding_dim=8, storage_file=os.path.join(tmp, "a"), no_faiss=True)
    db2.upsert([{K_VECTOR: v30[i], K_ID: f"id{i}"} for i in range(30)])
    db2.delete([f"id{i}" for i in range(20)])
    q = rng.random(8, dtype=np.float32)
    out["task2_vectors"] = v30.tolist()
    out["task2_query"] = q.tolist()
    out["task2_top25"] = db2.query(q, top_k=25)
    # tests/test_task48_tuning_knobs.py:39-60 debug attributes
    db2.query(q, top_k=3, where=lambda d: True)
    out["task48_k_eff_filtered"] = db2._last_k_eff
    db2.query(q, top_k=3)
    out["task48_strategy_small"] = db2._last_topk_strategy
    out["task48_k_eff_plain"] = db2._last_k_eff
    return out


def case_refstore():
    """A store saved by the reference itself (3 files, with one deleted row): load-compat fixture."""
    base = os.path.join(HERE, "refstore")
    for suffix in (".ids.json", ".vecs.npy", ".meta.json"):
        if os.path.exists(base + suffix):
            os.remove(base + suffix)
    rng = np.random.default_rng(11)
    raw = rng.standard_normal((12, 6)).astype(np.float32)
    db = PicoVectorDB(embedding_dim=6, storage_file=base, no_faiss=True)
    db.upsert([{K_VECTOR: raw[i], K_ID: f"doc{i}", "text": f"t{i}", "n": i} for i in range(12)])
    db.delete(["doc4"])
    db.store_additional_data(owner="golden", version=3)
    db.save()
    q = rng.standard_normal(6).astype(np.float32)
    res = db.query(q, top_k=4)
    with open(base + ".expect.json", "w") as f:
        json.dump({"raw": raw.tolist(), "query": q.tolist(), "top4": res}, f, indent=1, sort_keys=True)


def main():
    case_refstore()
    with tempfile.TemporaryDirectory() as tmp:
        np.savez_compressed(os.path.join(HERE, "task20.npz"), **case_task20(tmp))
        np.savez_compressed(
            os.path.join(HERE, "gauss_n600_d48.npz"),
            **case_gauss(tmp, "g1", 600, 48, 5, 10, seed=123),
        )
        np.savez_compressed(
            os.path.join(HERE, "gauss_n400_d384_del30.npz"),
            **case_gauss(tmp, "g2", 400, 384, 4, 10, seed=124, delete_frac=0.3),
        )
        np.savez_compressed(
            os.path.join(HERE, "gauss_n900_d20_k100.npz"),
            **case_gauss(tmp, "g3", 900, 20, 3, 100, seed=125, delete_frac=0.1),
        )
        np.savez_compressed(os.path.join(HERE, "normalize.npz"), **case_normalize())
        with open(os.path.join(HERE, "records.json"), "w") as f:
            json.dump(case_records(tmp), f, indent=1, sort_keys=True)
    print("golden fixtures written to", HERE)


if __name__ == "__main__":
    main()
