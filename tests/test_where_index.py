"""The columnar metadata index behind dict ``where`` filters must select exactly the rows the
reference's Python loop selects (picovdb/pico_vdb.py:615-638), through inserts, updates, deletes,
slot reuse, vacuum and reload."""
import numpy as np
import pytest

from picovdb_b200 import K_ID, K_VECTOR, PicoVectorDB

from _host_engine import HostEngine


@pytest.fixture(params=[pytest.param("host"), pytest.param("cuda", marks=pytest.mark.gpu)])
def make_db(request, tmp_path, monkeypatch):
    if request.param == "host":
        monkeypatch.setattr(PicoVectorDB, "_engine_factory", staticmethod(lambda dim, **kw: HostEngine(dim, **kw)))
    made = []

    def factory(dim=4, name="w", **kw):
        d = PicoVectorDB(embedding_dim=dim, storage_file=str(tmp_path / name), no_faiss=True, **kw)
        made.append(d)
        return d

    yield factory
    for d in made:
        d.close()


def loop_mask(db, where, ids=None):
    """The reference's rule, restated as a plain loop."""
    ((key, val),) = where.items()
    rows = range(len(db._ids)) if ids is None else [db._id2idx[i] for i in ids if i in db._id2idx]
    out = np.zeros(len(db._ids), bool)
    for r in rows:
        d = db._docs[r]
        if d is None:
            continue
        if isinstance(val, dict) and set(val) == {"$in"}:
            out[r] = d.get(key) in set(val["$in"])
        else:
            out[r] = d.get(key) == val
    return out


FILTERS = [
    {"cat": 3}, {"cat": 3.0}, {"cat": True}, {"cat": None}, {"cat": "3"}, {"tag": "b"}, {"tag": None},
    {"cat": {"$in": [1, 2, 99]}}, {"tag": {"$in": ["a", None]}}, {"missing": 1}, {"missing": None},
    {"cat": {"$in": []}},
]


def check_all(db):
    for f in FILTERS:
        got = db._candidate_mask(f, None)
        np.testing.assert_array_equal(got, loop_mask(db, f), err_msg=str(f))
    some = [i for i in list(db._id2idx)[::3]] + ["nope"]
    for f in FILTERS[:6]:
        np.testing.assert_array_equal(db._candidate_mask(f, some), loop_mask(db, f, some), err_msg=str(f))


def test_index_tracks_every_mutation(make_db, tmp_path):
    db = make_db()
    rng = np.random.default_rng(0)

    def rec(i):
        d = {K_VECTOR: rng.random(4).astype(np.float32), K_ID: f"r{i}", "cat": i % 5}
        if i % 3:
            d["tag"] = "abc"[i % 3]
        if i % 7 == 0:
            d["cat"] = True  # equals 1 under Python semantics
        return d

    db.upsert([rec(i) for i in range(60)])
    check_all(db)                                   # builds the columns lazily
    assert set(db._columns) >= {"cat", "tag", "missing"}
    db.upsert([rec(i) for i in range(60, 90)])      # appends with live columns
    db.upsert([{K_VECTOR: rng.random(4), K_ID: "r5", "cat": 3, "tag": "b"}])   # update in place
    check_all(db)
    db.delete([f"r{i}" for i in range(0, 90, 4)])
    check_all(db)
    db.upsert([{K_VECTOR: rng.random(4), K_ID: "new1", "cat": 3}, {K_VECTOR: rng.random(4), K_ID: "new2"}])  # slot reuse
    check_all(db)
    db.upsert_array(rng.random((10, 4)).astype(np.float32)) if not db._free else None
    db.vacuum()
    check_all(db)
    db.upsert_array(rng.random((7, 4)).astype(np.float32), docs=[{"cat": 3}] * 7)
    check_all(db)
    db.save()
    db2 = make_db()
    check_all(db2)
    # end to end: the filtered query only returns matching records, best first
    res = db2.query(rng.random(4).astype(np.float32), top_k=5, where={"cat": 3})
    assert res and all(r["cat"] == 3 for r in res)


def test_unhashable_values_fall_back_to_the_loop(make_db):
    db = make_db()
    db.upsert([
        {K_VECTOR: [1, 0, 0, 0], K_ID: "a", "tags": ["x", "y"]},
        {K_VECTOR: [0, 1, 0, 0], K_ID: "b", "tags": "x"},
    ])
    q = np.array([1, 1, 0, 0], np.float32)
    assert [r[K_ID] for r in db.query(q, where={"tags": "x"})] == ["b"]
    assert [r[K_ID] for r in db.query(q, where={"tags": ["x", "y"]})] == ["a"]
    assert not db._columns["tags"].ok


def test_unhashable_filter_value_does_not_disable_the_column(make_db):
    # an unhashable FILTER value says nothing about the stored data: that one query takes the Python
    # loop, the column index (and its device mirror) stays usable for the next one
    db = make_db()
    db.upsert([
        {K_VECTOR: [1, 0, 0, 0], K_ID: "a", "tag": "x"},
        {K_VECTOR: [0, 1, 0, 0], K_ID: "b", "tag": "y"},
    ])
    q = np.array([1, 1, 0, 0], np.float32)
    assert [r[K_ID] for r in db.query(q, where={"tag": "x"})] == ["a"]
    assert db.query(q, where={"tag": ["x", "y"]}) == [[]]          # no stored value equals the list
    assert db.query(q, where={"tag": {"nested": 1}}) == [[]]
    assert db._columns["tag"].ok
    assert [r[K_ID] for r in db.query(q, where={"tag": "y"})] == ["b"]


def _ids_of(res):
    if res and isinstance(res[0], list):   # no candidates: the reference returns [[]] even for one query
        return [[r[K_ID] for r in lst] for lst in res]
    return [r[K_ID] for r in res]


def test_device_where_matches_host_mask_path(make_db, monkeypatch):
    """query(where={...}) through the device-resident code column (pvdb_search_where) returns what
    the host-mask path returns, through every kind of mutation, with ids=, and keeps quirk Q2's
    k_eff = min(top_k + buffer, candidates)."""
    db = make_db(dim=8)
    rng = np.random.default_rng(3)
    assert hasattr(db._engine, "search_where")

    def rec(i):
        return {K_VECTOR: rng.standard_normal(8).astype(np.float32), K_ID: f"r{i}", "cat": i % 4,
                "tag": "xyz"[i % 3] if i % 5 else None}

    def both(q, **kw):
        got = db.query(q, **kw)
        k_eff = db._last_k_eff
        with monkeypatch.context() as m:
            m.setattr(PicoVectorDB, "_device_where", lambda self, where, ids: None)
            want = db.query(q, **kw)
            assert _ids_of(got) == _ids_of(want), kw
            if want:
                assert db._last_k_eff == k_eff
        return got

    def sweep():
        q = rng.standard_normal(8).astype(np.float32)
        for f in ({"cat": 1}, {"cat": {"$in": [0, 3]}}, {"tag": "y"}, {"tag": None}, {"cat": 77}, {"nokey": 1},
                  {"cat": {"$in": []}}, {"cat": 1.0}):
            both(q, top_k=6, where=f)
        some = list(db._id2idx)[::2] + ["ghost"]
        both(q, top_k=4, where={"cat": 2}, ids=some)
        both(q, top_k=500, where={"cat": 2})        # more than there are candidates
        res = db.query(np.stack([q, -q]), top_k=3, where={"cat": 0})
        assert len(res) == 2 and all(r["cat"] == 0 for lst in res for r in lst)

    db.upsert([rec(i) for i in range(200)])
    sweep()
    db.upsert([rec(i) for i in range(200, 260)])                       # appended rows reach the column
    db.upsert([{K_VECTOR: rng.standard_normal(8), K_ID: "r7", "cat": 1, "tag": "y"}])  # update in place
    sweep()
    db.delete([f"r{i}" for i in range(0, 260, 3)])
    sweep()
    db.upsert([{K_VECTOR: rng.standard_normal(8), K_ID: f"n{i}", "cat": 1} for i in range(40)])  # slot reuse
    sweep()
    db.vacuum()                                                        # compaction drops the device columns
    assert all(c.dev_slot is None for c in db._columns.values())
    sweep()
    assert any(c.dev_slot is not None for c in db._columns.values())


def test_device_where_column_eviction(make_db):
    """More filter keys than device column slots: the oldest column is evicted and re-sent on use."""
    db = make_db(dim=4)
    rng = np.random.default_rng(5)
    keys = [f"k{j}" for j in range(20)]
    db.upsert([{K_VECTOR: rng.standard_normal(4), K_ID: f"r{i}", **{k: (i + j) % 3 for j, k in enumerate(keys)}}
               for i in range(50)])
    q = rng.standard_normal(4).astype(np.float32)
    for rnd in range(2):
        for j, k in enumerate(keys):
            res = db.query(q, top_k=50, where={k: 1})
            want = {f"r{i}" for i in range(50) if (i + j) % 3 == 1}
            assert set(_ids_of(res)) == want
    slots = [c.dev_slot for c in db._columns.values() if c.dev_slot is not None]
    assert len(slots) == len(set(slots)) <= 16


def test_concurrent_filtered_queries_share_the_column_index(make_db):
    """Queries run under the read lock, concurrently; the lazily built column index and its device
    mirror are shared state -- many threads filtering on many keys must all get the loop's answer."""
    import threading

    db = make_db(dim=8)
    rng = np.random.default_rng(9)
    keys = [f"k{j}" for j in range(20)]  # more keys than device column slots: eviction happens under load
    n = 300
    db.upsert([{K_VECTOR: rng.standard_normal(8), K_ID: f"r{i}", **{k: (i * (j + 1)) % 4 for j, k in enumerate(keys)}}
               for i in range(n)])
    q = rng.standard_normal(8).astype(np.float32)
    errors = []

    def worker(t):
        try:
            for rep_ in range(6):
                for j, k in enumerate(keys):
                    if (j + t) % 3:
                        continue
                    want = {f"r{i}" for i in range(n) if (i * (j + 1)) % 4 == 1}
                    res = db.query(q, top_k=n, where={k: 1})
                    if res and isinstance(res[0], list):   # no candidates: the reference's [[]]
                        res = res[0]
                    got = {r[K_ID] for r in res}
                    if got != want:
                        errors.append((t, k, len(got), len(want)))
        except Exception as e:  # noqa: BLE001
            errors.append((t, repr(e)))

    threads = [threading.Thread(target=worker, args=(t,)) for t in range(6)]
    for th in threads:
        th.start()
    for th in threads:
        th.join()
    assert not errors, errors[:5]
