"""Host bookkeeping for 10^7-10^8-row stores: implicit ids / lazy documents
(``_rows.py``), ``upsert_array`` on stores with free slots / ``capacity=``, and the compact save form.
Host-engine tests run here; the ``cuda`` parameter and the 5M-row test need a B200."""
import json
import os
import tracemalloc

import numpy as np
import pytest

from picovdb_b200 import K_ID, K_METRICS, K_VECTOR, PicoVectorDB
from picovdb_b200 import db as dbmod
from picovdb_b200._rows import IdMap, RowSeq

from _host_engine import HostEngine


@pytest.fixture(params=[pytest.param("host"), pytest.param("cuda", marks=pytest.mark.gpu)])
def make_db(request, tmp_path, monkeypatch):
    if request.param == "host":
        monkeypatch.setattr(PicoVectorDB, "_engine_factory", staticmethod(lambda dim, **kw: HostEngine(dim, **kw)))
    made = []

    def factory(dim=8, name="bulk", **kw):
        d = PicoVectorDB(embedding_dim=dim, storage_file=str(tmp_path / name), **kw)
        made.append(d)
        return d

    yield factory
    for d in made:
        d.close()


def _gauss(n, dim, seed):
    return np.random.default_rng(seed).standard_normal((n, dim)).astype(np.float32)


# ------------------------------------------------------------------ the containers
def test_rowseq_behaves_like_a_list():
    seq = RowSeq(lambda i: {"id": i})
    seq.extend([{"id": "a"}, None])
    seq.extend_range(100, 5)
    seq.extend_range(105, 3)          # continues the range: still one chunk
    seq.append({"id": "z"})
    assert len(seq) == 11 and seq.implicit_rows == 8 and seq.implicit_ranges() == [(2, 8, 100)]
    assert seq[0] == {"id": "a"} and seq[1] is None and seq[2] == {"id": 100} and seq[9] == {"id": 107}
    assert seq[-1] == {"id": "z"} and seq[2:4] == [{"id": 100}, {"id": 101}]
    seq[3] = None                     # override inside the range (a delete)
    seq[0] = {"id": "b"}
    assert seq[3] is None and seq[0] == {"id": "b"} and list(seq)[3] is None
    assert seq == [{"id": "b"}, None, {"id": 100}, None] + [{"id": 102 + i} for i in range(6)] + [{"id": "z"}]
    with pytest.raises(IndexError):
        seq[11]
    kept = seq.take_sorted(np.array([0, 2, 4, 5, 7, 10]))
    assert kept == [{"id": "b"}, {"id": 100}, {"id": 102}, {"id": 103}, {"id": 105}, {"id": "z"}]
    assert kept.implicit_ranges() == [(1, 1, 100), (2, 2, 102), (4, 1, 105)]
    again = RowSeq.from_compact(seq._make, json.loads(json.dumps(seq.to_compact())))
    assert again == seq and again.implicit_rows == 8


def test_rowseq_getter_matches_getitem():
    """``getter()`` is what query() calls per returned row: the list's own __getitem__ for a store of
    explicit rows, a closure over the chunk table otherwise -- same values as ``seq[row]`` everywhere."""
    plain = RowSeq(lambda i: {"_id_": i}, [{"a": i} for i in range(50)])
    assert plain.getter().__self__ is plain._chunks[0]          # no wrapper at all
    mixed = RowSeq(lambda i: {"_id_": i}, [{"a": i} for i in range(5)])
    mixed.extend_range(100, 1000)
    mixed.extend([{"b": i} for i in range(7)])
    mixed.extend_range(5000, 3)
    mixed[17] = {"over": 1}
    mixed[2] = None
    get = mixed.getter()
    for row in range(len(mixed)):
        assert get(row) == mixed[row]
    assert get(17) == {"over": 1} and get(2) is None and get(5 + 999) == {"_id_": 1099} and get(1012) == {"_id_": 5000}


def test_idmap_behaves_like_a_dict():
    m = IdMap()
    m["a"] = 0
    m.add_range(10, 1000, 1)          # ids 10..1009 -> rows 1..1000
    assert len(m) == 1001 and m["a"] == 0 and m[10] == 1 and m.get(1009) == 1000 and m.get(1010) is None
    assert 500 in m and 5 not in m and "zz" not in m and [1, 2] not in m and m.get(3.5) is None
    assert m.pop(11) == 2 and 11 not in m and len(m) == 1000 and m.pop(11, "gone") == "gone"
    m[12] = 7777                      # re-pointing an id of the range: the explicit entry wins
    assert m[12] == 7777 and len(m) == 1000
    assert m.overlaps(1009, 5) and not m.overlaps(1010, 5) and not m.overlaps(11, 1) and m.overlaps(12, 1)
    with pytest.raises(ValueError):
        m.add_range(1000, 20, 5000)
    rows = m.sorted_rows()
    assert rows.size == 1000 and rows[0] == 0 and 2 not in rows and 7777 in rows
    assert dict(m)[10] == 1 and sorted(m.values()) == rows.tolist()
    assert m == dict(m.items())


# ------------------------------------------------------------------ the class on top of them
def test_bulk_rows_have_no_per_row_objects_and_lazy_documents(make_db):
    n, dim = 200_000, 8
    db = make_db(dim=dim)
    db.upsert([{K_VECTOR: np.ones(dim), K_ID: "first", "tag": "x"}])
    vecs = _gauss(n, dim, 1)
    tracemalloc.start()
    before = tracemalloc.get_traced_memory()[0]
    ids = db.upsert_array(vecs)
    grown = tracemalloc.get_traced_memory()[0] - before
    tracemalloc.stop()
    assert isinstance(ids, range) and ids == range(1, n + 1)
    # (the host test engine keeps the vectors in numpy: vecs.nbytes; per-row dicts / ints / dict
    # entries would add >= 200 bytes per row = 40 MB here)
    assert grown < vecs.nbytes + 2_000_000, f"bulk ingest allocated {grown} host bytes"
    assert len(db) == n + 1 and db.capacity() == n + 1 and db._ids.implicit_rows == n
    assert db._docs[5] == {K_ID: 5} and db._id2idx[n] == n and db._ids[n] == n
    hit = db.query(vecs[41], top_k=3)
    assert hit[0][K_ID] == 42 and set(hit[0]) == {K_ID, K_METRICS}
    # (get() takes one str id or a list of ids, as in the reference, pico_vdb.py:927-957)
    assert db.get([42]) == [{K_ID: 42}] and db.get([3, "first", 10**9]) == [{K_ID: 3}, {K_ID: "first", "tag": "x"}]
    # updates / deletes inside the range, ids / where filters, explicit rows after the range
    db.upsert([{K_VECTOR: vecs[0] * -1, K_ID: 7, "tag": "y"}])
    assert db._docs[7] == {K_ID: 7, "tag": "y"} and len(db) == n + 1
    assert db.delete([8, 9, "nope"]) == [8, 9] and len(db) == n - 1 and db._docs[8] is None
    assert 8 not in db._id2idx and db._ids[8] == 8          # a deleted slot keeps its id (pico_vdb.py:522)
    db.upsert([{K_VECTOR: vecs[100], K_ID: "late"}])         # reuses a freed slot of the range
    assert db._id2idx["late"] in (8, 9)
    assert [r[K_ID] for r in db.query(vecs[100], top_k=2)] in ([101, "late"], ["late", 101])
    assert [r[K_ID] for r in db.query(vecs[0], top_k=5, where={"tag": "y"})] == [7]
    assert [r[K_ID] for r in db.query(vecs[20], top_k=5, ids=[21, 22, 8])][0] == 21
    act = db._active_indices
    assert act.size == len(db) and set(act.tolist()) == set(db._id2idx.values())
    with pytest.raises(ValueError):
        db.upsert_array(vecs[:2], ids=[5, 6])                # ids of the range are taken
    more = db.upsert_array(vecs[:10])                        # one freed slot left: it is filled first
    assert list(more) == list(range(n + 1, n + 11)) and db._free == []
    assert db.query(vecs[3], top_k=1)[0][K_ID] in (4, n + 4)
    tail = db.upsert_array(vecs[:10])                        # no free slots: a second implicit range
    assert tail == range(n + 11, n + 21) and db._ids.implicit_rows == n + 10   # ids move past every id seen


def test_bulk_save_load_vacuum_roundtrip(make_db, monkeypatch):
    n, dim = 5000, 8
    vecs = _gauss(n, dim, 2)
    for threshold in (10**9, 100):        # reference-format lists, then the compact range form
        monkeypatch.setattr(dbmod, "COMPACT_ROWS_THRESHOLD", threshold)
        name = f"s{threshold}"
        db = make_db(dim=dim, name=name)
        db.upsert([{K_VECTOR: np.ones(dim), K_ID: "first", "tag": "x"}])
        db.upsert_array(vecs)
        db.delete([10, 11, 4000])
        db.upsert([{K_VECTOR: vecs[10], K_ID: 20, "note": "updated"}])
        want = [r[K_ID] for r in db.query(vecs[123], top_k=5)]
        db.save()
        with open(db._path + ".ids.json") as f:
            on_disk = json.load(f)
        assert isinstance(on_disk, dict) == (threshold == 100)
        if threshold != 100:
            assert on_disk[:3] == ["first", 1, 2] and len(on_disk) == n + 1   # what the reference would load
        again = make_db(dim=dim, name=name)
        assert len(again) == len(db) and [r[K_ID] for r in again.query(vecs[123], top_k=5)] == want
        assert again.get([20])[0]["note"] == "updated" and again.get([10]) == [] and again.get("first")["tag"] == "x"
        assert sorted(again._free) == sorted(db._free)
        assert again._ids.implicit_rows == (n if threshold == 100 else 0)
        again.vacuum()
        assert len(again) == len(db) and again.capacity() == len(db) and again._free == []
        assert [r[K_ID] for r in again.query(vecs[123], top_k=5)] == want
        assert again.get([4001, 4000]) == [{K_ID: 4001}]
        assert again._id2idx[4001] == again._ids[:].index(4001)


def test_upsert_array_fills_free_slots_and_respects_capacity(make_db):
    dim = 4
    db = make_db(dim=dim, name="cap", capacity=6)
    v = _gauss(8, dim, 3)
    ids = db.upsert_array(v[:4], ids=["a", "b", "c", "d"])
    assert ids == ["a", "b", "c", "d"] and len(db) == 4 and len(db._free) == 2
    assert db.query(v[2], top_k=1)[0][K_ID] == "c"
    with pytest.raises(ValueError, match="Database capacity exceeded"):
        db.upsert_array(v[4:8])
    assert len(db._free) == 2
    db.delete(["b"])
    got = db.upsert_array(v[4:7], docs=[{"j": j} for j in range(3)])
    assert len(db) == 6 and len(db._free) == 0 and db.get([got[1]])[0]["j"] == 1
    assert db.query(v[5], top_k=1)[0][K_ID] == got[1]
    grow = make_db(dim=dim, name="grow")
    grow.upsert_array(v[:3])
    grow.delete([1])
    out = grow.upsert_array(v[3:6], ids=["x", "y", "z"])     # one freed slot, two appended rows
    assert out == ["x", "y", "z"] and grow._id2idx["x"] == 1 and grow.capacity() == 5
    assert grow.query(v[4], top_k=1)[0][K_ID] == "y"


@pytest.mark.gpu
def test_five_million_bulk_rows_through_the_class(tmp_path):
    """5M x 64 rows through PicoVectorDB.upsert_array + query(): bounded host memory, correct ids."""
    import resource

    n, dim = 5_000_000, 64
    db = PicoVectorDB(embedding_dim=dim, storage_file=str(tmp_path / "big"), bf16_mirror=True)
    rss0 = resource.getrusage(resource.RUSAGE_SELF).ru_maxrss
    rng = np.random.default_rng(5)
    keep = {}
    for c0 in range(0, n, 500_000):
        block = rng.standard_normal((500_000, dim)).astype(np.float32)
        db.upsert_array(block)
        keep[c0 + 17] = block[17].copy()
    rss1 = resource.getrusage(resource.RUSAGE_SELF).ru_maxrss
    assert (rss1 - rss0) < 1_500_000, f"host RSS grew by {(rss1 - rss0) / 1e6:.2f} GB for {n} bulk rows"
    assert len(db) == n and db._ids.implicit_rows == n
    qs = np.stack(list(keep.values()))
    res = db.query(qs, top_k=3)
    assert [r[0][K_ID] for r in res] == list(keep.keys())
    assert all(abs(r[0][K_METRICS] - 1.0) < 1e-5 for r in res)
    db.delete([17, 18])
    assert db.query(keep[17], top_k=1)[0][K_ID] != 17
    db.save()
    assert os.path.getsize(str(tmp_path / "big.ids.json")) < 10_000    # ranges, not 5M numbers
    again = PicoVectorDB(embedding_dim=dim, storage_file=str(tmp_path / "big"), bf16_mirror=True)
    assert len(again) == n - 2
    assert [r[0][K_ID] for r in again.query(qs[1:], top_k=1)] == list(keep.keys())[1:]
    again.close()
    db.close()
