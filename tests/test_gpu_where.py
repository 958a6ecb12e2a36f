"""pvdb_search_where: dict ``where`` filters evaluated on the device from a code column
(picovdb/pico_vdb.py:615-638 followed by :683-714) must select exactly the rows the oracle selects
with the equivalent boolean prefilter, on the scan path and the tensor-core batch path.  Needs a B200."""
import numpy as np
import pytest

from oracle import picovdb_oracle as O

pytestmark = pytest.mark.gpu

F32_RTOL, F32_ATOL = 1e-5, 2e-6


@pytest.fixture
def store():
    from picovdb_b200.engine import DeviceStore

    made = []

    def make(dim, **kw):
        s = DeviceStore(dim, **kw)
        made.append(s)
        return s

    yield make
    for s in made:
        s.close()


def _oracle(vecs, q, k, active, pf):
    qn = O.prepare_queries(q, vecs.shape[1])[0]
    return O.search(vecs, qn, k, active, pf)


@pytest.mark.parametrize("n_wanted", [1, 3, 8, 9, 40])
def test_scan_where_matches_oracle(store, n_wanted):
    n, dim, k = 20_011, 96, 10
    rng = np.random.default_rng(n_wanted)
    s = store(dim)
    s.upsert_range(rng.standard_normal((n, dim)).astype(np.float32), 0)
    vecs = s.download()
    codes = rng.integers(-1, 200, n).astype(np.int32)          # -1 = key absent
    s.column_write(0, codes)
    dead = rng.choice(n, n // 5, replace=False)
    s.delete_rows(dead)
    active = np.ones(n, bool)
    active[dead] = False
    wanted = rng.choice(200, n_wanted, replace=False).tolist()
    pf = np.isin(codes, wanted)
    q = rng.standard_normal((1, dim)).astype(np.float32)
    sc, rows, cand = s.search_where(q, k, 0, wanted, precision="f32")
    assert cand == int((pf & active).sum())
    ref_s, ref_r = _oracle(vecs, q, k, active, pf)
    np.testing.assert_array_equal(rows, ref_r)
    np.testing.assert_allclose(sc, ref_s, rtol=F32_RTOL, atol=F32_ATOL)
    # with an extra row mask (the ids= argument of query())
    extra = rng.random(n) < 0.5
    sc, rows, cand = s.search_where(q, k, 0, wanted, extra=extra, precision="f32")
    assert cand == int((pf & active & extra).sum())
    ref_s, ref_r = _oracle(vecs, q, k, active, pf & extra)
    np.testing.assert_array_equal(rows, ref_r)
    np.testing.assert_allclose(sc, ref_s, rtol=F32_RTOL, atol=F32_ATOL)


def test_where_edge_cases(store):
    n, dim = 1000, 16
    rng = np.random.default_rng(0)
    s = store(dim)
    s.upsert_range(rng.standard_normal((n, dim)).astype(np.float32), 0)
    vecs = s.download()
    codes = (np.arange(n) % 7).astype(np.int32)
    s.column_write(2, codes[:600])                            # rows >= 600 never written: absent
    q = rng.standard_normal((1, dim)).astype(np.float32)
    sc, rows, cand = s.search_where(q, 5, 2, [3])
    pf = (codes == 3) & (np.arange(n) < 600)
    assert cand == int(pf.sum())
    np.testing.assert_array_equal(rows, _oracle(vecs, q, 5, None, pf)[1])
    # scattered update of some rows, including one past the written range
    upd = np.array([1, 2, 999, 650], dtype=np.int64)
    s.column_write(2, np.array([3, -1, 3, 3], np.int32), rows=upd)
    pf[[1, 999, 650]] = True
    pf[2] = False
    sc, rows, cand = s.search_where(q, 5, 2, [3])
    assert cand == int(pf.sum())
    np.testing.assert_array_equal(rows, _oracle(vecs, q, 5, None, pf)[1])
    # no code matches: zero candidates, padded result
    sc, rows, cand = s.search_where(q, 5, 2, [99])
    assert cand == 0 and (rows == -1).all() and np.isneginf(sc).all()
    # fewer candidates than k
    s.column_write(3, np.where(np.arange(n) < 3, 1, 0).astype(np.int32))
    sc, rows, cand = s.search_where(q, 5, 3, [1])
    assert cand == 3 and (rows[0, :3] >= 0).all() and (rows[0, 3:] == -1).all()
    # a column that was never written is an error, as is a bad column number
    with pytest.raises(Exception):
        s.search_where(q, 5, 7, [1])
    with pytest.raises(Exception):
        s.column_write(16, codes)
    # rows grow after the column was written: new rows are absent until written
    s.upsert_range(rng.standard_normal((500, dim)).astype(np.float32), n)
    sc, rows, cand = s.search_where(q, 5, 3, [1])
    assert cand == 3
    s.column_drop(3)
    with pytest.raises(Exception):
        s.search_where(q, 5, 3, [1])


def test_batch_where_matches_oracle(store):
    n, dim, nq, k = 30_000, 128, 70, 10
    rng = np.random.default_rng(11)
    s = store(dim)
    s.upsert_range(rng.standard_normal((n, dim)).astype(np.float32), 0)
    vecs = s.download()
    codes = rng.integers(0, 12, n).astype(np.int32)
    s.column_write(1, codes)
    dead = rng.choice(n, 3000, replace=False)
    s.delete_rows(dead)
    active = np.ones(n, bool)
    active[dead] = False
    q = rng.standard_normal((nq, dim)).astype(np.float32)
    sc, rows, cand = s.search_where(q, k, 1, [2, 5], precision="tf32")
    pf = np.isin(codes, [2, 5])
    assert cand == int((pf & active).sum())
    ref_s, ref_r = _oracle(vecs, q, k, active, pf)
    assert O.recall_at_k(rows, ref_r) >= 0.999
    same = rows == ref_r
    np.testing.assert_allclose(sc[same], ref_s[same], rtol=F32_RTOL, atol=F32_ATOL)
    assert active[rows].all() and pf[rows].all()
