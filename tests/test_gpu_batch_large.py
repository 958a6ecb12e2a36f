"""Batched tensor-core path at sizes where its two-pass (sample + main) schedule and the
multi-launch split (> 4096 queries) are active.  Needs a B200."""
import numpy as np
import pytest

from oracle import picovdb_oracle as O

pytestmark = pytest.mark.gpu

F32_RTOL, F32_ATOL = 1e-5, 2e-6


def _gauss(n, dim, seed):
    return np.random.default_rng(seed).standard_normal((n, dim)).astype(np.float32)


@pytest.fixture
def store_factory():
    from picovdb_b200.engine import DeviceStore

    made = []

    def make(dim, **kw):
        s = DeviceStore(dim, **kw)
        made.append(s)
        return s

    yield make
    for s in made:
        s.close()


def _check(sc, rows, store, qn, k, ref_s, ref_r, min_same):
    assert np.all(np.diff(sc, axis=1) <= 0)
    for qi in range(0, qn.shape[0], 97):  # spot-check: each score is its own row's exact dot product
        exact = store[rows[qi]] @ qn[qi]
        np.testing.assert_allclose(sc[qi], exact, rtol=F32_RTOL, atol=F32_ATOL)
    same = rows == ref_r
    assert same.mean() >= min_same, same.mean()
    np.testing.assert_allclose(sc[same], ref_s[same], rtol=F32_RTOL, atol=F32_ATOL)
    assert O.recall_at_k(rows, ref_r) >= min_same


@pytest.mark.parametrize("k", [10, 100])
def test_two_pass_batch_matches_oracle(store_factory, k):
    # 300k x 64, 2048 queries: 1172 tiles x 16 query tiles -> the sample pass (1/16 of the tiles) runs
    dim, n, nq = 64, 300_000, 2048
    s = store_factory(dim)
    s.upsert_range(_gauss(n, dim, 5), 0)
    dead = np.random.default_rng(6).choice(n, n // 5, replace=False)
    s.delete_rows(dead)
    store = s.download()
    active = np.ones(n, bool)
    active[dead] = False
    qn, _ = O.prepare_queries(_gauss(nq, dim, 7), dim)
    ref_s, ref_r = O.search_chunked(store, qn, k, active)
    sc, rows = s.search(qn, k, precision="tf32", normalized=True)
    _check(sc, rows, store, qn, k, ref_s, ref_r, 0.998)
    assert active[rows].all()


def test_more_than_4096_queries_and_prefilter(store_factory):
    dim, n, nq, k = 32, 120_000, 5000, 10
    s = store_factory(dim, bf16_mirror=True)
    s.upsert_range(_gauss(n, dim, 15), 0)
    store = s.download()
    qn, _ = O.prepare_queries(_gauss(nq, dim, 17), dim)
    pf = (np.arange(n) % 4) != 1
    ref_s, ref_r = O.search_chunked(store, qn, k, None, pf)
    for prec in ("tf32", "bf16"):
        sc, rows = s.search(qn, k, prefilter=pf, precision=prec, normalized=True)
        _check(sc, rows, store, qn, k, ref_s, ref_r, 0.995)
        assert pf[rows].all()


def test_batch_results_do_not_depend_on_timing(store_factory):
    """The candidate pools fill in a timing-dependent order, the result must not: many repeats of the
    same search return exactly the oracle's rows every time (this caught a finalize race that
    garbled one query in a few percent of the runs)."""
    dim, n, nq, k = 32, 120_000, 5000, 10
    s = store_factory(dim, bf16_mirror=True)
    s.upsert_range(_gauss(n, dim, 15), 0)
    store = s.download()
    qn, _ = O.prepare_queries(_gauss(nq, dim, 17), dim)
    pf = (np.arange(n) % 4) != 1
    ref_s, ref_r = O.search_chunked(store, qn, k, None, pf)
    for prec, repeats in (("tf32", 10), ("bf16", 20)):
        first = None
        for _ in range(repeats):
            sc, rows = s.search(qn, k, prefilter=pf, precision=prec, normalized=True)
            if first is None:
                first = (sc.copy(), rows.copy())
                assert (rows == ref_r).mean() >= 0.9995          # near-ties of the low-precision pass aside
                same = rows == ref_r
                np.testing.assert_allclose(sc[same], ref_s[same], rtol=F32_RTOL, atol=F32_ATOL)
            else:
                np.testing.assert_array_equal(rows, first[1])
                np.testing.assert_array_equal(sc, first[0])


@pytest.mark.parametrize("env", [
    {"PVDB_BATCH_TILE_BLOCK": "1"},
    {"PVDB_BATCH_TILE_BLOCK": "3"},
    {"PVDB_BATCH_TILE_BLOCK": "8"},
    {"PVDB_BATCH_TILE_BLOCK": "5", "PVDB_BATCH_NO_CLUSTER": "1"},
    {"PVDB_BATCH_TILE_BLOCK": "4", "PVDB_BATCH_PAIR": "1"},
    {"PVDB_BATCH_TILE_BLOCK": "7", "PVDB_BATCH_CLUSTER": "4"},
])
def test_work_item_schedules_give_the_same_answer(store_factory, monkeypatch, env):
    """Visits are grouped into work items of R consecutive database tiles per query tile (VisitSeq in
    csrc/batch.cu); R, the cluster size and the pair-MMA variant only change WHO scores WHICH tile WHEN.
    Every schedule must return the oracle's rows -- tile counts that are not multiples of R, an odd number
    of query tiles (padding tile in a cluster) and a prefilter included."""
    for key, val in env.items():
        monkeypatch.setenv(key, val)
    dim, n, nq, k = 48, 150_001, 1100, 10   # 586 tiles (last one partial), 9 query tiles
    s = store_factory(dim, bf16_mirror=True)
    s.upsert_range(_gauss(n, dim, 25), 0)
    store = s.download()
    qn, _ = O.prepare_queries(_gauss(nq, dim, 27), dim)
    pf = (np.arange(n) % 5) != 2
    ref_s, ref_r = O.search_chunked(store, qn, k, None, pf)
    for prec in ("tf32", "bf16"):
        sc, rows = s.search(qn, k, prefilter=pf, precision=prec, normalized=True)
        _check(sc, rows, store, qn, k, ref_s, ref_r, 0.995)
        assert pf[rows].all()
