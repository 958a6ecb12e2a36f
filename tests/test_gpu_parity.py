"""Parity of the CUDA path (through the C ABI) with the oracle.  All tests need a B200.

Tolerances (BASELINE.json north_star): fp32 paths -- scores within 1e-5 relative of the
reference, ids identical wherever the score gap exceeds that tolerance; bf16 paths -- 1e-2.
Index outputs of integer work (row ids, padding, active bitmap) are compared exactly.
"""
import os

import numpy as np
import pytest

from oracle import picovdb_oracle as O

pytestmark = pytest.mark.gpu

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")
F32_RTOL = 1e-5
F32_ATOL = 2e-6  # scores near zero: |err| <= eps * sum|q_i v_i| <= ~1e-6 for unit vectors
BF16_RTOL = 1e-2
BF16_ATOL = 4e-3


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


def _gauss(n, dim, seed):
    return np.random.default_rng(seed).standard_normal((n, dim)).astype(np.float32)


# ------------------------------------------------------------------ write side
@pytest.mark.parametrize("dim", [1, 2, 3, 5, 16, 100, 384, 768, 1024, 1536])
def test_upsert_normalises_like_reference(store_factory, dim):
    # reference _normalize (pico_vdb.py:58-68) incl. zero -> e0; fp32 tolerance 1e-6 as in the
    # reference's own tests (tests/test_more.py:258-260)
    raw = _gauss(70, dim, dim) * np.float32(37.5)
    raw[3] = 0.0
    s = store_factory(dim, bf16_mirror=True)
    s.upsert_rows(raw, np.arange(70))
    got = s.download()
    want = O.normalize_rows(raw)
    np.testing.assert_allclose(got, want, rtol=1e-6, atol=1e-7)
    np.testing.assert_array_equal(got[3], want[3])
    assert got.dtype == np.float32 and got.flags["C_CONTIGUOUS"]
    info = s.info()
    assert (info.rows, info.active, info.dim) == (70, 70, dim)
    assert info.ld_f32 % 4 == 0 and info.ld_bf16 % 8 == 0
    np.testing.assert_array_equal(s.fetch_rows([5, 3, 69]), got[[5, 3, 69]])


def test_golden_normalize_vectors(store_factory):
    g = np.load(os.path.join(GOLDEN, "normalize.npz"))
    for key in [k for k in g.files if k.startswith("in_")]:
        dim = int(key[3:]) if key != "in_34" else 2
        s = store_factory(dim)
        s.upsert_range(g[key], 0)
        np.testing.assert_allclose(s.download(), g["out_" + key[3:]], rtol=1e-6, atol=1e-7)


def test_scatter_delete_bitmap_and_growth(store_factory):
    dim = 24
    s = store_factory(dim)
    raw = _gauss(5000, dim, 1)
    rows = np.random.default_rng(2).permutation(9000)[:5000]
    s.upsert_rows(raw, rows)  # scattered rows, forces growth, leaves holes
    info = s.info()
    assert info.rows == rows.max() + 1 and info.active == 5000 and info.capacity >= info.rows
    want = np.zeros((info.rows, dim), np.float32)
    want[rows] = O.normalize_rows(raw)
    np.testing.assert_allclose(s.download(), want, rtol=1e-6, atol=1e-7)
    mask = np.zeros(info.rows, bool)
    mask[rows] = True
    np.testing.assert_array_equal(s.active_mask(), mask)
    dead = rows[:1234]
    s.delete_rows(dead)
    mask[dead] = False
    want[dead] = 0
    np.testing.assert_array_equal(s.active_mask(), mask)
    np.testing.assert_array_equal(s.download()[dead], 0)
    assert s.info().active == 5000 - 1234
    # overwrite some rows and re-activate a deleted one
    s.upsert_rows(raw[:10], dead[:10])
    want[dead[:10]] = O.normalize_rows(raw[:10])
    np.testing.assert_allclose(s.download(), want, rtol=1e-6, atol=1e-7)
    assert s.info().active == 5000 - 1234 + 10


def test_upload_download_compact_roundtrip(store_factory):
    dim = 10
    v = O.normalize_rows_fast(_gauss(333, dim, 3))
    active = np.random.default_rng(4).random(333) > 0.3
    v[~active] = 0
    s = store_factory(dim, bf16_mirror=True)
    s.upload(v, 0, active)
    np.testing.assert_array_equal(s.download(), v)  # raw load: bit exact
    np.testing.assert_array_equal(s.active_mask(), active)
    keep = np.flatnonzero(active)
    s.compact(keep)
    assert s.info().rows == keep.size == s.info().active
    np.testing.assert_array_equal(s.download(), v[keep])
    qn, _ = O.prepare_queries(_gauss(2, dim, 5), dim)
    sc, rows = s.search(qn, 5)
    ref_s, ref_r = O.search(v[keep], qn, 5)
    O.compare_topk(sc, rows, ref_s, ref_r, rtol=F32_RTOL, atol=F32_ATOL)
    # the mirror was compacted too
    sc16, rows16 = s.search(qn, 5, precision="bf16")
    O.compare_topk(sc16, rows16, ref_s, ref_r, rtol=BF16_RTOL, atol=BF16_ATOL)


def test_fixed_capacity_error(store_factory):
    from picovdb_b200 import _native as N

    s = store_factory(4, reserve_rows=1024, fixed_capacity=True)
    s.upsert_range(_gauss(1024, 4, 0), 0)
    with pytest.raises(N.NativeError) as ei:
        s.upsert_range(_gauss(1, 4, 0), 1024)
    assert ei.value.code == N.PVDB_ERR_CAPACITY and "capacity exceeded" in ei.value.message


# ------------------------------------------------------------------ single-query scan
@pytest.mark.parametrize("dim", [3, 16, 20, 48, 100, 384, 768, 1024, 2052])
@pytest.mark.parametrize("k", [1, 10, 100])
def test_scan_matches_oracle(store_factory, dim, k):
    n = 3001
    raw = _gauss(n, dim, 100 + dim)
    s = store_factory(dim, bf16_mirror=True)
    s.upsert_range(raw, 0)
    store = s.download()
    queries = _gauss(3, dim, 7)
    queries[1] = 0.0  # zero query -> e0 (pico_vdb.py:585-590)
    qn, _ = O.prepare_queries(queries, dim)
    ref_s, ref_r = O.search(store, qn, k)
    sc, rows = s.search(queries, k, precision="f32")
    st = O.compare_topk(sc, rows, ref_s, ref_r, rtol=F32_RTOL, atol=F32_ATOL)
    assert st["recall"] >= 0.999
    # pre-normalised flag gives the same answer
    sc2, rows2 = s.search(qn, k, precision="f32", normalized=True)
    O.compare_topk(sc2, rows2, ref_s, ref_r, rtol=F32_RTOL, atol=F32_ATOL)
    # bf16 mirror scan: looser tolerance, recall reported against the fp32 reference
    sc16, rows16 = s.search(queries, k, precision="bf16")
    fin = np.isfinite(ref_s)
    assert np.array_equal(np.isfinite(sc16), fin)
    if dim >= 48 and k >= 10:
        assert O.recall_at_k(rows16, ref_r) >= 0.9
    # each returned score must be the bf16-rounded dot product of its own row within tolerance
    for qi in range(3):
        exact = store[rows16[qi]] @ qn[qi]
        np.testing.assert_allclose(sc16[qi], exact, rtol=BF16_RTOL, atol=BF16_ATOL)


def test_scan_masks_prefilter_and_padding(store_factory):
    dim, n, k = 64, 5000, 10
    raw = _gauss(n, dim, 11)
    s = store_factory(dim)
    s.upsert_range(raw, 0)
    dead = np.random.default_rng(1).choice(n, size=int(0.3 * n), replace=False)
    s.delete_rows(dead)
    store = s.download()
    active = np.ones(n, bool)
    active[dead] = False
    qn, _ = O.prepare_queries(_gauss(4, dim, 12), dim)
    cat = np.arange(n) % 10
    for pf in (None, cat == 0, cat % 2 == 0, np.zeros(n, bool), np.arange(n) == 4321):
        ref_s, ref_r = O.search(store, qn, k, active, pf)
        sc, rows = s.search(qn, k, prefilter=pf, precision="f32", normalized=True)
        O.compare_topk(sc, rows, ref_s, ref_r, rtol=F32_RTOL, atol=F32_ATOL)
        live = rows[rows >= 0]
        assert active[live].all() and (pf is None or np.asarray(pf)[live].all())
    # fewer candidates than k: -inf / -1 padding, exact
    few = np.zeros(n, bool)
    few[np.flatnonzero(active)[:3]] = True
    sc, rows = s.search(qn, k, prefilter=few, precision="f32", normalized=True)
    assert (rows[:, 3:] == -1).all() and np.isneginf(sc[:, 3:]).all() and (rows[:, :3] >= 0).all()


@pytest.mark.parametrize("k", [129, 300, 1000])
def test_scan_large_k_paging(store_factory, k):
    dim, n = 32, 1500
    s = store_factory(dim)
    s.upsert_range(_gauss(n, dim, 21), 0)
    store = s.download()
    qn, _ = O.prepare_queries(_gauss(2, dim, 22), dim)
    ref_s, ref_r = O.search(store, qn, k)
    sc, rows = s.search(qn, k, precision="f32", normalized=True)
    O.compare_topk(sc, rows, ref_s, ref_r, rtol=F32_RTOL, atol=F32_ATOL)
    for qi in range(2):
        assert len(set(rows[qi].tolist())) == k  # no row returned twice across pages


def test_ties_resolve_to_lowest_row(store_factory):
    dim = 8
    v = np.zeros((200, dim), np.float32)
    v[:, 0] = 1.0  # all rows identical -> all scores equal
    s = store_factory(dim)
    s.upsert_range(v, 0)
    sc, rows = s.search(v[:1], 7)
    assert rows[0].tolist() == list(range(7)) and np.allclose(sc, 1.0)


def test_empty_store_and_row_base(store_factory):
    s = store_factory(4)
    sc, rows = s.search(np.ones((2, 4), np.float32), 3)
    assert (rows == -1).all() and np.isneginf(sc).all()
    s.upsert_range(np.eye(4, dtype=np.float32), 0)
    s.set_row_base(1000)
    sc, rows = s.search(np.eye(4, dtype=np.float32)[2:3], 2)
    assert rows[0, 0] == 1002 and sc[0, 0] == pytest.approx(1.0)


# ------------------------------------------------------------------ golden fixtures via the C ABI
@pytest.mark.parametrize(
    "name,k", [("gauss_n600_d48.npz", 10), ("gauss_n400_d384_del30.npz", 10), ("gauss_n900_d20_k100.npz", 100)]
)
def test_golden_fixtures_cabi(store_factory, name, k):
    g = np.load(os.path.join(GOLDEN, name))
    n, dim = g["raw"].shape
    s = store_factory(dim)
    s.upsert_range(g["raw"], 0)
    dead = np.flatnonzero(g["deleted"])
    if dead.size:
        s.delete_rows(dead)
    np.testing.assert_allclose(s.download(), g["store"], rtol=1e-6, atol=1e-7)
    cat = np.arange(n) % 10
    subset = np.zeros(n, bool)
    subset[::7] = True
    for key, pf in (("", None), ("where_eq", cat == 0), ("where_in", np.isin(cat, [1, 2, 3])),
                    ("where_fn", cat % 2 == 0), ("subset", subset)):
        sc, rows = s.search(g["queries"], k, prefilter=pf, precision="f32")
        ref_r = g[f"ids_{key}" if key else "ids"]
        ref_s = g[f"scores_{key}" if key else "scores"]
        O.compare_topk(sc, rows, ref_s, ref_r, rtol=F32_RTOL, atol=F32_ATOL)


# ------------------------------------------------------------------ size-independent properties
def test_properties_at_scale(store_factory):
    """1M x 128 (0.5 GB): self-query returns self with score 1; results are sorted; idempotent;
    deleting the winner promotes the runner-up; agrees with the chunked oracle."""
    import torch

    dim, n, k = 128, 1_000_000, 10
    s = store_factory(dim, reserve_rows=n)
    gen = torch.Generator(device="cuda").manual_seed(123)
    chunk = 250_000
    for r0 in range(0, n, chunk):
        x = torch.randn(chunk, dim, device="cuda", generator=gen)
        s.upsert_range_dev(x.data_ptr(), r0, chunk)
    torch.cuda.synchronize()
    probe = np.array([0, 123_456, 999_999])
    qv = s.fetch_rows(probe)
    sc, rows = s.search(qv, k, precision="f32")
    assert rows[:, 0].tolist() == probe.tolist()
    np.testing.assert_allclose(sc[:, 0], 1.0, rtol=1e-5)
    assert np.all(np.diff(sc, axis=1) <= 0)
    sc_b, rows_b = s.search(qv, k, precision="f32")
    np.testing.assert_array_equal(rows, rows_b)
    np.testing.assert_array_equal(sc, sc_b)
    store = s.download()
    ref_s, ref_r = O.search_chunked(store, O.prepare_queries(qv, dim)[0], k)
    O.compare_topk(sc, rows, ref_s, ref_r, rtol=F32_RTOL, atol=F32_ATOL)
    s.delete_rows(probe)
    sc2, rows2 = s.search(qv, k, precision="f32")
    np.testing.assert_array_equal(rows2[:, : k - 1], rows[:, 1:])


def test_merge_topk_kernel(store_factory):
    import torch
    from picovdb_b200.engine import merge_topk_dev

    rng = np.random.default_rng(5)
    nl, nq, k = 8, 37, 10
    parts_s = np.sort(rng.standard_normal((nl, nq, k)).astype(np.float32), axis=2)[:, :, ::-1].copy()
    parts_r = rng.permutation(nl * nq * k).reshape(nl, nq, k).astype(np.int64)
    parts_s[3, :, 6:] = -np.inf  # a short shard
    parts_r[3, :, 6:] = -1
    ds = torch.from_numpy(parts_s).cuda()
    dr = torch.from_numpy(parts_r).cuda()
    out_s = torch.empty(nq, k, dtype=torch.float32, device="cuda")
    out_r = torch.empty(nq, k, dtype=torch.int64, device="cuda")
    merge_topk_dev(0, ds.data_ptr(), dr.data_ptr(), nl, nq, k, out_s.data_ptr(), out_r.data_ptr(),
                   torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    ref_s, ref_r = O.merge_topk(list(parts_s), list(parts_r), k)
    np.testing.assert_array_equal(out_r.cpu().numpy(), ref_r)
    np.testing.assert_array_equal(out_s.cpu().numpy(), ref_s)


# ------------------------------------------------------------------ batched tensor-core path
def _check_batch(sc, rows, store, qn, k, active, pf, ref_r, score_rtol, score_atol, min_recall):
    """Returned scores must be the dot product of their own row (within tolerance), lists sorted,
    rows unique / active / inside the prefilter, and recall against the fp32 reference high."""
    nq = qn.shape[0]
    assert sc.shape == (nq, k) and rows.shape == (nq, k)
    for qi in range(nq):
        live = rows[qi] >= 0
        n_live = int(live.sum())
        assert n_live == int((ref_r[qi] >= 0).sum()), f"query {qi}: result count differs"
        assert live[:n_live].all() and np.isneginf(sc[qi, n_live:]).all()
        r = rows[qi, :n_live]
        assert len(set(r.tolist())) == n_live
        if active is not None:
            assert active[r].all()
        if pf is not None:
            assert np.asarray(pf)[r].all()
        exact = store[r] @ qn[qi]
        np.testing.assert_allclose(sc[qi, :n_live], exact, rtol=score_rtol, atol=score_atol)
        assert np.all(np.diff(sc[qi, :n_live]) <= 0)
    rec = O.recall_at_k(rows, ref_r)
    assert rec >= min_recall, f"recall {rec}"
    return rec


@pytest.mark.parametrize("dim", [16, 100, 384, 768, 1024])
@pytest.mark.parametrize("nq,k", [(5, 10), (64, 1), (130, 10), (300, 100)])
def test_batch_tf32_rescored_matches_oracle(store_factory, dim, nq, k):
    n = 5003  # not a multiple of the 256-row tile
    s = store_factory(dim)
    s.upsert_range(_gauss(n, dim, 200 + dim), 0)
    store = s.download()
    queries = _gauss(nq, dim, 31)
    queries[2] = 0.0
    qn, _ = O.prepare_queries(queries, dim)
    ref_s, ref_r = O.search(store, qn, k)
    sc, rows = s.search(queries, k, precision="tf32")
    _check_batch(sc, rows, store, qn, k, None, None, ref_r, F32_RTOL, F32_ATOL, 0.995)
    # auto precision picks the same path for a batch this large
    sc2, rows2 = s.search(queries, k)
    np.testing.assert_array_equal(rows2, rows)
    np.testing.assert_array_equal(sc2, sc)


def test_batch_masks_and_prefilter(store_factory):
    dim, n, k, nq = 64, 7001, 10, 40
    s = store_factory(dim, bf16_mirror=True)
    s.upsert_range(_gauss(n, dim, 41), 0)
    dead = np.random.default_rng(1).choice(n, size=int(0.3 * n), replace=False)
    s.delete_rows(dead)
    store = s.download()
    active = np.ones(n, bool)
    active[dead] = False
    qn, _ = O.prepare_queries(_gauss(nq, dim, 42), dim)
    cat = np.arange(n) % 10
    few = np.zeros(n, bool)
    few[np.flatnonzero(active)[:4]] = True
    for pf in (None, cat == 0, cat % 2 == 0, np.zeros(n, bool), few):
        ref_s, ref_r = O.search(store, qn, k, active, pf)
        for prec, rtol, atol, rec in (("tf32", F32_RTOL, F32_ATOL, 0.995), ("bf16", F32_RTOL, F32_ATOL, 0.98)):
            sc, rows = s.search(qn, k, prefilter=pf, precision=prec, normalized=True)
            _check_batch(sc, rows, store, qn, k, active, pf, ref_r, rtol, atol, rec)


def test_batch_without_rescoring_reports_tensor_core_scores(store_factory):
    dim, n, k, nq = 384, 4000, 10, 33
    s = store_factory(dim, bf16_mirror=True)
    s.upsert_range(_gauss(n, dim, 51), 0)
    store = s.download()
    qn, _ = O.prepare_queries(_gauss(nq, dim, 52), dim)
    ref_s, ref_r = O.search(store, qn, k)
    sc, rows = s.search(qn, k, precision="tf32", normalized=True, rescore=False)
    _check_batch(sc, rows, store, qn, k, None, None, ref_r, 2e-3, 2e-3, 0.9)   # tf32 inputs: ~5e-4 relative
    sc, rows = s.search(qn, k, precision="bf16", normalized=True, rescore=False)
    _check_batch(sc, rows, store, qn, k, None, None, ref_r, BF16_RTOL, BF16_ATOL, 0.9)


def test_batch_bf16_only_store(store_factory):
    dim, n, k, nq = 384, 3000, 10, 20
    s = store_factory(dim, keep_f32=False, bf16_mirror=True)
    raw = _gauss(n, dim, 61)
    s.upsert_range(raw, 0)
    ref_store = O.normalize_rows(raw)
    np.testing.assert_allclose(s.download(), ref_store, rtol=8e-3, atol=1e-3)  # bf16 storage
    qn, _ = O.prepare_queries(_gauss(nq, dim, 62), dim)
    ref_s, ref_r = O.search(ref_store, qn, k)
    sc, rows = s.search(qn, k, normalized=True)
    _check_batch(sc, rows, ref_store, qn, k, None, None, ref_r, BF16_RTOL, BF16_ATOL, 0.9)
    sc1, rows1 = s.search(qn[:1], k, normalized=True)  # single query -> bf16 scan
    _check_batch(sc1, rows1, ref_store, qn[:1], k, None, None, ref_r[:1], BF16_RTOL, BF16_ATOL, 0.9)


def test_batch_large_matches_chunked_oracle(store_factory):
    """C1-like shape scaled to what the oracle finishes in seconds: 100k x 256, 1000 queries."""
    dim, n, k, nq = 256, 100_000, 10, 1000
    s = store_factory(dim)
    s.upsert_range(_gauss(n, dim, 71), 0)
    store = s.download()
    qn, _ = O.prepare_queries(_gauss(nq, dim, 72), dim)
    ref_s, ref_r = O.search_chunked(store, qn, k)
    sc, rows = s.search(qn, k, normalized=True)
    rec = _check_batch(sc, rows, store, qn, k, None, None, ref_r, F32_RTOL, F32_ATOL, 0.999)
    # where the reference ids agree exactly the scores must agree at 1e-5 as well
    same = rows == ref_r
    np.testing.assert_allclose(sc[same], ref_s[same], rtol=F32_RTOL, atol=F32_ATOL)
    assert same.mean() >= 0.999, (rec, same.mean())


def test_batch_through_picovectordb(tmp_path):
    from picovdb_b200 import PicoVectorDB, K_ID

    dim, n = 32, 2000
    db = PicoVectorDB(embedding_dim=dim, storage_file=str(tmp_path / "b"))
    v = _gauss(n, dim, 81)
    db.upsert_array(v)
    qs = _gauss(50, dim, 82)
    res = db.query(qs, top_k=5)
    qn, _ = O.prepare_queries(qs, dim)
    ref_s, ref_r = O.search(db._vectors, qn, 5)
    got = np.array([[r[K_ID] for r in rows] for rows in res])
    assert (got == ref_r).mean() >= 0.995
    db.close()


@pytest.mark.parametrize("precision,mirror", [("tf32", False), ("bf16", True)])
def test_batch_duplicate_rows_tie_order(store_factory, precision, mirror):
    """Every vector is stored three times (rows r, r + n, r + 2n): scores tie exactly, so the result
    must list the copies by ascending row and never lose the lowest copy at a selection boundary --
    the two accumulator halves and the CTAs share thresholds, which must stay non-strict for this."""
    n, dim, nq, k = 6000, 64, 200, 9
    base = _gauss(n, dim, 77)
    s = store_factory(dim, bf16_mirror=mirror)
    s.upsert_range(np.concatenate([base, base, base]), 0)
    q = _gauss(nq, dim, 78)
    q[:5] = base[:5]
    sc, rows = s.search(q, k, precision=precision)
    store = s.download()
    qn = O.prepare_queries(q, dim)[0]
    exact = (store[:n] @ qn.T).T                                   # (nq, n) fp32 scores of the distinct vectors
    best = np.argsort(-exact, axis=1, kind="stable")[:, : k // 3]
    want = (best[:, :, None] + np.arange(3)[None, None, :] * n).reshape(nq, k)
    if precision == "tf32":
        np.testing.assert_array_equal(rows, want)
    else:  # bf16 candidates, fp32 re-scored: near-ties between distinct vectors may swap
        assert (rows == want).mean() > 0.97
    trip = rows.reshape(nq, k // 3, 3)
    assert (np.diff(trip, axis=2) == n).all()                      # copies in ascending row order
    sc3 = sc.reshape(nq, k // 3, 3)
    assert (sc3.max(axis=2) == sc3.min(axis=2)).all()              # and with identical scores
    assert rows[:5, 0].tolist() == [0, 1, 2, 3, 4]


# ------------------------------------------------------------------ exactness guard of the tensor paths
def _near_duplicate_corpus(n, dim, n_dup, scale, seed):
    """Gaussian rows plus `n_dup` near-copies of one vector: base + scale * noise, normalised."""
    rng = np.random.default_rng(seed)
    rows = rng.standard_normal((n, dim)).astype(np.float32)
    base = rng.standard_normal(dim).astype(np.float32)
    base /= np.linalg.norm(base)
    where = rng.choice(n, n_dup, replace=False)
    noise = rng.standard_normal((n_dup, dim)).astype(np.float32) / np.sqrt(dim)
    rows[where] = base[None, :] + np.float32(scale) * noise
    queries = base[None, :] + 0.7 * rng.standard_normal((12, dim)).astype(np.float32) / np.sqrt(dim)
    return rows, queries.astype(np.float32), where


@pytest.mark.parametrize("prec,mirror", [("tf32", False), ("bf16", True)])
def test_guard_near_duplicates_fall_back_to_exact_scan(store_factory, prec, mirror):
    """500 near-duplicates whose mutual score differences (~1e-3 * noise) sit far below the tensor
    core's input rounding error: ranking them by tf32 / bf16 scores is a lottery, so top-10 + 32 (54)
    candidates miss true top-10 rows.  The guard must notice (the re-scored 10th best is not provably
    above the weakest kept candidate + rounding bound) and re-run those queries on the exact scan:
    ids then equal the fp32 scan's / the oracle's.  Without the guard the same call is wrong."""
    dim, n, k = 128, 20_000, 10
    raw, queries, _ = _near_duplicate_corpus(n, dim, 500, 1e-3, 7)
    s = store_factory(dim, bf16_mirror=mirror)
    s.upsert_range(raw, 0)
    store = s.download()
    qn, _ = O.prepare_queries(queries, dim)
    ref_s, ref_r = O.search(store, qn, k)
    exact_s, exact_r = s.search(queries, k, precision="f32")          # one exact scan per query
    O.compare_topk(exact_s, exact_r, ref_s, ref_r, rtol=F32_RTOL, atol=F32_ATOL)

    got_s, got_r = s.search(queries, k, precision=prec)               # guarded tensor-core batch
    flagged, _ = s.guard_stats()
    assert flagged == len(queries), f"every query sits in the near-duplicate cluster, {flagged} flagged"
    np.testing.assert_array_equal(got_r, exact_r)
    np.testing.assert_array_equal(got_s, exact_s)

    raw_s, raw_r = s.search(queries, k, precision=prec, guard=False)  # what round 1 returned
    assert s.guard_stats()[0] == 0
    assert O.recall_at_k(raw_r, exact_r) < 0.9, "the corpus must actually defeat the fixed slack"


def test_guard_quiet_on_well_separated_data_and_batch_equals_single(store_factory):
    """Gaussian data: nothing is flagged, and a query returns the same ids alone (exact scan) and
    inside a batch (tensor-core pass + re-scoring), on an fp32 store and on a bf16-only store."""
    dim, n, k, nq = 384, 30_000, 10, 150
    raw = _gauss(n, dim, 81)
    queries = _gauss(nq, dim, 82)
    for kw, prec in (({}, "tf32"), ({"keep_f32": False, "bf16_mirror": True}, "bf16")):
        s = store_factory(dim, **kw)
        s.upsert_range(raw, 0)
        b_s, b_r = s.search(queries, k, precision=prec)
        assert s.guard_stats()[0] == 0
        for qi in range(0, nq, 7):
            o_s, o_r = s.search(queries[qi:qi + 1], k)                # single query: exact scan of this store
            np.testing.assert_array_equal(b_r[qi], o_r[0])
            np.testing.assert_allclose(b_s[qi], o_s[0], rtol=2e-6, atol=1e-6)


def test_guard_large_k_slack(store_factory):
    """k = 100 takes the wider slack (k_sel = 160): still exact, nothing flagged on Gaussian rows."""
    dim, n, k, nq = 768, 40_000, 100, 64
    s = store_factory(dim)
    s.upsert_range(_gauss(n, dim, 91), 0)
    store = s.download()
    qn, _ = O.prepare_queries(_gauss(nq, dim, 92), dim)
    ref_s, ref_r = O.search(store, qn, k)
    sc, rows = s.search(qn, k, precision="tf32", normalized=True)
    assert s.guard_stats()[0] == 0
    O.compare_topk(sc, rows, ref_s, ref_r, rtol=F32_RTOL, atol=F32_ATOL)


# ------------------------------------------------------------------ the large-N oracle is itself pinned
def test_torch_oracle_is_pinned_to_the_numpy_oracle():
    """tests/_torch_oracle.py (chunked fp32 torch brute force, the expected value at BASELINE's full
    sizes) against the numpy oracle at 100k rows: same ids, same scores, masks and zero vectors too."""
    import torch

    from _torch_oracle import TorchOracle

    dim, n, k = 96, 100_000, 25
    dev = torch.device("cuda", 0)
    raw = _gauss(n, dim, 101) * np.float32(3.0)
    raw[17] = 0.0
    queries = _gauss(9, dim, 102)
    queries[4] = 0.0
    active = np.ones(n, bool)
    active[np.random.default_rng(3).choice(n, n // 4, replace=False)] = False
    store = O.normalize_rows(raw)
    qn, _ = O.prepare_queries(queries, dim)
    for elig in (None, active):
        orc = TorchOracle(queries, k, dev)
        for r0 in range(0, n, 30_000):                           # ragged chunks
            x = torch.from_numpy(raw[r0:r0 + 30_000]).to(dev)
            e = None if elig is None else torch.from_numpy(elig[r0:r0 + 30_000]).to(dev)
            orc.update(x, r0, e)
        got_s, got_r = orc.result()
        ref_s, ref_r = O.search(store, qn, k, elig)
        O.compare_topk(got_s, got_r, ref_s, ref_r, rtol=2e-6, atol=1e-6)


# ------------------------------------------------------------------ streamed persistence (8(f) row 3)
def test_streamed_download_upload_roundtrip_and_bf16_file(store_factory, tmp_path):
    """Multi-block pinned pipelines (rows chosen so several 32 MiB blocks and a ragged tail occur):
    download == what was uploaded, download(out=memmap) writes the mapped file, and the bf16 bit
    patterns of a bf16-only store round-trip through download_bf16 / upload_bf16 exactly."""
    from numpy.lib.format import open_memmap

    dim, n = 256, 100_003           # 102 MB of fp32 rows: four pipeline blocks
    raw = _gauss(n, dim, 111)
    s = store_factory(dim, bf16_mirror=True)
    s.upsert_range(raw, 0)                                 # streamed H2D + fused normalise
    want = O.normalize_rows(raw)
    got = s.download()
    np.testing.assert_allclose(got, want, rtol=1e-6, atol=1e-7)
    mm = open_memmap(str(tmp_path / "m.npy"), mode="w+", dtype=np.float32, shape=(n, dim))
    s.download(0, n, out=mm)
    mm.flush()
    np.testing.assert_array_equal(np.load(str(tmp_path / "m.npy")), got)
    active = np.ones(n, bool)
    active[::3] = False
    t = store_factory(dim)
    t.upload(got, 0, active)                               # streamed raw load
    np.testing.assert_array_equal(t.download(), got)
    np.testing.assert_array_equal(t.active_mask(), active)
    b = store_factory(dim, keep_f32=False, bf16_mirror=True)
    b.upsert_range(raw, 0)
    bits = b.download_bf16()
    assert bits.dtype == np.uint16 and bits.shape == (n, dim)
    as_f32 = (bits.astype(np.uint32) << 16).view(np.float32)
    np.testing.assert_array_equal(as_f32, b.download())
    np.testing.assert_allclose(as_f32, want, rtol=8e-3, atol=1e-3)
    c = store_factory(dim, keep_f32=False, bf16_mirror=True)
    c.upload_bf16(bits, 0, active)
    np.testing.assert_array_equal(c.download_bf16(), bits)
    np.testing.assert_array_equal(c.active_mask(), active)
    q = _gauss(3, dim, 112)
    np.testing.assert_array_equal(c.search(q, 5, prefilter=None)[1], b.search(q, 5, prefilter=active)[1])
    with pytest.raises(Exception):
        s.upload_bf16(bits[:32], 0)                        # a store with fp32 rows loads fp32


def test_bf16_only_db_saves_and_loads_its_mirror(tmp_path):
    from picovdb_b200 import K_ID, K_VECTOR, PicoVectorDB

    dim, n = 32, 3000
    vecs = _gauss(n, dim, 113)
    path = str(tmp_path / "b16")
    db = PicoVectorDB(embedding_dim=dim, storage_file=path, keep_f32=False, bf16_mirror=True)
    db.upsert([{K_VECTOR: vecs[i], K_ID: f"r{i}"} for i in range(n)])
    db.delete(["r5"])
    want = [r[K_ID] for r in db.query(vecs[7], top_k=5)]
    db.save()
    assert os.path.exists(path + ".vecs.bf16.npy") and not os.path.exists(path + ".vecs.npy")
    assert os.path.getsize(path + ".vecs.bf16.npy") < n * dim * 2 + 1024
    again = PicoVectorDB(embedding_dim=dim, storage_file=path, keep_f32=False, bf16_mirror=True)
    assert [r[K_ID] for r in again.query(vecs[7], top_k=5)] == want and len(again) == n - 1
    np.testing.assert_array_equal(again._vectors, db._vectors)
    again.close()
    db.close()
    ref = PicoVectorDB(embedding_dim=dim, storage_file=path, keep_f32=False, bf16_mirror=True, save_dtype="f32")
    ref.save()                                             # the reference's fp32 .npy on request
    assert os.path.exists(path + ".vecs.npy") and not os.path.exists(path + ".vecs.bf16.npy")
    assert np.load(path + ".vecs.npy").shape == (n, dim)
    ref.close()


# ------------------------------------------------------------------ concurrent readers of one store
def test_concurrent_searches_on_one_store(store_factory):
    """Eight threads search the same handle at once (the store mutex is held only while a call enqueues
    its work; every call waits for its own event and owns its own pinned result slot): every thread
    must get exactly what a sequential call returns, for single queries, filtered queries and batches."""
    import threading

    dim, n, k = 64, 200_000, 10
    s = store_factory(dim, bf16_mirror=True)
    s.upsert_range(_gauss(n, dim, 121), 0)
    queries = _gauss(96, dim, 122)
    pf = (np.arange(n) % 5) != 0
    want_one = [s.search(queries[i:i + 1], k) for i in range(96)]
    want_pf = [s.search(queries[i:i + 1], k, prefilter=pf) for i in range(0, 96, 8)]
    want_batch = s.search(queries, k, precision="tf32")
    errors = []

    def worker(t):
        try:
            for rep in range(3):
                for i in range(t, 96, 8):
                    got = s.search(queries[i:i + 1], k)
                    assert np.array_equal(got[1], want_one[i][1]) and np.array_equal(got[0], want_one[i][0])
                j = t
                got = s.search(queries[8 * j:8 * j + 1], k, prefilter=pf)
                assert np.array_equal(got[1], want_pf[j][1])
                if t % 4 == 0:
                    got = s.search(queries, k, precision="tf32")
                    assert np.array_equal(got[1], want_batch[1]) and np.array_equal(got[0], want_batch[0])
        except Exception as exc:  # noqa: BLE001
            errors.append(repr(exc))

    threads = [threading.Thread(target=worker, args=(t,)) for t in range(8)]
    for th in threads:
        th.start()
    for th in threads:
        th.join()
    assert not errors, errors[:3]


@pytest.mark.parametrize("dim", [8, 24, 100, 128, 200, 384, 500, 512, 520, 768, 1000, 1536])
@pytest.mark.parametrize("k", [1, 10, 100])
def test_bf16_scan_on_mma_matches_the_cuda_core_scan_and_the_exact_dot(store_factory, monkeypatch, dim, k):
    """bf16 rows are scored on mma.sync with the fp32 query split into three bf16 terms (scan_mma_topk_kernel;
    query fragments in registers up to dim 512, in shared memory beyond).  Products are exact and accumulation is fp32, so the scores must agree with the
    CUDA-core kernel (PVDB_SCAN_NO_MMA=1) and with the float64 dot product of the stored bf16 rows to ~1e-6
    -- the fp32 tolerance, not the bf16 one -- with deleted rows, a dense and a sparse (prefilter) walk."""
    n = 5003
    s = store_factory(dim, keep_f32=False, bf16_mirror=True)
    s.upsert_range(_gauss(n, dim, 300 + dim), 0)
    dead = np.random.default_rng(3).choice(n, n // 7, replace=False)
    s.delete_rows(dead)
    rows_f32 = s.download()                      # the stored bf16 values, widened
    active = np.ones(n, bool)
    active[dead] = False
    pf = (np.arange(n) % 3) != 1
    queries = _gauss(4, dim, 11)
    queries[2] = 0.0
    qn, _ = O.prepare_queries(queries, dim)
    for prefilter in (None, pf):
        live = active if prefilter is None else (active & prefilter)
        for qi in range(4):
            exact = rows_f32.astype(np.float64) @ qn[qi].astype(np.float64)
            exact[~live] = -np.inf
            monkeypatch.delenv("PVDB_SCAN_NO_MMA", raising=False)
            sc, rows = s.search(queries[qi:qi + 1], k, prefilter=prefilter)
            monkeypatch.setenv("PVDB_SCAN_NO_MMA", "1")
            sc_c, rows_c = s.search(queries[qi:qi + 1], k, prefilter=prefilter)
            monkeypatch.delenv("PVDB_SCAN_NO_MMA", raising=False)
            assert np.all(np.diff(sc[0]) <= 0)
            assert live[rows[0]].all()
            np.testing.assert_allclose(sc[0], exact[rows[0]], rtol=F32_RTOL, atol=F32_ATOL)
            np.testing.assert_allclose(sc[0], sc_c[0], rtol=F32_RTOL, atol=F32_ATOL)
            # the same rows, except where two scores are closer than the accumulation-order noise
            kth = np.sort(exact)[::-1][k - 1]
            for r in set(rows[0].tolist()) ^ set(rows_c[0].tolist()):
                assert abs(exact[r] - kth) <= 4e-6, (r, exact[r], kth)


def test_bf16_mma_scan_pages_large_k(store_factory):
    """k > 128 runs the scan in pages bounded by the previous page's last key; the mma kernel takes the same
    bound.  Every returned row must be live, scores sorted, and the set equal to the exact top k."""
    dim, n, k = 200, 4001, 300
    s = store_factory(dim, keep_f32=False, bf16_mirror=True)
    s.upsert_range(_gauss(n, dim, 77), 0)
    dead = np.arange(0, n, 9)
    s.delete_rows(dead)
    rows_f64 = s.download().astype(np.float64)
    q = _gauss(1, dim, 78)
    qn, _ = O.prepare_queries(q, dim)
    exact = rows_f64 @ qn[0].astype(np.float64)
    exact[dead] = -np.inf
    sc, rows = s.search(q, k)
    assert np.all(np.diff(sc[0]) <= 0) and len(set(rows[0].tolist())) == k
    np.testing.assert_allclose(sc[0], exact[rows[0]], rtol=F32_RTOL, atol=F32_ATOL)
    want = np.argsort(-exact, kind="stable")[:k]
    kth = exact[want[-1]]
    for r in set(rows[0].tolist()) ^ set(want.tolist()):
        assert abs(exact[r] - kth) <= 4e-6


# ------------------------------------------------------------------ several queries per pass
@pytest.mark.parametrize("kw", [{}, {"keep_f32": False, "bf16_mirror": True}], ids=["f32", "bf16"])
@pytest.mark.parametrize("dim", [8, 20, 100, 384, 520, 1024, 1536])
def test_exact_batches_share_a_pass_and_equal_single_queries(store_factory, monkeypatch, kw, dim):
    """An exact batch (precision f32 / the bf16 rows of a bf16-only store) is scanned several queries per pass
    (4 over fp32 rows, 2 over bf16 rows on mma.sync; scan_kernel.cuh).  Per query the arithmetic is the
    single-query kernel's, so every query must come back with the SAME BITS it gets alone -- raw and
    pre-normalised queries, zero queries, deleted rows, dense and sparse walks, ragged group sizes -- and k > 32
    must quietly take the single-query kernels.  The oracle pins the values."""
    n = 4099
    prec = "f32" if not kw else "bf16"
    s = store_factory(dim, **kw)
    s.upsert_range(_gauss(n, dim, 500 + dim), 0)
    dead = np.random.default_rng(5).choice(n, n // 5, replace=False)
    s.delete_rows(dead)
    active = np.ones(n, bool)
    active[dead] = False
    store = s.download()
    pf = (np.arange(n) % 4) != 2
    from picovdb_b200._native import kernel_launches as launches

    for nq in (2, 3, 4, 5, 9):
        queries = _gauss(nq, dim, 40 + nq)
        queries[nq // 2] = 0.0
        qn, _ = O.prepare_queries(queries, dim)
        for k in (1, 10, 32, 33):
            for prefilter in (None, pf):
                l0 = launches()
                b_s, b_r = s.search(queries, k, prefilter=prefilter, precision=prec, scan_only=True)
                used = launches() - l0
                width = 4 if prec == "f32" else 2
                want = nq if k > 32 else (nq // width + (1 if nq % width >= 2 else 0) + (1 if nq % width == 1 else 0))
                assert used == want, (nq, k, used, want)
                n_s, n_r = s.search(qn, k, prefilter=prefilter, precision=prec, normalized=True, scan_only=True)
                np.testing.assert_array_equal(n_r, b_r)   # (host-normalised queries differ in the last bit)
                np.testing.assert_allclose(n_s, b_s, rtol=2e-6, atol=1e-6)
                for qi in range(nq):
                    o_s, o_r = s.search(queries[qi:qi + 1], k, prefilter=prefilter, precision=prec)
                    np.testing.assert_array_equal(b_r[qi], o_r[0])
                    np.testing.assert_array_equal(b_s[qi], o_s[0])
                if prec == "f32":
                    ref_s, ref_r = O.search(store, qn, k, active, prefilter)
                    O.compare_topk(b_s, b_r, ref_s, ref_r, rtol=F32_RTOL, atol=F32_ATOL)
    # the switch: one launch per query again, same bits
    queries = _gauss(6, dim, 77)
    b_s, b_r = s.search(queries, 10, precision=prec, scan_only=True)
    monkeypatch.setenv("PVDB_SCAN_NO_MULTI", "1")
    l0 = launches()
    o_s, o_r = s.search(queries, 10, precision=prec, scan_only=True)
    assert launches() - l0 == 6
    monkeypatch.delenv("PVDB_SCAN_NO_MULTI", raising=False)
    np.testing.assert_array_equal(b_r, o_r)
    np.testing.assert_array_equal(b_s, o_s)


def test_exact_batches_edge_cases(store_factory):
    """Several queries per pass at the edges: rows too wide for four queries to share shared memory (the call
    quietly takes one pass per query), a store whose rows are all deleted (padding only), fewer live rows than k,
    a store smaller than one warp step, and 33 queries (8 groups of 4 + a lone remainder) against the oracle."""
    from picovdb_b200._native import kernel_launches as launches

    # dim 12000: 4 x 48 KB of queries do not fit next to the lists -> single-query kernels, same answers
    dim, n = 12000, 300
    s = store_factory(dim)
    s.upsert_range(_gauss(n, dim, 1), 0)
    store = s.download()
    q = _gauss(3, dim, 2)
    qn, _ = O.prepare_queries(q, dim)
    l0 = launches()
    sc, rows = s.search(q, 5, precision="f32")
    assert launches() - l0 == 3
    ref_s, ref_r = O.search(store, qn, 5)
    O.compare_topk(sc, rows, ref_s, ref_r, rtol=F32_RTOL, atol=F32_ATOL)

    # the widest rows that still share passes (147 KB / 90 KB of dynamic shared memory per block)
    for dim, kw, prec, width in ((8192, {}, "f32", 4), (4096, {"keep_f32": False, "bf16_mirror": True}, "bf16", 2)):
        s = store_factory(dim, **kw)
        s.upsert_range(_gauss(200, dim, 6), 0)
        q = _gauss(width + 1, dim, 7)
        l0 = launches()
        sc, rows = s.search(q, 10, precision=prec, scan_only=True)
        assert launches() - l0 == 2                      # one shared pass + the lone remainder
        for i in range(width + 1):
            o_s, o_r = s.search(q[i:i + 1], 10, precision=prec)
            np.testing.assert_array_equal(rows[i], o_r[0])
            np.testing.assert_array_equal(sc[i], o_s[0])

    for kw, prec in (({}, "f32"), ({"keep_f32": False, "bf16_mirror": True}, "bf16")):
        dim, n, k = 72, 700, 10
        s = store_factory(dim, **kw)
        s.upsert_range(_gauss(n, dim, 3), 0)
        q = _gauss(33, dim, 4)
        single = [s.search(q[i:i + 1], k, precision=prec) for i in range(33)]
        sc, rows = s.search(q, k, precision=prec, scan_only=True)
        np.testing.assert_array_equal(rows, np.concatenate([r for _, r in single]))
        np.testing.assert_array_equal(sc, np.concatenate([x for x, _ in single]))
        if prec == "f32":
            qn, _ = O.prepare_queries(q, dim)
            ref_s, ref_r = O.search(s.download(), qn, k)
            O.compare_topk(sc, rows, ref_s, ref_r, rtol=F32_RTOL, atol=F32_ATOL)
        # fewer live rows than k, then none at all
        s.delete_rows(np.arange(4, n))
        sc, rows = s.search(q[:5], k, precision=prec, scan_only=True)
        assert (rows[:, :4] >= 0).all() and (rows[:, :4] < 4).all() and (rows[:, 4:] == -1).all()
        assert np.isneginf(sc[:, 4:]).all() and np.all(np.diff(sc[:, :4], axis=1) <= 0)
        s.delete_rows(np.arange(0, 4))
        sc, rows = s.search(q[:5], k, precision=prec, scan_only=True)
        assert (rows == -1).all() and np.isneginf(sc).all()
        # a store smaller than one warp step
        t = store_factory(dim, **kw)
        t.upsert_range(_gauss(3, dim, 5), 0)
        sc, rows = t.search(q[:4], 2, precision=prec, scan_only=True)
        for i in range(4):
            o_s, o_r = t.search(q[i:i + 1], 2, precision=prec)
            np.testing.assert_array_equal(rows[i], o_r[0])
            np.testing.assert_array_equal(sc[i], o_s[0])
