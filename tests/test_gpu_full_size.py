"""Parity at BASELINE.json's full sizes.  The numpy oracle cannot hold these matrices, so the expected
values come from an INDEPENDENT chunked fp32 brute force in plain torch (tests/_torch_oracle.py, pinned
to the numpy oracle at oracle-sized inputs by tests/test_gpu_parity.py) that scores the raw rows while
they are generated -- plus size-independent properties: a stored row queried against the store returns
itself with score 1, lists are sorted, results are idempotent.  Needs a B200."""
import numpy as np
import pytest
import torch

from oracle import picovdb_oracle as O

from _torch_oracle import TorchOracle

pytestmark = pytest.mark.gpu


def _fill(store, rows, dim, seed, oracles=(), eligible=None):
    """Generate + upsert the rows chunk by chunk; every TorchOracle in `oracles` scores the same raw
    chunk (`eligible`: optional global bool mask (torch, on the device) per oracle)."""
    dev = torch.device("cuda", 0)
    gen = torch.Generator(device=dev).manual_seed(seed)
    chunk = max(1, (256 << 20) // (dim * 4))
    stream = torch.cuda.current_stream().cuda_stream
    for r0 in range(0, rows, chunk):
        m = min(chunk, rows - r0)
        x = torch.randn(m, dim, device=dev, generator=gen)
        store.upsert_range_dev(x.data_ptr(), r0, m, stream=stream)
        for i, orc in enumerate(oracles):
            el = None if eligible is None or eligible[i] is None else eligible[i][r0:r0 + m]
            orc.update(x, r0, el)
        torch.cuda.synchronize()


def _agree(a_rows, a_sc, b_rows, b_sc, min_same, rtol):
    same = a_rows == b_rows
    assert same.mean() >= min_same, same.mean()
    np.testing.assert_allclose(a_sc[same], b_sc[same], rtol=rtol, atol=2e-6)


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


def test_c2_full_size_single_query(store_factory):
    """C2: 1M x 1024 fp32, single query top-10."""
    n, dim, k = 1_000_000, 1024, 10
    s = store_factory(dim, reserve_rows=n)
    q_rand = np.random.default_rng(99).standard_normal((16, dim)).astype(np.float32)
    orc = TorchOracle(q_rand, k, torch.device("cuda", 0))
    _fill(s, n, dim, 123, [orc])
    ref_s, ref_r = orc.result()
    got_s, got_r = s.search(q_rand, k, precision="f32", scan_only=True)
    O.compare_topk(got_s, got_r, ref_s, ref_r, rtol=1e-5, atol=2e-6)       # exact scan vs independent brute force
    got_s, got_r = s.search(q_rand, k, precision="tf32")                    # tensor-core batch + guard
    O.compare_topk(got_s, got_r, ref_s, ref_r, rtol=1e-5, atol=2e-6)
    probe = np.array([0, 31, 500_000, 999_999])
    qv = s.fetch_rows(probe)
    np.testing.assert_allclose(np.linalg.norm(qv, axis=1), 1.0, rtol=1e-6)
    sc, rows = s.search(qv, k, precision="f32", scan_only=True)
    assert rows[:, 0].tolist() == probe.tolist()
    np.testing.assert_allclose(sc[:, 0], 1.0, rtol=1e-5)
    assert np.all(np.diff(sc, axis=1) <= 0) and (rows >= 0).all() and (rows < n).all()
    sc2, rows2 = s.search(qv, k, precision="f32", scan_only=True)
    np.testing.assert_array_equal(rows, rows2)
    np.testing.assert_array_equal(sc, sc2)
    # the runner-up list must be what the exact dot products of those rows say
    for qi in range(len(probe)):
        exact = s.fetch_rows(rows[qi]) @ qv[qi]
        np.testing.assert_allclose(sc[qi], exact, rtol=1e-5, atol=2e-6)
    # tensor-core batch path on the same queries agrees with the scan
    sb, rb = s.search(qv, k, precision="tf32")
    _agree(rb, sb, rows, sc, 0.97, 1e-5)


def test_c3_full_size_batch(store_factory):
    """C3: 10M x 768 fp32 / tf32 + re-scoring, 4096-query batch, top-100."""
    n, dim, nq, k = 10_000_000, 768, 4096, 100
    s = store_factory(dim, reserve_rows=n)
    rng = np.random.default_rng(99)
    q = rng.standard_normal((nq, dim)).astype(np.float32)
    pick_ref = np.arange(64, 64 + 48)                       # 48 random queries get an independent expected value
    orc = TorchOracle(q[pick_ref], k, torch.device("cuda", 0))
    _fill(s, n, dim, 123, [orc])
    own = rng.choice(n, 64, replace=False)
    q[:64] = s.fetch_rows(own)  # 64 queries are stored rows: they must find themselves first
    sc, rows = s.search(q, k, precision="tf32")
    ref_s, ref_r = orc.result()
    stats = O.compare_topk(sc[pick_ref], rows[pick_ref], ref_s, ref_r, rtol=1e-5, atol=2e-6)
    assert stats["recall"] >= 0.999, stats
    assert s.guard_stats()[0] <= 8, "well separated Gaussian rows must hardly ever need the exact-scan fallback"
    assert rows.shape == (nq, k) and (rows >= 0).all() and (rows < n).all()
    assert np.all(np.diff(sc, axis=1) <= 0)
    assert rows[:64, 0].tolist() == own.tolist()
    np.testing.assert_allclose(sc[:64, 0], 1.0, rtol=1e-5)
    for qi in range(nq):  # no row twice in a list
        if qi % 512 == 0:
            assert len(set(rows[qi].tolist())) == k
    # exact fp32 scan on a few of the queries: same ids (up to near-ties), same scores
    pick = np.array([0, 63, 64, 1000, 4095])
    se, re_ = s.search(q[pick], k, precision="f32", scan_only=True)
    _agree(rows[pick], sc[pick], re_, se, 0.99, 1e-5)


def test_c4_full_size_masks(store_factory):
    """C4: 5M x 384 fp32, 30 % deleted + metadata prefilters, top-10: dense scan, sparse scan and the
    batch path must select the same rows, all of them live and inside the prefilter."""
    n, dim, k = 5_000_000, 384, 10
    s = store_factory(dim, reserve_rows=n)
    dead = np.random.default_rng(1).choice(n, int(0.3 * n), replace=False)
    active = np.ones(n, bool)
    active[dead] = False
    cat = np.arange(n) % 10
    q = np.random.default_rng(99).standard_normal((6, dim)).astype(np.float32)
    filters = (None, cat == 0, cat % 2 == 0)
    dev = torch.device("cuda", 0)
    oracles = [TorchOracle(q, k, dev) for _ in filters]
    elig = [torch.from_numpy(active if pf is None else (active & pf)).to(dev) for pf in filters]
    _fill(s, n, dim, 123, oracles, elig)
    s.delete_rows(dead)
    assert s.info().active == int(active.sum())
    for pf, orc in zip(filters, oracles):
        ref_s, ref_r = orc.result()
        sc_all, rows_all = s.search(q, k, prefilter=pf, precision="f32", scan_only=True)
        O.compare_topk(sc_all, rows_all, ref_s, ref_r, rtol=1e-5, atol=2e-6)   # masked scans vs brute force
        sc, rows = s.search(q[:2], k, prefilter=pf, precision="f32", scan_only=True)
        assert active[rows].all() and (pf is None or pf[rows].all())
        assert np.all(np.diff(sc, axis=1) <= 0)
        sb, rb = s.search(q, k, prefilter=pf, precision="tf32")
        assert active[rb].all() and (pf is None or pf[rb].all())
        O.compare_topk(sb, rb, ref_s, ref_r, rtol=1e-5, atol=2e-6)             # masked tensor-core epilogue
        _agree(rb[:2], sb[:2], rows, sc, 0.95, 1e-5)
        if pf is not None:
            # a prefilter that is all ones takes the sparse walk over the same rows as pf=None
            import os
            os.environ["PVDB_SCAN_NO_SPARSE"] = "1"
            try:
                sd, rd = s.search(q[:2], k, prefilter=pf, precision="f32", scan_only=True)
            finally:
                del os.environ["PVDB_SCAN_NO_SPARSE"]
            np.testing.assert_array_equal(rd, rows)
            np.testing.assert_array_equal(sd, sc)


def test_c5_shard_size_bf16(store_factory):
    """C5: one GPU's shard (12.5M x 384) of the 100M-row bf16-only store, single query and batch."""
    n, dim, k = 12_500_000, 384, 10
    s = store_factory(dim, reserve_rows=n, keep_f32=False, bf16_mirror=True)
    q_rand = np.random.default_rng(99).standard_normal((300, dim)).astype(np.float32)
    orc = TorchOracle(q_rand[:32], k, torch.device("cuda", 0))
    _fill(s, n, dim, 123, [orc])
    ref_s, ref_r = orc.result()                       # fp32 scores of the rows BEFORE the bf16 rounding
    one_s, one_r = s.search(q_rand[:32], k, precision="bf16", scan_only=True)
    O.compare_topk(one_s, one_r, ref_s, ref_r, rtol=1e-2, atol=4e-3)
    bat_s, bat_r = s.search(q_rand, k, precision="bf16")
    O.compare_topk(bat_s[:32], bat_r[:32], ref_s, ref_r, rtol=1e-2, atol=4e-3)
    np.testing.assert_array_equal(bat_r[:32], one_r)  # re-scored from the mirror + guard: batch == single
    probe = np.array([7, 6_000_000, n - 1])
    qv = s.fetch_rows(probe)  # bf16-rounded stored rows
    sc, rows = s.search(qv, k, precision="bf16", scan_only=True, normalized=False)
    assert rows[:, 0].tolist() == probe.tolist()
    np.testing.assert_allclose(sc[:, 0], 1.0, rtol=1e-2)
    q = np.random.default_rng(99).standard_normal((300, dim)).astype(np.float32)
    q[:3] = qv
    sb, rb = s.search(q, k, precision="bf16")
    assert rb[:3, 0].tolist() == probe.tolist() and np.all(np.diff(sb, axis=1) <= 0)
    se, re_ = s.search(q[:8], k, precision="bf16", scan_only=True)
    same = rb[:8] == re_
    assert same.mean() >= 0.9          # bf16 x bf16 products vs bf16 x fp32: near-ties may swap
    np.testing.assert_allclose(sb[:8][same], se[same], rtol=1e-2, atol=4e-3)
