"""Peer-memory exchange and the single-process device group (csrc/exchange.cu, csrc/group.cu).

A one-GPU box can only run the exchange against its own mailbox (PVDB_EXCHANGE_SELF=1: the kernels
still publish, raise the flag, wait for it and merge -- every step except the remote write); the
multi-device tests need >= 2 GPUs.  Cross-process parity of the fused exchange with the NCCL path is
in tests/test_gpu_sharded.py.
"""
import os

import numpy as np
import pytest

from oracle import picovdb_oracle as O

pytestmark = pytest.mark.gpu


def _gauss(n, dim, seed):
    return np.random.default_rng(seed).standard_normal((n, dim)).astype(np.float32)


def _ngpu():
    import torch

    return torch.cuda.device_count()


def test_self_mailbox_exchange_matches_plain_search(monkeypatch):
    from picovdb_b200.engine import DeviceStore, Exchange

    monkeypatch.setenv("PVDB_EXCHANGE_SELF", "1")
    dim, n = 96, 30_000
    st = DeviceStore(dim, device=0, bf16_mirror=True)
    st.upsert_range(_gauss(n, dim, 1), 0)
    st.delete_rows(np.arange(0, n, 7))
    st.set_row_base(1_000_000)                      # as a shard: result rows are global
    ex = Exchange(0, 1, 0, 1 << 16)
    queries = _gauss(300, dim, 2)
    pf = (np.arange(n) % 3) == 0
    try:
        launches = 0
        for k in (1, 10, 32, 33, 128):
            for prefilter in (None, pf):
                want_s, want_r = st.search(queries[:5], k, prefilter=prefilter, precision="f32")
                for rep in range(3):                # consecutive launches alternate the slot parity
                    got_s, got_r = st.search_exchange(ex, queries[:5], k, prefilter=prefilter, precision="f32")
                    # scan path: k <= 32 scans several queries per pass into scratch and exchanges once;
                    # beyond that one scan per query with the exchange fused into its last block
                    launches += 1 if k <= 32 else 5
                    np.testing.assert_array_equal(got_r, want_r)
                    np.testing.assert_array_equal(got_s, want_s)
                got_s, got_r = st.search_exchange(ex, queries[3:4], k, prefilter=prefilter, precision="f32")
                launches += 1                       # a lone query: always the fused form
                np.testing.assert_array_equal(got_r, want_r[3:4])
                np.testing.assert_array_equal(got_s, want_s[3:4])
                assert want_r.min() >= 1_000_000
        assert ex.launches() == launches
        for prec in ("tf32", "bf16"):               # batch path: one exchange + merge launch per call
            want_s, want_r = st.search(queries, 10, precision=prec)
            got_s, got_r = st.search_exchange(ex, queries, 10, precision=prec)
            launches += 1
            np.testing.assert_array_equal(got_r, want_r)
            np.testing.assert_array_equal(got_s, want_s)
        assert ex.launches() == launches
        with pytest.raises(Exception):              # beyond the fused exchange: the caller must fall back
            st.search_exchange(ex, queries[:2], 129, precision="f32")
    finally:
        ex.close()
        st.close()


def test_group_of_one_device_equals_a_store(tmp_path):
    from picovdb_b200.engine import DeviceStore
    from picovdb_b200.group import GroupStore

    dim, n, k = 64, 9000, 10
    raw = _gauss(n, dim, 3)
    g = GroupStore(dim, [0], reserve_rows=n)
    st = DeviceStore(dim, device=0)
    try:
        g.upsert_range(raw, 0)
        st.upsert_range(raw, 0)
        dead = np.arange(5, n, 11)
        g.delete_rows(dead)
        st.delete_rows(dead)
        q = _gauss(40, dim, 4)
        for nq in (1, 40):
            for pf in (None, (np.arange(n) % 4) == 1):
                a = g.search(q[:nq], k, prefilter=pf)
                b = st.search(q[:nq], k, prefilter=pf)
                np.testing.assert_array_equal(a[1], b[1])
                np.testing.assert_array_equal(a[0], b[0])
        np.testing.assert_array_equal(g.download(), st.download())
        np.testing.assert_array_equal(g.active_mask(), st.active_mask())
        np.testing.assert_array_equal(g.fetch_rows([3, 8999, 0]), st.fetch_rows([3, 8999, 0]))
    finally:
        g.close()
        st.close()


def test_db_devices_kwarg_single_device(tmp_path):
    from picovdb_b200 import K_ID, K_VECTOR, PicoVectorDB

    db = PicoVectorDB(embedding_dim=8, storage_file=str(tmp_path / "one"), devices=[0])
    vecs = _gauss(50, 8, 5)
    db.upsert([{K_VECTOR: vecs[i], K_ID: str(i)} for i in range(50)])
    assert db.query(vecs[7], top_k=1)[0][K_ID] == "7"
    with pytest.raises(ValueError, match="capacity"):
        PicoVectorDB(embedding_dim=8, storage_file=str(tmp_path / "two"), devices=[0, 1])
    db.close()


needs2 = pytest.mark.skipif("_ngpu() < 2", reason="needs >= 2 GPUs")


@needs2
def test_two_stores_on_two_devices_in_one_process():
    """function attributes are per device -- the tensor-core batch path must work on the
    second GPU of a process as well as on the first."""
    from picovdb_b200.engine import DeviceStore

    dim, n, k = 128, 20_000, 10
    raw = _gauss(n, dim, 6)
    q = _gauss(64, dim, 7)
    out = []
    for dev in (0, 1):
        st = DeviceStore(dim, device=dev, bf16_mirror=True)
        st.upsert_range(raw, 0)
        out.append((st.search(q, k, precision="tf32"), st.search(q, k, precision="bf16"), st.search(q[:1], k)))
        st.close()
    for a, b in zip(out[0], out[1]):
        np.testing.assert_array_equal(a[1], b[1])
        np.testing.assert_array_equal(a[0], b[0])


@needs2
def test_group_over_devices_equals_one_store():
    from picovdb_b200.engine import DeviceStore
    from picovdb_b200.group import GroupStore

    ndev = min(_ngpu(), 4)
    dim, n = 96, 50_001
    raw = _gauss(n, dim, 8)
    g = GroupStore(dim, list(range(ndev)), reserve_rows=n, bf16_mirror=True)
    st = DeviceStore(dim, device=0, bf16_mirror=True)
    try:
        g.upsert_range(raw, 0)
        st.upsert_range(raw, 0)
        dead = np.arange(2, n, 5)
        g.delete_rows(dead)
        st.delete_rows(dead)
        q = _gauss(300, dim, 9)
        pf = (np.arange(n) % 10) < 3
        for nq, k, prec in ((1, 10, "f32"), (7, 100, "f32"), (300, 10, "tf32"), (300, 10, "bf16"), (3, 300, "f32")):
            for prefilter in (None, pf):
                for rep in range(2):
                    a = g.search(q[:nq], k, prefilter=prefilter, precision=prec)
                b = st.search(q[:nq], k, prefilter=prefilter, precision=prec)
                np.testing.assert_array_equal(a[1], b[1], err_msg=f"{nq} {k} {prec}")
                np.testing.assert_allclose(a[0], b[0], rtol=0, atol=0)
        np.testing.assert_array_equal(g.download(), st.download())
        keep = np.flatnonzero(st.active_mask())
        g.compact(keep)
        st.compact(keep)
        a, b = g.search(q[:9], 10), st.search(q[:9], 10)
        np.testing.assert_array_equal(a[1], b[1])
    finally:
        g.close()
        st.close()


@needs2
def test_db_over_two_devices_matches_oracle(tmp_path):
    from picovdb_b200 import K_ID, K_VECTOR, PicoVectorDB

    dim, n = 48, 3000
    vecs = _gauss(n, dim, 10)
    items = [{K_VECTOR: vecs[i], K_ID: f"r{i}", "cat": i % 4} for i in range(n)]
    db = PicoVectorDB(embedding_dim=dim, storage_file=str(tmp_path / "grp"), devices=[0, 1], capacity=4096)
    odb = O.OracleDB(dim)
    db.upsert(items)
    odb.upsert(items)
    gone = [f"r{i}" for i in range(0, n, 6)]
    db.delete(gone)
    odb.delete(gone)
    q = _gauss(5, dim, 11)
    ids = lambda res: [[r[K_ID] for r in lst] for lst in res]  # noqa: E731
    assert ids(db.query(q, top_k=7)) == ids(odb.query(q, top_k=7))
    assert ids(db.query(q, top_k=7, where={"cat": 1})) == ids(odb.query(q, top_k=7, where={"cat": 1}))
    assert [r[K_ID] for r in db.query(q[0], top_k=3)] == [r[K_ID] for r in odb.query(q[0], top_k=3)]
    db.save()
    db.close()
    again = PicoVectorDB(embedding_dim=dim, storage_file=str(tmp_path / "grp"), devices=[0, 1], capacity=4096)
    assert ids(again.query(q, top_k=7)) == ids(odb.query(q, top_k=7))
    again.vacuum()
    assert ids(again.query(q, top_k=7)) == ids(odb.query(q, top_k=7))
    again.close()
