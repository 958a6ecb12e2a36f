"""Store growth: many small appends (the reference's bench/many_upserts.py pattern, where every
upsert re-allocates and copies the whole matrix, pico_vdb.py:451-462) keep every earlier row intact
while the device buffers grow, and searches see all of them.  Needs a B200."""
import time

import numpy as np
import pytest

from oracle import picovdb_oracle as O

pytestmark = pytest.mark.gpu


def test_many_small_appends_grow_in_place():
    from picovdb_b200.engine import DeviceStore

    dim, step, rounds = 96, 700, 120
    rng = np.random.default_rng(0)
    raw = rng.standard_normal((step * rounds, dim)).astype(np.float32)
    s = DeviceStore(dim, bf16_mirror=True)
    caps = []
    t0 = time.perf_counter()
    for i in range(rounds):
        s.upsert_range(raw[i * step:(i + 1) * step], i * step)
        caps.append(int(s.info().capacity))
    elapsed = time.perf_counter() - t0
    assert caps[-1] >= step * rounds and len(set(caps)) >= 5   # grew several times
    got = s.download()
    np.testing.assert_allclose(got, O.normalize_rows_fast(raw), rtol=1e-6, atol=1e-7)
    assert s.info().active == step * rounds
    qn, _ = O.prepare_queries(raw[[5, 40000, step * rounds - 1]], dim)
    sc, rows = s.search(qn, 3, precision="f32", normalized=True)
    assert rows[:, 0].tolist() == [5, 40000, step * rounds - 1]
    sc16, rows16 = s.search(qn, 3, precision="bf16", normalized=True)
    assert rows16[:, 0].tolist() == rows[:, 0].tolist()
    assert elapsed < 20.0
    s.close()


def test_picovectordb_single_item_upserts(tmp_path):
    from picovdb_b200 import K_ID, K_VECTOR, PicoVectorDB

    db = PicoVectorDB(embedding_dim=32, storage_file=str(tmp_path / "many"))
    rng = np.random.default_rng(1)
    vecs = rng.standard_normal((1500, 32)).astype(np.float32)
    for i in range(1500):                      # bench/many_upserts.py:31-39
        db.upsert([{K_VECTOR: vecs[i], K_ID: i}])
    assert len(db) == 1500
    res = db.query(vecs[1234], top_k=1)
    assert res[0][K_ID] == 1234
    np.testing.assert_allclose(db._vectors, O.normalize_rows_fast(vecs), rtol=1e-6, atol=1e-7)
    db.close()
