"""Multi-GPU row-sharded search on real GPUs (NCCL): skipped unless >= 2 devices are visible."""
import os
import socket
import subprocess
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = r"""
import os, sys, json
import numpy as np, torch, torch.distributed as dist
sys.path.insert(0, os.environ["PVDB_ROOT"])
from oracle import picovdb_oracle as O
from picovdb_b200.engine import DeviceStore
from picovdb_b200.sharded import ShardedSearch, shard_range
rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"]); lr = int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
n, dim, k = 40000, 96, 10
rng = np.random.default_rng(7)
full = rng.standard_normal((n, dim)).astype(np.float32)
queries = rng.standard_normal((9, dim)).astype(np.float32)
r0, r1 = shard_range(n, world, rank)
st = DeviceStore(dim, device=lr)
st.upsert_range(full[r0:r1], 0)
sh = ShardedSearch(st, r0)
s, r = sh.search(queries, k, precision="f32")
s1, r1_ = sh.search(queries[:1], k, precision="f32")
sb, rb = sh.search(queries, k)  # batch path (tf32 + rescoring) per shard, then merge
mode = sh.exchange_mode
# the same calls through the round-1 exchange (NCCL all-gather + merge kernel) must give the same bits
os.environ["PVDB_NO_PEER_EXCHANGE"] = "1"
s_n, r_n = sh.search(queries, k, precision="f32")
sb_n, rb_n = sh.search(queries, k)
del os.environ["PVDB_NO_PEER_EXCHANGE"]
assert np.array_equal(r, r_n) and np.array_equal(s, s_n), "fused exchange != nccl exchange (scan path)"
assert np.array_equal(rb, rb_n) and np.array_equal(sb, sb_n), "fused exchange != nccl exchange (batch path)"
# device-resident entry point, a larger k (four list slots per lane) and many calls in a row (slot parity)
qd = torch.from_numpy(queries).cuda()
for kk in (1, 33, 100):
    for rep in range(3):
        sd, rd = sh.search_dev(qd, kk, precision="f32", scan_only=True)
        torch.cuda.synchronize()
    os.environ["PVDB_NO_PEER_EXCHANGE"] = "1"
    sd_n, rd_n = sh.search_dev(qd, kk, precision="f32", scan_only=True)
    torch.cuda.synchronize()
    del os.environ["PVDB_NO_PEER_EXCHANGE"]
    assert torch.equal(rd.cpu(), rd_n.cpu()) and torch.equal(sd.cpu(), sd_n.cpu()), f"k={kk}"
big = rng.standard_normal((700, dim)).astype(np.float32)
sB, rB = sh.search(big, 10)
os.environ["PVDB_NO_PEER_EXCHANGE"] = "1"
sB_n, rB_n = sh.search(big, 10)
del os.environ["PVDB_NO_PEER_EXCHANGE"]
assert np.array_equal(rB, rB_n) and np.array_equal(sB, sB_n), "700-query batch"
sh.close()
if rank == 0:
    print("EXCHANGE_MODE", mode)
    store = O.normalize_rows(full)
    qn, _ = O.prepare_queries(queries, dim)
    ref_s, ref_r = O.search(store, qn, k)
    O.compare_topk(s, r, ref_s, ref_r, rtol=1e-5, atol=2e-6)
    O.compare_topk(s1, r1_, ref_s[:1], ref_r[:1], rtol=1e-5, atol=2e-6)
    assert O.recall_at_k(rb, ref_r) >= 0.99
    print("SHARDED_OK", world)
st.close()
dist.destroy_process_group()
"""


def test_sharded_search_nccl(tmp_path):
    import torch

    ngpu = torch.cuda.device_count()
    if ngpu < 2:
        pytest.skip("needs >= 2 GPUs")
    world = 2 if ngpu < 4 else 4
    script = tmp_path / "worker.py"
    script.write_text(WORKER)
    with socket.socket() as sck:
        sck.bind(("127.0.0.1", 0))
        port = sck.getsockname()[1]
    env = dict(os.environ, PVDB_ROOT=ROOT)
    out = subprocess.run(
        [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
         "--master-addr", "127.0.0.1", "--master-port", str(port), str(script)],
        capture_output=True, text=True, timeout=600, env=env,
    )
    assert out.returncode == 0, out.stdout[-3000:] + out.stderr[-3000:]
    assert f"SHARDED_OK {world}" in out.stdout


DB_WORKER = r"""
import os, sys
import numpy as np, torch, torch.distributed as dist
sys.path.insert(0, os.environ["PVDB_ROOT"])
from oracle import picovdb_oracle as O
from picovdb_b200 import K_ID, K_VECTOR, PicoVectorDB
from picovdb_b200.sharded import ShardedPicoVectorDB as ShardedDB, shard_range
rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"]); lr = int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dist.init_process_group("nccl", device_id=torch.device("cuda", lr))

path = os.path.join(os.environ["PVDB_TMP"], "db")
n, dim = 5000, 64
rng = np.random.default_rng(11)
vecs = rng.standard_normal((n, dim)).astype(np.float32)
db = ShardedDB(embedding_dim=dim, storage_file=path, capacity=8192, no_faiss=True, device=lr)
db.upsert([{K_VECTOR: vecs[i], K_ID: f"r{i}", "cat": i % 5} for i in range(n)])
where = dict(db._id2idx)                      # id -> global row (dealt out over the shards)
dead = [f"r{i}" for i in range(0, n, 9)]
db.delete(dead)
q = rng.standard_normal((6, dim)).astype(np.float32)
res = db.query(q, top_k=8)
res_w = db.query(q, top_k=8, where={"cat": 2})            # evaluated on every shard's GPU (search_where)
some = [f"r{i}" for i in range(0, n, 3)]
res_wi = db.query(q, top_k=8, where={"cat": {"$in": [1, 4]}}, ids=some)
k_eff_wi = db._last_k_eff
one = db.query(q[0], top_k=8)
got_vec = db.get(["r1"], include_vector=True)[0][K_VECTOR]
db.save()
db.close()
db2 = ShardedDB(embedding_dim=dim, storage_file=path, capacity=8192, no_faiss=True, device=lr)
res2 = db2.query(q, top_k=8)
db2.vacuum()                                   # compaction moves rows between the shards
res3 = db2.query(q, top_k=8)
res3_w = db2.query(q, top_k=8, where={"cat": 2})   # columns were dropped by the compaction and are re-sent
n_after = len(db2)
db2.close()
if rank == 0:
    store = O.normalize_rows(vecs)
    alive = np.ones(n, bool); alive[::9] = False
    qn, _ = O.prepare_queries(q, dim)
    ref_s, ref_r = O.search(store, qn, 8, alive)
    ids = lambda rs: [[r[K_ID] for r in lst] for lst in rs]
    want = [[f"r{j}" for j in row] for row in ref_r]
    assert ids(res) == want and ids(res2) == want and [r[K_ID] for r in one] == want[0]
    assert ids(res3) == want and ids(res3_w) == ids(res_w) and n_after == int(alive.sum())
    ref_s2, ref_r2 = O.search(store, qn, 8 + 32, alive, (np.arange(n) % 5) == 2)
    assert ids(res_w) == [[f"r{j}" for j in row[:8]] for row in ref_r2]
    pf = np.isin(np.arange(n) % 5, [1, 4]) & (np.arange(n) % 3 == 0)
    ref_s3, ref_r3 = O.search(store, qn, 8 + 32, alive, pf)
    assert ids(res_wi) == [[f"r{j}" for j in row[:8]] for row in ref_r3]
    assert k_eff_wi == min(40, int((pf & alive).sum()))
    np.testing.assert_allclose(got_vec, store[1], rtol=1e-6, atol=1e-7)
    saved = np.load(path + ".vecs.npy")
    assert saved.shape == (8192, dim)
    np.testing.assert_allclose(saved[where["r1"]], store[1], rtol=1e-6, atol=1e-7)
    np.testing.assert_allclose(saved[where["r4999"]], store[4999], rtol=1e-6, atol=1e-7)
    assert not saved[where["r0"]].any() and not saved[where["r9"]].any()   # deleted rows are zero-filled, as in the reference
    r0_, r1_ = shard_range(8192, world, 0)
    on0 = sum(1 for r in where.values() if r < r1_)
    assert abs(on0 - n // world) <= 1                                        # rows are dealt out evenly
    print("SHARDED_DB_OK", world)
dist.destroy_process_group()
"""


def test_sharded_db_nccl(tmp_path):
    """The drop-in class over a row-sharded store (ShardedStore): SPMD calls on every rank."""
    import torch

    ngpu = torch.cuda.device_count()
    if ngpu < 2:
        pytest.skip("needs >= 2 GPUs")
    world = 2
    script = tmp_path / "db_worker.py"
    script.write_text(DB_WORKER)
    with socket.socket() as sck:
        sck.bind(("127.0.0.1", 0))
        port = sck.getsockname()[1]
    env = dict(os.environ, PVDB_ROOT=ROOT, PVDB_TMP=str(tmp_path))
    out = subprocess.run(
        [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
         "--master-addr", "127.0.0.1", "--master-port", str(port), str(script)],
        capture_output=True, text=True, timeout=600, env=env,
    )
    assert out.returncode == 0, out.stdout[-3000:] + out.stderr[-3000:]
    assert f"SHARDED_DB_OK {world}" in out.stdout
