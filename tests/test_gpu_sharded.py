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
if rank == 0:
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
