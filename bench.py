#!/usr/bin/env python
"""Headline benchmark: exact top-10 cosine search, queries/sec (BASELINE.json `metric`).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Workload at every N (the north-star target config, BASELINE.json configs[4]; 76.8 GB, fits one
B200): **C5 = 100,000,000 x 384 unit-norm rows kept as a bf16 mirror only**, synthetic N(0,1) ->
L2-normalised, DB seed 123 (+rank), query seed 99.  Rows are sharded contiguously over the N ranks
(strong scaling: the database is fixed).

The JSON line carries three measurements of that store:

`value` / `roofline` / `e2e`  -- SINGLE queries, top-10, through the HBM-bound scan kernel.  One "step"
    = `--queries-per-step` (32) independent single queries, each its own call, answered by its own pass of
    the scan kernel over the whole (shard of the) mirror; for N > 1 every rank scans its shard and the per-rank
    top-10 lists are exchanged and merged (picovdb_b200/sharded.py).
    value    = queries/s, queries and results resident in HBM (CUDA events, max over ranks).
    e2e      = queries/s through the public host API (`pvdb_search`: ONE query per call, query in host
               memory, H2D + kernels + D2H + stream sync inside the timed region, wall clock).
    roofline = the scan kernel alone (pre-normalised query => the only kernel launched), CUDA events
               around back-to-back launches; algorithmic bytes = rows*dim*2 + rows/8 per launch.
`exact_batch` -- the same device-resident queries, `--queries-per-step` per exact (scan-only) call: the scan
    kernels then score several queries per pass over the matrix (2 over bf16 rows, 4 over fp32 rows),
    bit-identical to lone queries.  Informational; `value` is always one call and one pass per query.
`batch`  -- the same store answering a 4096-query batch, top-10, through the tcgen05 tensor-core kernel
    (+ exact re-scoring, exactness guard): queries/s device-resident and end to end with host buffers,
    `roofline.bound = "tensor"` with algorithmic flops 2*Q*N*dim against the measured bf16 peak (burst
    and sustained), recall@10 against the exact scan of the same store AND against an independent
    fp32 brute force (torch matmul over the pre-rounding rows) on 64 sampled queries.
`c2`     -- (N = 1 only) BASELINE configs[1]: 1,000,000 x 1024 fp32, single-query top-10, with the
    oracle port timed on the SAME full-size matrix on the host (`c2.cpu_baseline`).

`parity` -- checked inside this run, before timing: single-query and batch results of 64 sampled
    queries against the fp32 brute force (bf16 tolerance 1e-2; ids must agree wherever the reference
    score gap exceeds it) and, for N > 1, the device-side exchange + merge against a host merge of
    every rank's local list (bit-equal) with every row inside its owner's shard.
`cpu_baseline` / `--impl reference` = the oracle's numpy restatement of the reference path
    (oracle/picovdb_oracle.py: sgemv + argpartition + argsort, all host threads).  The reference
    keeps fp32 rows only: 100M x 384 fp32 = 154 GB does not fit the host, so the CPU arm times single
    queries over a BOUNDED fp32 sample of the same rows and scales the time linearly in the row count
    (an explicitly labelled extrapolation; the scan is linear in rows).

Inputs per pass (>= 9.6 GB per GPU) are far larger than the 126 MB L2, so no L2 flush is needed
between iterations (stated in `config.l2`).
"""
from __future__ import annotations

import os
import sys


def usable_cores() -> int:
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except AttributeError:
        return os.cpu_count() or 1


if "reference" in sys.argv:
    # torch.distributed.run exports OMP_NUM_THREADS=1 to every rank; the reference arm is a CPU
    # measurement that must use all the host threads it can at every N -- set before OpenBLAS loads
    for _v in ("OMP_NUM_THREADS", "OPENBLAS_NUM_THREADS", "MKL_NUM_THREADS"):
        os.environ[_v] = str(usable_cores())

import argparse  # noqa: E402
import json  # noqa: E402
import statistics  # noqa: E402
import subprocess  # noqa: E402
import threading  # noqa: E402
import time  # noqa: E402

import numpy as np  # noqa: E402

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "queries/sec (top-10, exact)"
UNIT = "queries/s"
N_REF = 64  # sampled queries checked against the independent fp32 brute force


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="c5", choices=["c2", "c5"],
                    help="c5 (default, the north-star target): 100M x 384 bf16 mirror only; "
                         "c2 (BASELINE configs[1]): 1M x 1024 fp32")
    ap.add_argument("--rows", type=int, default=None)
    ap.add_argument("--dim", type=int, default=None)
    ap.add_argument("--k", type=int, default=10)
    ap.add_argument("--queries-per-step", type=int, default=32)
    ap.add_argument("--batch-queries", type=int, default=4096)
    ap.add_argument("--batch-iters", type=int, default=5)
    ap.add_argument("--no-batch", action="store_true")
    ap.add_argument("--no-c2", action="store_true")
    ap.add_argument("--cpu-queries", type=int, default=16, help="single queries timed for cpu_baseline")
    ap.add_argument("--cpu-sample-rows", type=int, default=8_000_000)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    args.warmup = max(3, args.warmup)  # timing hygiene: never fewer than 3 untimed warm-up steps
    apply_workload(args, args.workload)
    return args


def apply_workload(args, name: str):
    args.workload = name
    if name == "c5":
        args.rows = args.rows or 100_000_000
        args.dim = args.dim or 384
        args.store_dtype, args.precision, args.elem_bytes, args.tol = "bf16", "bf16", 2, 1e-2
    else:
        args.rows = args.rows or 1_000_000
        args.dim = args.dim or 1024
        args.store_dtype, args.precision, args.elem_bytes, args.tol = "f32", "f32", 4, 1e-5
    return args


def workload_config(args):
    world = args.gpus
    what = "bf16 mirror only" if args.store_dtype == "bf16" else "fp32"
    return {
        "workload": (f"{args.workload.upper()}: {args.rows}x{args.dim} unit-norm rows ({what}), "
                     f"single-query top-{args.k} (HBM-bound scan) + {args.batch_queries}-query batch "
                     f"(tcgen05 path) on the same store"),
        "rows": args.rows,
        "dim": args.dim,
        "k": args.k,
        "queries_per_step": args.queries_per_step,
        "batch_queries": args.batch_queries,
        "sharding": f"rows/{world} contiguous per rank" if world > 1 else "none",
        "exchange": "per-rank top-k lists exchanged + merged on the device, once per query (batch: once per batch)"
                    if world > 1 else "none",
        "l2": (f"inputs ({args.rows * args.dim * args.elem_bytes / world / 1e9:.1f} GB per pass per GPU) "
               "larger than L2 (126 MB): no flush needed"),
        "seeds": {"db": 123, "queries": 99},
    }


# ----------------------------------------------------------------------------- clocks
class ClockSampler:
    FIELDS = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index = index
        self.proc = None
        self.lines: list[str] = []

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits",
                 "-lms", "50"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True,
            )
            threading.Thread(target=self._pump, daemon=True).start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self) -> dict:
        if self.proc is not None:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            parts = [p.strip() for p in ln.split(",")]
            if len(parts) < 7:
                continue
            try:
                sm.append(float(parts[0]))
                mx.append(float(parts[1]))
            except ValueError:
                continue
            for name, val in zip(names, parts[3:7]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        return {"sm_mhz": statistics.median(sm), "sm_max_mhz": max(mx), "reasons": sorted(reasons), "samples": len(sm)}


# ----------------------------------------------------------------------------- CPU arm
def host_matrix(rows: int, dim: int, seed: int) -> np.ndarray:
    """Same distribution as the device generator: N(0,1) rows, L2-normalised (fp32)."""
    from concurrent.futures import ThreadPoolExecutor

    from oracle import picovdb_oracle as O

    out = np.empty((rows, dim), dtype=np.float32)
    chunk = 65536

    def fill(i0):
        i1 = min(rows, i0 + chunk)
        g = np.random.default_rng([seed, i0])
        out[i0:i1] = O.normalize_rows_fast(g.standard_normal((i1 - i0, dim), dtype=np.float32))

    with ThreadPoolExecutor(max_workers=min(32, usable_cores())) as ex:
        list(ex.map(fill, range(0, rows, chunk)))
    return out


def set_blas_threads() -> int:
    """Pin the BLAS pool to every usable core (it would otherwise follow OMP_NUM_THREADS, which
    torch.distributed.run sets to 1) and return the thread count actually in effect."""
    want = usable_cores()
    try:
        from threadpoolctl import threadpool_info, threadpool_limits

        threadpool_limits(limits=want, user_api="blas")
        got = [p.get("num_threads", 1) for p in threadpool_info() if p.get("user_api") == "blas"]
        return max(got or [want])
    except Exception:
        return want


def cpu_sample_rows(args) -> int:
    """Rows of the fp32 sample the CPU arm scans: the whole matrix when it fits comfortably in host
    RAM, else a bounded sample (time is then scaled linearly to the full row count)."""
    try:
        import psutil

        avail = psutil.virtual_memory().available
    except Exception:
        avail = 16 << 30
    fit = int(avail / 2.5 / (args.dim * 4))
    return max(1, min(args.rows, args.cpu_sample_rows, fit))


def time_oracle(args, n_queries: int, warm: int = 2):
    """Seconds per single query of the oracle port on this host (median), plus a description."""
    from oracle import picovdb_oracle as O

    rows = cpu_sample_rows(args)
    mat = host_matrix(rows, args.dim, 123)
    qs = np.random.default_rng(99).standard_normal((n_queries + warm, args.dim)).astype(np.float32)
    times = []
    for i in range(n_queries + warm):
        t0 = time.perf_counter()
        qn, _ = O.prepare_queries(qs[i], args.dim)
        O.search(mat, qn, args.k)
        dt = time.perf_counter() - t0
        if i >= warm:
            times.append(dt)
    per_query = statistics.median(times)
    sample = f"{n_queries} single queries over {rows}x{args.dim} fp32 on the host"
    if rows != args.rows:
        sample += (f" (bounded sample of the {args.rows}-row workload: the reference keeps fp32 rows only; "
                   f"per-query time scaled linearly x{args.rows / rows:.1f} to {args.rows} rows -- extrapolation)")
        per_query = per_query * (args.rows / rows)
    del mat
    return per_query, sample, times


def run_reference_arm(args, world, rank):
    """`--impl reference`: the reference's CPU implementation of the path = the oracle port
    (a Python reference cannot travel to the GPU box)."""
    if rank != 0:
        return
    threads = set_blas_threads()
    per_step_q = max(1, min(args.queries_per_step, 2))
    n = per_step_q * args.steps
    t0 = time.perf_counter()
    per_query, sample, times = time_oracle(args, n, warm=max(1, min(args.warmup, 3)))
    qps = 1.0 / per_query
    line = {
        "impl": "reference",
        "metric": METRIC,
        "value": qps,
        "unit": UNIT,
        "n_gpus": args.gpus,
        "steps": args.steps,
        "warmup": args.warmup,
        "ms_per_step": per_query * 1e3 * per_step_q,
        "higher_is_better": True,
        "scaling": "strong",
        "vs_baseline": None,
        "dtype": "f32",
        "data": "synthetic",
        "config": workload_config(args),
        "cpu_baseline": {"value": qps, "unit": UNIT, "cores": threads, "kind": "port",
                         "sample": f"{sample}; {per_step_q} queries per step"},
        "e2e": {"value": qps, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
        "host": {"cpu_count": os.cpu_count(), "usable_cores": usable_cores(), "blas_threads": threads,
                 "omp_num_threads_env": os.environ.get("OMP_NUM_THREADS"), "wall_s": time.perf_counter() - t0},
    }
    emit_json(line)


# ----------------------------------------------------------------------------- GPU arm
def load_peaks() -> dict:
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            j = json.load(f)
        return {"hbm": float(j["hbm_gbs"]), "bf16": float(j["bf16_tflops"]),
                "bf16_sustained": float(j.get("bf16_tflops_sustained", j["bf16_tflops"])),
                "source": "measured (MEASURED_PEAKS.json)"}
    return {"hbm": 6650.0, "bf16": 1590.0, "bf16_sustained": 1400.0, "source": "fallback (B200_PROFILING.md)"}


def load_traffic(key: str):
    path = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(path):
        with open(path) as f:
            return json.load(f).get(key)
    return None


def fill_store(torch, store, dev, n_local, dim, seed, ref_qn, k):
    """Generate this rank's rows on the device (N(0,1), normalised + stored by the fused upsert
    kernel) and, on the way, keep an independent fp32 brute-force top-k of the `ref_qn` queries over
    the SAME rows before any rounding: torch normalise + fp32 matmul + topk, chunk by chunk."""
    gen = torch.Generator(device=dev).manual_seed(seed)
    chunk = 131072
    stream = torch.cuda.current_stream().cuda_stream
    best_s = torch.full((ref_qn.shape[0], k), float("-inf"), device=dev)
    best_r = torch.full((ref_qn.shape[0], k), -1, dtype=torch.int64, device=dev)
    for c0 in range(0, n_local, chunk):
        m = min(chunk, n_local - c0)
        x = torch.randn(m, dim, device=dev, generator=gen)
        store.upsert_range_dev(x.data_ptr(), c0, m, stream=stream)  # fused normalise + scatter
        xn = torch.nn.functional.normalize(x, dim=1)
        sc = ref_qn @ xn.T                                           # fp32 (TF32 matmul is off by default)
        s, r = torch.topk(sc, min(k, m), dim=1)
        cs = torch.cat([best_s, s], dim=1)
        cr = torch.cat([best_r, r + c0], dim=1)
        best_s, idx = torch.topk(cs, k, dim=1)
        best_r = torch.gather(cr, 1, idx)
        torch.cuda.synchronize()
    return best_s, best_r


def host_merge(scores_list, rows_list, k):
    """numpy k-way merge of per-rank (Q, k) lists: (score desc, row asc), -1 rows last."""
    s = np.concatenate(scores_list, axis=1)
    r = np.concatenate(rows_list, axis=1)
    out_s = np.full((s.shape[0], k), -np.inf, dtype=np.float32)
    out_r = np.full((s.shape[0], k), -1, dtype=np.int64)
    for q in range(s.shape[0]):
        ok = r[q] >= 0
        order = np.lexsort((r[q][ok], -s[q][ok].astype(np.float64)))[:k]
        out_s[q, : order.size] = s[q][ok][order]
        out_r[q, : order.size] = r[q][ok][order]
    return out_s, out_r


def check_against_reference(got_s, got_r, ref_s, ref_r, tol):
    """The north star's id rule: scores of common rows agree within `tol` (relative, + small
    absolute), and a reference row may be missing from the result only if its reference score is
    within the tolerance band of the reference k-th score.  Returns (recall, violations)."""
    hits = total = bad = 0
    for q in range(ref_r.shape[0]):
        kth = ref_s[q, -1]
        band = 2.0 * tol * max(abs(float(kth)), 1e-3)
        got = {int(r): float(s) for r, s in zip(got_r[q], got_s[q])}
        for r, s in zip(ref_r[q], ref_s[q]):
            total += 1
            g = got.get(int(r))
            if g is None:
                if s > kth + band:
                    bad += 1
                continue
            hits += 1
            if abs(g - float(s)) > tol * abs(float(s)) + 1e-5:
                bad += 1
    return hits / max(total, 1), bad


def run_b200(args, world, rank, local_rank):
    import torch
    import torch.distributed as dist

    from picovdb_b200 import _native as N
    from picovdb_b200.engine import DeviceStore
    from picovdb_b200.sharded import ShardedSearch, shard_range

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the product has no CPU path (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        # stdout carries exactly one JSON line: NCCL's version banner / warnings go to stderr
        os.environ.setdefault("NCCL_DEBUG", "WARN")
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
        dist.init_process_group("nccl", device_id=dev)
    peaks = load_peaks()
    stream = torch.cuda.current_stream().cuda_stream

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(vals):
        t = torch.tensor(vals, dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return [float(x) for x in t.cpu()]

    def measure(wl, sampler=None, with_batch=True):
        """Build the store of workload `wl` on this rank's shard, check it, time it."""
        r0, r1 = shard_range(wl.rows, world, rank)
        n_local = r1 - r0
        k, qps = wl.k, wl.queries_per_step
        total_steps = wl.steps + wl.warmup
        qgen = torch.Generator(device="cpu").manual_seed(99)
        n_pool = max(qps * min(total_steps, 8), N_REF)
        nb = wl.batch_queries if with_batch else 0
        q_host = torch.randn(max(n_pool, nb), wl.dim, generator=qgen).pin_memory()
        q_dev = q_host.to(dev)
        q_np = q_host.numpy()
        ref_qn = torch.nn.functional.normalize(q_dev[:N_REF], dim=1).contiguous()

        store = DeviceStore(wl.dim, device=local_rank, reserve_rows=max(n_local, 1),
                            keep_f32=wl.store_dtype == "f32", bf16_mirror=wl.store_dtype == "bf16")
        t_fill = time.perf_counter()
        loc_s, loc_r = fill_store(torch, store, dev, n_local, wl.dim, 123 + rank, ref_qn, k)
        loc_r = torch.where(loc_r >= 0, loc_r + r0, loc_r)
        fill_s = time.perf_counter() - t_fill
        sharded = ShardedSearch(store, r0)

        def gather_lists(s, r):
            """every rank's (Q, k) lists on the host, in rank order"""
            if world == 1:
                return [s.cpu().numpy()], [r.cpu().numpy()]
            ss = [torch.empty_like(s) for _ in range(world)]
            rr = [torch.empty_like(r) for _ in range(world)]
            dist.all_gather(ss, s.contiguous())
            dist.all_gather(rr, r.contiguous())
            return [x.cpu().numpy() for x in ss], [x.cpu().numpy() for x in rr]

        # ---- parity, before any timing ---------------------------------------------------------
        parity = {}
        ref_s, ref_r = host_merge(*gather_lists(loc_s, loc_r), k)        # fp32 brute force, all shards
        # (a) single queries through the sharded device path
        s_one = torch.empty((N_REF, k), dtype=torch.float32, device=dev)
        r_one = torch.empty((N_REF, k), dtype=torch.int64, device=dev)
        for i in range(N_REF):
            s, r = sharded.search_dev(q_dev[i:i + 1], k, precision=wl.precision, scan_only=True)
            s_one[i], r_one[i] = s[0], r[0]
        torch.cuda.synchronize()
        one_s, one_r = s_one.cpu().numpy(), r_one.cpu().numpy()
        assert np.all(np.diff(one_s, axis=1) <= 0) and one_r.min() >= 0 and one_r.max() < wl.rows
        rec, bad = check_against_reference(one_s, one_r, ref_s, ref_r, wl.tol)
        parity["single_recall_vs_fp32"] = rec
        parity["single_violations"] = bad
        # (b) N > 1: the device exchange + merge against a host merge of every rank's local list
        if world > 1:
            l_s = torch.empty((N_REF, k), dtype=torch.float32, device=dev)
            l_r = torch.empty((N_REF, k), dtype=torch.int64, device=dev)
            store.search_dev(q_dev.data_ptr(), N_REF, k, l_s.data_ptr(), l_r.data_ptr(), precision=wl.precision,
                             stream=stream, scan_only=True)
            torch.cuda.synchronize()
            ls, lr = gather_lists(l_s, l_r)
            in_shard = True
            for j in range(world):
                a, b = shard_range(wl.rows, world, j)
                live = lr[j] >= 0
                in_shard = in_shard and bool(np.all((lr[j][live] >= a) & (lr[j][live] < b)))
            m_s, m_r = host_merge(ls, lr, k)
            parity["merge_equals_host_merge"] = bool(np.array_equal(m_r, one_r) and np.array_equal(m_s, one_s))
            parity["rows_in_owner_shard"] = in_shard
        # (c) the batch path on the same queries
        batch_out = None
        if with_batch:
            qb = q_dev[:nb]
            sb, rb = sharded.search_dev(qb, k, precision=wl.precision)
            torch.cuda.synchronize()
            bs, br = sb[:N_REF].cpu().numpy(), rb[:N_REF].cpu().numpy()
            rec_b, bad_b = check_against_reference(bs, br, ref_s, ref_r, wl.tol)
            same = float(np.mean([len(set(br[q]) & set(one_r[q])) / k for q in range(N_REF)]))
            parity["batch_recall_vs_fp32"] = rec_b
            parity["batch_violations"] = bad_b
            parity["batch_recall_vs_exact_scan"] = same
            batch_out = {"recall_at_10_vs_exact_scan": same, "recall_at_10_vs_fp32": rec_b,
                         "recall_queries": N_REF}
        flags = torch.tensor([parity.get("single_violations", 0) + parity.get("batch_violations", 0),
                              0 if parity.get("merge_equals_host_merge", True) else 1,
                              0 if parity.get("rows_in_owner_shard", True) else 1], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(flags, op=dist.ReduceOp.MAX)
        parity["checked"] = True
        parity["ok"] = bool(flags.sum().item() == 0)
        assert parity["ok"], f"parity check failed: {parity}"

        # ---- value: device-resident single queries --------------------------------------------
        # ONE query per call: several queries in one exact call would share passes over the matrix (4 per pass
        # on fp32 rows, 2 on bf16 rows) -- that figure is reported separately as `exact_batch`
        def step_device(i):
            sl = q_dev[(i % (n_pool // qps)) * qps:][:qps]
            out = None
            for j in range(qps):
                out = sharded.search_dev(sl[j:j + 1], k, precision=wl.precision, scan_only=True)
            return out

        def step_device_shared(i):
            sl = q_dev[(i % (n_pool // qps)) * qps:][:qps]
            return sharded.search_dev(sl, k, precision=wl.precision, scan_only=True)

        if sampler is not None:
            sampler.start()
        for i in range(wl.warmup):
            step_device(i)
        barrier()
        # nvidia-smi needs ~1 s before its first sample: when the whole timed region is shorter than that,
        # keep the GPUs under the same load with extra untimed steps (same count on every rank)
        pe0, pe1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        pe0.record()
        for i in range(3):
            step_device(i)
        pe1.record()
        torch.cuda.synchronize()
        step_ms = max(max_over_ranks([pe0.elapsed_time(pe1) / 3.0])[0], 1e-3)
        if step_ms * wl.steps < 1500.0:
            for i in range(min(20000, int((1500.0 - step_ms * wl.steps) / step_ms) + 1)):
                step_device(i)
        barrier()
        launches0 = N.kernel_launches()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev0.record()
        for i in range(wl.steps):
            step_device(wl.warmup + i)
        ev1.record()
        barrier()
        launches = N.kernel_launches() - launches0
        dev_ms = ev0.elapsed_time(ev1)

        # ---- exact_batch: the same queries, `qps` per exact call (passes over the matrix are shared) ----
        for i in range(2):
            step_device_shared(i)
        barrier()
        sh_steps = max(1, wl.steps // 2)
        xe0, xe1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        xe0.record()
        for i in range(sh_steps):
            step_device_shared(wl.warmup + i)
        xe1.record()
        barrier()
        shared_ms = xe0.elapsed_time(xe1)

        # ---- e2e: public host API, ONE query per call, host query in, host result out ---------
        def one_e2e(qv):
            return sharded.search(qv, k, precision=wl.precision) if world > 1 else \
                store.search(qv, k, precision=wl.precision)

        def step_e2e(i):
            base = (i % (n_pool // qps)) * qps
            out = None
            for j in range(qps):
                out = one_e2e(q_np[base + j: base + j + 1])
            return out

        for i in range(min(wl.warmup, 3)):
            step_e2e(i)
        barrier()
        t0 = time.perf_counter()
        for i in range(wl.steps):
            step_e2e(wl.warmup + i)
        barrier()
        e2e_s = time.perf_counter() - t0

        # ---- batch: tensor-core path, device resident and end to end ---------------------------
        b_ms = b_e2e_ms = None
        if with_batch:
            qb = q_dev[:nb]
            for _ in range(3):
                sharded.search_dev(qb, k, precision=wl.precision)
            barrier()
            bt = []
            for _ in range(wl.batch_iters):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                sharded.search_dev(qb, k, precision=wl.precision)
                e1.record()
                torch.cuda.synchronize()
                bt.append(e0.elapsed_time(e1))
            b_ms = statistics.median(bt)
            qb_np = q_np[:nb]
            one_e2e(qb_np)
            barrier()
            et = []
            for _ in range(max(2, wl.batch_iters // 2)):
                t0 = time.perf_counter()
                one_e2e(qb_np)
                et.append((time.perf_counter() - t0) * 1e3)
            b_e2e_ms = statistics.median(et)
        clocks = sampler.stop() if sampler is not None else None

        # ---- roofline of the dominant single-query kernel: scan launches only ------------------
        n_qn = min(64, n_pool)
        qn_dev = torch.nn.functional.normalize(q_dev[:n_qn], dim=1).contiguous()
        out_s = torch.empty(k, dtype=torch.float32, device=dev)
        out_r = torch.empty(k, dtype=torch.int64, device=dev)
        n_scan = 50
        for j in range(5):
            store.search_dev(qn_dev[j].data_ptr(), 1, k, out_s.data_ptr(), out_r.data_ptr(), precision=wl.precision,
                             normalized=True, stream=stream)
        torch.cuda.synchronize()
        l0 = N.kernel_launches()
        ev0.record()
        for j in range(n_scan):
            store.search_dev(qn_dev[j % n_qn].data_ptr(), 1, k, out_s.data_ptr(), out_r.data_ptr(),
                             precision=wl.precision, normalized=True, stream=stream)
        ev1.record()
        torch.cuda.synchronize()
        assert N.kernel_launches() - l0 == n_scan, "roofline loop must launch exactly one kernel per query"
        scan_ms = ev0.elapsed_time(ev1) / n_scan

        guard_last, guard_total = store.guard_stats()
        dev_ms, e2e_ms, scan_ms, b_ms, b_e2e_ms, fill_s, guard_total, shared_ms = max_over_ranks(
            [dev_ms, e2e_s * 1e3, scan_ms, b_ms or 0.0, b_e2e_ms or 0.0, fill_s, float(guard_total), shared_ms])
        exchange_mode = sharded.exchange_mode
        sharded.close()
        store.close()
        del sharded, store
        torch.cuda.empty_cache()

        n_queries = wl.steps * qps
        algo_bytes = n_local * wl.dim * wl.elem_bytes + n_local / 8
        achieved = algo_bytes / (scan_ms / 1e3) / 1e9
        res = {
            "value": n_queries / (dev_ms / 1e3),
            "ms_per_step": dev_ms / wl.steps,
            "e2e": {
                "value": n_queries / (e2e_ms / 1e3),
                "unit": UNIT,
                "h2d_bytes_per_step": qps * wl.dim * 4,
                "d2h_bytes_per_step": qps * k * 12,
                "api": ("DeviceStore.search -> pvdb_search" if world == 1 else "ShardedSearch.search")
                       + " (one query per call; host query in, host result out; blocking)",
            },
            "gpu_launches": int(launches),
            "roofline": {
                "bound": "hbm",
                "kernel": ("scan_mma_topk_kernel (bf16 rows on mma.sync, fp32 query split into three bf16 terms; masked "
                           "GEMV + fused top-k)" if wl.store_dtype == "bf16" and not os.environ.get("PVDB_SCAN_NO_MMA")
                           else f"scan_topk_kernel<{wl.store_dtype}> (masked GEMV + fused top-k)"),
                "achieved": achieved,
                "peak": peaks["hbm"],
                "unit": "GB/s",
                "frac": achieved / peaks["hbm"],
                "peak_source": peaks["source"] + " hbm_gbs",
                "algorithmic_bytes_per_launch": algo_bytes,
                "us_per_launch": scan_ms * 1e3,
                "rows_per_gpu": n_local,
                "traffic": load_traffic(f"scan_{wl.store_dtype}_{n_local}x{wl.dim}"),
            },
            "exact_batch": {
                "value": sh_steps * qps / (shared_ms / 1e3),
                "unit": UNIT,
                "queries_per_call": qps,
                "note": ("device-resident exact (scan-only) calls of several queries: the scan kernels score "
                         + ("2 queries per pass over bf16 rows (mma.sync)" if wl.store_dtype == "bf16"
                            else "4 queries per pass over fp32 rows")
                         + ", bit-identical to lone queries; NOT the single-query figure (`value`)"),
            },
            "clocks": clocks,
            "parity": parity,
            "fill_seconds": fill_s,
            "exchange": exchange_mode,
        }
        if with_batch:
            flops = 2.0 * nb * wl.rows * wl.dim
            tf = flops / (b_ms / 1e3) / 1e12
            kind = "bf16" if wl.store_dtype == "bf16" else "tf32 (peak taken as bf16 / 2)"
            scale = 1.0 if wl.store_dtype == "bf16" else 0.5
            batch_out.update({
                "guard_flagged_queries": int(guard_total),   # fell back to the exact scan (all batch calls of this run)
                "queries": nb,
                "k": k,
                "ms_per_batch": b_ms,
                "value": nb / (b_ms / 1e3),
                "unit": UNIT,
                "roofline": {
                    "bound": "tensor",
                    "kernel": f"batch_topk_kernel<{kind}> (tcgen05.mma + fused mask/top-k epilogue) + finalize (re-score)",
                    "flops": flops,
                    "achieved": tf,
                    "unit": "TFLOP/s",
                    "peak": peaks["bf16"] * scale * world,
                    "frac": tf / (peaks["bf16"] * scale * world),
                    "peak_sustained": peaks["bf16_sustained"] * scale * world,
                    "frac_sustained": tf / (peaks["bf16_sustained"] * scale * world),
                    "peak_source": peaks["source"] + f" bf16_tflops / bf16_tflops_sustained x {world} GPU(s)",
                },
                "e2e": {
                    "value": nb / (b_e2e_ms / 1e3),
                    "unit": UNIT,
                    "ms_per_batch": b_e2e_ms,
                    "h2d_bytes_per_batch": nb * wl.dim * 4,
                    "d2h_bytes_per_batch": nb * k * 12,
                    "api": ("DeviceStore.search -> pvdb_search" if world == 1 else "ShardedSearch.search")
                           + " (host query batch in, host results out; blocking)",
                },
            })
            res["batch"] = batch_out
        return res

    # ---- the main workload -------------------------------------------------------------------
    sampler = ClockSampler(local_rank)
    main = measure(args, sampler=sampler, with_batch=not args.no_batch)

    # ---- secondary block: BASELINE configs[1] at full size next to the CPU port (N = 1 only) ----
    c2 = None
    if world == 1 and args.workload == "c5" and not args.no_c2:
        wl2 = argparse.Namespace(**vars(args))
        wl2.rows = wl2.dim = None
        apply_workload(wl2, "c2")
        wl2.steps, wl2.warmup = max(5, min(args.steps, 10)), 3
        r2 = measure(wl2, sampler=None, with_batch=False)
        c2 = {"workload": f"C2: {wl2.rows}x{wl2.dim} fp32 unit-norm rows, single-query top-{wl2.k}",
              "value": r2["value"], "unit": UNIT, "e2e": r2["e2e"], "roofline": r2["roofline"],
              "exact_batch": r2["exact_batch"], "parity": r2["parity"]}
        if not args.no_cpu_baseline:
            threads = set_blas_threads()
            per_query, sample, _ = time_oracle(wl2, args.cpu_queries)
            c2["cpu_baseline"] = {"value": 1.0 / per_query, "unit": UNIT, "cores": threads, "kind": "port",
                                  "sample": sample}

    if rank == 0:
        line = {
            "metric": METRIC,
            "value": main["value"],
            "unit": UNIT,
            "n_gpus": world,
            "steps": args.steps,
            "warmup": args.warmup,
            "ms_per_step": main["ms_per_step"],
            "higher_is_better": True,
            "scaling": "strong",
            "vs_baseline": None,
            "dtype": args.store_dtype,
            "data": "synthetic",
            "config": workload_config(args),
            "e2e": main["e2e"],
            "gpu_launches": main["gpu_launches"],
            "roofline": main["roofline"],
            "clocks": main["clocks"],
            "parity_checked": bool(main["parity"]["checked"] and main["parity"]["ok"]),
            "parity": main["parity"],
            "fill_seconds": main["fill_seconds"],
        }
        line["exchange_mode"] = main["exchange"]   # mechanism actually used (config stays identical to the reference arm's)
        line["exact_batch"] = main["exact_batch"]
        if "batch" in main:
            line["batch"] = main["batch"]
        if c2 is not None:
            line["c2"] = c2
        if world == 1 and not args.no_cpu_baseline:
            threads = set_blas_threads()
            per_query, sample, _ = time_oracle(args, args.cpu_queries)
            line["cpu_baseline"] = {"value": 1.0 / per_query, "unit": UNIT, "cores": threads, "kind": "port",
                                    "sample": sample}
        emit_json(line)
    if world > 1:
        dist.destroy_process_group()


_JSON_OUT = None


def emit_json(line: dict) -> None:
    """The one JSON line of the contract, written to the process's ORIGINAL stdout."""
    data = (json.dumps(line) + "\n").encode()
    if _JSON_OUT is None:
        sys.stdout.write(data.decode())
        sys.stdout.flush()
    else:
        os.write(_JSON_OUT, data)


def _reserve_stdout() -> None:
    """Keep stdout for the JSON line only: libraries that print to fd 1 (NCCL's version banner does,
    whatever NCCL_DEBUG_FILE says) are pointed at stderr for the rest of the run."""
    global _JSON_OUT
    sys.stdout.flush()
    _JSON_OUT = os.dup(1)
    os.dup2(2, 1)


def main():
    args = parse_args()
    _reserve_stdout()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference_arm(args, world, rank)
        return
    if world != args.gpus and world == 1 and args.gpus > 1:
        raise SystemExit("bench.py: --gpus N > 1 must be launched with torch.distributed.run (one rank per GPU)")
    run_b200(args, world, rank, local_rank)


if __name__ == "__main__":
    main()
