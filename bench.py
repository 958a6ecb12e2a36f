#!/usr/bin/env python
"""Headline benchmark: exact top-10 cosine search, queries/sec (BASELINE.json `metric`).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Workload at every N: BASELINE.json configs[1] -- 1,000,000 x 1024 fp32 unit-norm rows (synthetic
N(0,1) -> L2-normalised, DB seed 123 (+rank), query seed 99), single-query top-10 through the
HBM-bound scan kernel.  One "step" = `--queries-per-step` (32) independent single queries, each
answered by its own pass of the scan kernel over the whole (shard of the) matrix.  For N > 1 the
rows are sharded contiguously over the ranks (strong scaling: the database is fixed), every rank
scans its shard for each query of the step, and ONE all-gather + merge kernel per step combines
the per-rank top-10 lists (picovdb_b200/sharded.py).

`value`  = queries/s with queries and results resident in HBM (CUDA events, max over ranks).
`e2e`    = queries/s through the public host API (`DeviceStore.search`, i.e. the `pvdb_search`
           C-ABI call): ONE query per call, query in pinned host memory, H2D + kernels + D2H +
           stream sync inside the timed region (wall clock around the blocking calls).
`roofline` = the scan kernel alone (pre-normalised query => the only kernel launched), CUDA
           events around back-to-back launches; algorithmic bytes = rows*dim*4 + rows/8.
`cpu_baseline` / `--impl reference` = the oracle's numpy restatement of the reference path
           (oracle/picovdb_oracle.py: sgemv + argpartition + argsort, all host threads) on the same
           1M x 1024 workload, a bounded number of single queries.

The input (4.1 GB per query pass) is far larger than the 126 MB L2, so no L2 flush is needed
between iterations (stated in `config.l2`).
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "queries/sec (top-10, exact)"
UNIT = "queries/s"


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="c2", choices=["c2", "c5"],
                    help="c2 (default, BASELINE configs[1]): 1M x 1024 fp32; c5: 100M x 384 bf16 mirror only "
                         "(the north-star target config, single queries), sharded over the ranks")
    ap.add_argument("--rows", type=int, default=None)
    ap.add_argument("--dim", type=int, default=None)
    ap.add_argument("--k", type=int, default=10)
    ap.add_argument("--queries-per-step", type=int, default=32)
    ap.add_argument("--cpu-queries", type=int, default=24, help="single queries timed for cpu_baseline")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    args.warmup = max(3, args.warmup)  # timing hygiene: never fewer than 3 untimed warm-up steps
    if args.workload == "c5":
        args.rows = args.rows or 100_000_000
        args.dim = args.dim or 384
        args.store_dtype, args.precision, args.elem_bytes = "bf16", "bf16", 2
    else:
        args.rows = args.rows or 1_000_000
        args.dim = args.dim or 1024
        args.store_dtype, args.precision, args.elem_bytes = "f32", "f32", 4
    return args


def workload_config(args, world):
    return {
        "workload": (f"{args.workload.upper()}: {args.rows}x{args.dim} {args.store_dtype} unit-norm rows, "
                     f"single-query top-{args.k} (HBM-bound scan)"),
        "rows": args.rows,
        "dim": args.dim,
        "k": args.k,
        "queries_per_step": args.queries_per_step,
        "sharding": f"rows/{world} contiguous per rank" if world > 1 else "none",
        "exchange": "one all-gather + merge kernel per step" if world > 1 else "none",
        "l2": f"inputs ({args.rows * args.dim * args.elem_bytes / world / 1e9:.1f} GB per pass per GPU) larger than L2 (126 MB): no flush needed",
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

    with ThreadPoolExecutor(max_workers=min(32, os.cpu_count() or 1)) as ex:
        list(ex.map(fill, range(0, rows, chunk)))
    return out


def blas_threads() -> int:
    try:
        from threadpoolctl import threadpool_info

        return max([p.get("num_threads", 1) for p in threadpool_info() if p.get("user_api") == "blas"] or [1])
    except Exception:
        return os.cpu_count() or 1


def cpu_rows_that_fit(args) -> int:
    try:
        import psutil

        avail = psutil.virtual_memory().available
    except Exception:
        avail = 16 << 30
    need = args.rows * args.dim * 4
    if args.workload == "c5":
        return min(args.rows, 2_000_000)  # the reference holds fp32 only: 100M x 384 = 154 GB; bounded sample
    if need * 1.3 < avail:
        return args.rows
    return max(1, int(avail / 1.3 / (args.dim * 4)))


def time_oracle(args, n_queries: int, warm: int = 2):
    """Seconds per single query of the oracle port on this host (median), plus a description."""
    from oracle import picovdb_oracle as O

    rows = cpu_rows_that_fit(args)
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
    scale = rows / args.rows  # < 1 only when the full matrix does not fit in host RAM
    sample = f"{n_queries} single queries over {rows}x{args.dim} fp32 on the host"
    if rows != args.rows:
        sample += f" (bounded sample; time scaled linearly to {args.rows} rows)"
        per_query = per_query / scale
    return per_query, sample, times


def run_reference_arm(args, world, rank):
    """`--impl reference`: the reference's CPU implementation of the path = the oracle port
    (a Python reference cannot travel to the GPU box)."""
    if rank != 0:
        return
    threads = blas_threads()
    per_step_q = max(1, min(args.queries_per_step, 4))
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
        "config": workload_config(args, 1),
        "cpu_baseline": {"value": qps, "unit": UNIT, "cores": threads, "kind": "port",
                         "sample": f"{sample}; {per_step_q} queries per step"},
        "e2e": {"value": qps, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
        "host": {"cpu_count": os.cpu_count(), "blas_threads": threads, "wall_s": time.perf_counter() - t0},
    }
    emit_json(line)


# ----------------------------------------------------------------------------- GPU arm
def load_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def load_traffic(key: str):
    path = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(path):
        with open(path) as f:
            return json.load(f).get(key)
    return None


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

    r0, r1 = shard_range(args.rows, world, rank)
    n_local = r1 - r0
    store = DeviceStore(args.dim, device=local_rank, reserve_rows=max(n_local, 1),
                        keep_f32=args.store_dtype == "f32", bf16_mirror=args.store_dtype == "bf16")
    gen = torch.Generator(device=dev).manual_seed(123 + rank)
    chunk = 131072
    stream = torch.cuda.current_stream().cuda_stream
    for c0 in range(0, n_local, chunk):
        m = min(chunk, n_local - c0)
        x = torch.randn(m, args.dim, device=dev, generator=gen)
        store.upsert_range_dev(x.data_ptr(), c0, m, stream=stream)  # fused normalise + scatter
        torch.cuda.synchronize()
    sharded = ShardedSearch(store, r0)

    qps, k = args.queries_per_step, args.k
    total_steps = args.steps + args.warmup
    qgen = torch.Generator(device="cpu").manual_seed(99)
    n_pool = qps * min(total_steps, 8)
    q_host = torch.randn(n_pool, args.dim, generator=qgen).pin_memory()
    q_dev = q_host.to(dev)
    q_np = q_host.numpy()

    def step_device(i):
        sl = q_dev[(i % (n_pool // qps)) * qps:][:qps]
        return sharded.search_dev(sl, k, precision=args.precision, scan_only=True)  # one scan pass per query

    # ---- sanity: results sorted, rows valid, all ranks agree (full parity lives in tests/)
    s0, r0_ = step_device(0)
    torch.cuda.synchronize()
    s_chk, r_chk = s0.cpu().numpy(), r0_.cpu().numpy()
    assert np.all(np.diff(s_chk, axis=1) <= 0) and r_chk.min() >= 0 and r_chk.max() < args.rows

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- value: device-resident inputs/outputs (clock sampling starts before the warm-up so that
    # short timed regions still get samples taken under load)
    sampler = ClockSampler(local_rank)
    sampler.start()
    for i in range(args.warmup):
        step_device(i)
    barrier()
    # nvidia-smi needs ~1 s before its first sample: when the whole timed region is shorter than that
    # (sharded runs: a few ms per step), keep the GPUs under the same load with extra untimed steps --
    # the same number on every rank, the steps contain a collective
    pe0, pe1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    pe0.record()
    for i in range(3):
        step_device(i)
    pe1.record()
    torch.cuda.synchronize()
    probe = torch.tensor([pe0.elapsed_time(pe1) / 3.0], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(probe, op=dist.ReduceOp.MAX)
    step_ms = max(float(probe.item()), 1e-3)
    if step_ms * args.steps < 1500.0:
        for i in range(min(20000, int((1500.0 - step_ms * args.steps) / step_ms) + 1)):
            step_device(i)
    barrier()
    launches0 = N.kernel_launches()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for i in range(args.steps):
        step_device(args.warmup + i)
    ev1.record()
    barrier()
    launches = N.kernel_launches() - launches0
    dev_ms = ev0.elapsed_time(ev1)

    # ---- e2e: public host API, ONE query per call, pinned host query, result back on the host
    def step_e2e(i):
        base = (i % (n_pool // qps)) * qps
        out = None
        for j in range(qps):
            out = sharded.search(q_np[base + j: base + j + 1], k, precision=args.precision) if world > 1 else \
                store.search(q_np[base + j: base + j + 1], k, precision=args.precision)
        return out

    for i in range(min(args.warmup, 3)):
        step_e2e(i)
    barrier()
    t0 = time.perf_counter()
    for i in range(args.steps):
        step_e2e(args.warmup + i)
    barrier()
    e2e_s = time.perf_counter() - t0
    clocks = sampler.stop()

    # ---- roofline of the dominant kernel: scan launches only (pre-normalised query)
    n_qn = min(64, n_pool)
    qn_dev = torch.nn.functional.normalize(q_dev[:n_qn], dim=1).contiguous()
    out_s = torch.empty(k, dtype=torch.float32, device=dev)
    out_r = torch.empty(k, dtype=torch.int64, device=dev)
    n_scan = 50
    for j in range(5):
        store.search_dev(qn_dev[j].data_ptr(), 1, k, out_s.data_ptr(), out_r.data_ptr(), precision=args.precision,
                         normalized=True, stream=stream)
    torch.cuda.synchronize()
    l0 = N.kernel_launches()
    ev0.record()
    for j in range(n_scan):
        store.search_dev(qn_dev[j % n_qn].data_ptr(), 1, k, out_s.data_ptr(), out_r.data_ptr(),
                         precision=args.precision, normalized=True, stream=stream)
    ev1.record()
    torch.cuda.synchronize()
    assert N.kernel_launches() - l0 == n_scan, "roofline loop must launch exactly one kernel per query"
    scan_ms = ev0.elapsed_time(ev1) / n_scan

    times = torch.tensor([dev_ms, e2e_s * 1e3, scan_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(times, op=dist.ReduceOp.MAX)
    dev_ms, e2e_ms, scan_ms = [float(x) for x in times.cpu()]

    if rank == 0:
        n_queries = args.steps * qps
        value = n_queries / (dev_ms / 1e3)
        e2e_value = n_queries / (e2e_ms / 1e3)
        peak, peak_src = load_peaks()
        algo_bytes = n_local * args.dim * args.elem_bytes + n_local / 8
        achieved = algo_bytes / (scan_ms / 1e3) / 1e9
        line = {
            "metric": METRIC,
            "value": value,
            "unit": UNIT,
            "n_gpus": world,
            "steps": args.steps,
            "warmup": args.warmup,
            "ms_per_step": dev_ms / args.steps,
            "higher_is_better": True,
            "scaling": "strong",
            "vs_baseline": None,
            "dtype": args.store_dtype,
            "data": "synthetic",
            "config": workload_config(args, world),
            "e2e": {
                "value": e2e_value,
                "unit": UNIT,
                "h2d_bytes_per_step": qps * args.dim * 4,
                "d2h_bytes_per_step": qps * k * 12,
                "api": "DeviceStore.search -> pvdb_search (one query per call; pinned host query; blocking)",
            },
            "gpu_launches": int(launches),
            "roofline": {
                "bound": "hbm",
                "kernel": f"scan_topk_kernel<{args.store_dtype}> (masked GEMV + fused top-k)",
                "achieved": achieved,
                "peak": peak,
                "unit": "GB/s",
                "frac": achieved / peak,
                "peak_source": peak_src,
                "algorithmic_bytes_per_launch": algo_bytes,
                "us_per_launch": scan_ms * 1e3,
                "traffic": load_traffic(f"scan_{args.store_dtype}_{args.rows}x{args.dim}") if world == 1 else None,
            },
            "clocks": clocks,
        }
        if world == 1 and not args.no_cpu_baseline:
            per_query, sample, _ = time_oracle(args, args.cpu_queries)
            line["cpu_baseline"] = {"value": 1.0 / per_query, "unit": UNIT, "cores": blas_threads(), "kind": "port",
                                    "sample": sample}
        emit_json(line)
    store.close()
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
