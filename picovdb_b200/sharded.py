"""Row-sharded search across the GPUs of one box: one process per GPU, ``torch.distributed``.

SURVEY.md 8(e): the database rows are split into contiguous, equal blocks (global row r lives on
rank ``r // rows_per_rank``); every rank holds the full query batch, scans its own shard with the
same kernels as the single-GPU path and produces a local (Q, k) list whose row indices are already
global (``row_base``); the per-rank lists are exchanged and k-way merged, after which every rank
holds the same final (Q, k).  The reference has no counterpart (it is single process); the merged
result equals what one store holding all rows would return.

The exchange (k <= 128): every rank owns a mailbox in HBM that its peers write over NVLink (CUDA IPC);
the kernel that finishes a local list stores it into every peer's mailbox, raises a flag, waits for
the peers' flags and merges -- inside the scan kernel for single queries, one extra kernel after a
tensor-core batch (csrc/exchange.cuh).  No collective call and no extra launch sits on the query
path; torch.distributed only carries the one-time handle exchange.  If the mailboxes cannot be set
up (no peer access), or for k > 128, the round-1 path is used: ONE NCCL all-gather of the packed
{rows, scores} blocks + the merge kernel.

torch is used here for what it is good at: process-group plumbing, device buffers, streams.
"""
from __future__ import annotations

import os
from typing import Callable, Optional

import numpy as np
import torch
import torch.distributed as dist


def shard_range(total_rows: int, world_size: int, rank: int) -> tuple[int, int]:
    """Contiguous block of rows owned by ``rank``: [row0, row1).  Blocks are ceil(total/world)
    rows, rounded up to a multiple of 32 so every shard's bitmap starts on a word boundary."""
    per = -(-total_rows // world_size)
    per = (per + 31) // 32 * 32
    row0 = min(total_rows, rank * per)
    return row0, min(total_rows, row0 + per)


def owner_of(row: int, total_rows: int, world_size: int) -> int:
    per = -(-total_rows // world_size)
    per = (per + 31) // 32 * 32
    return min(world_size - 1, row // per)


def packed_result_bytes(nq: int, k: int) -> int:
    """Bytes of one rank's packed result block: nq*k int64 rows, then nq*k fp32 scores, padded to
    16 bytes so consecutive blocks keep both arrays aligned."""
    return (nq * k * 12 + 15) // 16 * 16


class ShardedSearch:
    """Search over a row-sharded store.  Every rank constructs one around its local shard.

    ``local``   engine with ``search`` / ``search_dev`` / ``set_row_base`` (a ``DeviceStore``)
    ``row0``    global index of the shard's first row
    ``merge``   test hook only: a host merge function for process groups without a GPU (gloo);
                with CUDA tensors the merge always runs in the library's kernel.
    """

    def __init__(self, local, row0: int, group=None, merge: Optional[Callable] = None) -> None:
        self.local = local
        self.row0 = int(row0)
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self._host_merge = merge
        local.set_row_base(self.row0)
        self._bufs: dict[tuple, dict[str, torch.Tensor]] = {}
        self._ex = None            # engine.Exchange once connected
        self._ex_failed = False    # peer mailboxes unavailable on this box: stay on the NCCL path

    # ------------------------------------------------------------------ peer-memory exchange
    MAX_EXCHANGE_K = 128

    def _exchange(self, nq: int, k: int):
        """The connected exchange end sized for nq x k results, or None (use NCCL).  Collective:
        every rank calls it with the same arguments, so (re)creation happens in lockstep."""
        if (self.world == 1 or k > self.MAX_EXCHANGE_K or self._ex_failed or self._host_merge is not None
                or not torch.cuda.is_available() or os.environ.get("PVDB_NO_PEER_EXCHANGE")):
            return None
        need = nq * k
        if self._ex is not None and self._ex.slot_keys >= need:
            return self._ex
        from . import _native as N
        from .engine import Exchange

        dev = torch.device("cuda", torch.cuda.current_device())
        if self._ex is not None:      # outgrown: nobody may still be inside a kernel that uses it
            self.close()
        ok, ex = 1, None
        handle = bytes(N.IPC_HANDLE_BYTES)
        try:
            ex = Exchange(dev.index, self.world, self.rank, max(need, 1 << 16))
            handle = ex.ipc_handle()
        except Exception:
            ok = 0
        mine = torch.frombuffer(bytearray(handle), dtype=torch.uint8).to(dev)
        everyone = torch.empty(self.world * N.IPC_HANDLE_BYTES, dtype=torch.uint8, device=dev)
        dist.all_gather_into_tensor(everyone, mine, group=self.group)
        flag = torch.tensor([ok], dtype=torch.int32, device=dev)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=self.group)
        if int(flag.item()) == 1:
            try:
                ex.connect_ipc(everyone.cpu().numpy().tobytes())
            except Exception:
                ok = 0
            flag = torch.tensor([ok], dtype=torch.int32, device=dev)
            dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=self.group)
        if int(flag.item()) != 1:
            if ex is not None:
                ex.close()
            self._ex_failed = True
            return None
        torch.cuda.synchronize()
        dist.barrier(group=self.group)   # every rank's mailbox is mapped everywhere before the first launch
        self._ex = ex
        return ex

    @property
    def exchange_mode(self) -> str:
        if self.world == 1:
            return "none"
        return "peer-memory mailboxes (fused)" if self._ex is not None else "nccl all-gather + merge kernel"

    # ------------------------------------------------------------------ device path
    def _buffers(self, nq: int, k: int, device) -> dict[str, torch.Tensor]:
        key = (nq, k)
        b = self._bufs.get(key)
        if b is None:
            nbytes = packed_result_bytes(nq, k)
            # merged result: one packed block {rows int64 | scores fp32} so the host path needs ONE copy
            out = torch.empty(nbytes, dtype=torch.uint8, device=device)
            n_out = nq * k
            b = {
                "local": torch.empty(nbytes, dtype=torch.uint8, device=device),
                "all": torch.empty(self.world * nbytes, dtype=torch.uint8, device=device),
                "out": out,
                "rows": out[: n_out * 8].view(torch.int64).view(nq, k),
                "scores": out[n_out * 8 : n_out * 12].view(torch.float32).view(nq, k),
            }
            self._bufs[key] = b
        return b

    def search_dev(self, d_queries: torch.Tensor, k: int, precision: str = "auto", normalized: bool = False,
                   d_prefilter: Optional[torch.Tensor] = None, scan_only: bool = False) -> tuple[torch.Tensor, torch.Tensor]:
        """CUDA tensors in (``(Q, dim)`` fp32), CUDA tensors out; enqueued on torch's current stream.
        The returned tensors are reused by the next call with the same (Q, k)."""
        from .engine import merge_topk_dev

        nq = d_queries.shape[0]
        b = self._buffers(nq, k, d_queries.device)
        stream = torch.cuda.current_stream().cuda_stream
        n_out = nq * k
        if self.world == 1:
            # nothing to merge: the shard's result IS the result (rows already carry row_base)
            self.local.search_dev(
                d_queries.data_ptr(), nq, k, b["scores"].data_ptr(), b["rows"].data_ptr(),
                d_prefilter=d_prefilter.data_ptr() if d_prefilter is not None else 0,
                precision=precision, normalized=normalized, stream=stream, scan_only=scan_only,
            )
            return b["scores"], b["rows"]
        ex = self._exchange(nq, k)
        if ex is not None:
            # fused: the kernels exchange the lists over peer memory and write the merged result
            self.local.search_exchange_dev(
                ex, d_queries.data_ptr(), nq, k, b["scores"].data_ptr(), b["rows"].data_ptr(),
                d_prefilter=d_prefilter.data_ptr() if d_prefilter is not None else 0,
                precision=precision, normalized=normalized, stream=stream, scan_only=scan_only,
            )
            return b["scores"], b["rows"]
        loc = b["local"]
        # packed block: rows (int64) first, then scores (fp32)
        self.local.search_dev(
            d_queries.data_ptr(), nq, k, loc.data_ptr() + n_out * 8, loc.data_ptr(),
            d_prefilter=d_prefilter.data_ptr() if d_prefilter is not None else 0,
            precision=precision, normalized=normalized, stream=stream, scan_only=scan_only,
        )
        if self.world == 1:
            gathered = loc
        else:
            dist.all_gather_into_tensor(b["all"], loc, group=self.group)
            gathered = b["all"]
        block = packed_result_bytes(nq, k)
        merge_topk_dev(
            d_queries.device.index or 0, gathered.data_ptr() + n_out * 8, gathered.data_ptr(), self.world, nq, k,
            b["scores"].data_ptr(), b["rows"].data_ptr(), stream=stream,
            scores_stride=block // 4, rows_stride=block // 8,
        )
        return b["scores"], b["rows"]

    # ------------------------------------------------------------------ host-buffer path
    def search(self, queries: np.ndarray, k: int, prefilter: Optional[np.ndarray] = None,
               precision: str = "auto") -> tuple[np.ndarray, np.ndarray]:
        """numpy in / numpy out.  ``prefilter`` is this rank's slice of the row mask."""
        q = np.ascontiguousarray(queries, dtype=np.float32)
        if self._host_merge is None and torch.cuda.is_available():
            nq = q.shape[0]
            ex = self._exchange(nq, k)
            if ex is not None:
                # one C-ABI call: H2D, scan + fused exchange + merge, result written to pinned host memory
                return self.local.search_exchange(ex, q, k, prefilter=prefilter, precision=precision)
            st = self._staging(nq, k, q.shape[1])
            st["hq"].numpy()[...] = q                       # pinned staging -> async H2D
            st["dq"].copy_(st["hq"], non_blocking=True)
            dp = None
            if prefilter is not None:
                from .engine import pack_row_mask

                dp = torch.from_numpy(pack_row_mask(prefilter).view(np.int32)).cuda(non_blocking=True)
            self.search_dev(st["dq"], k, precision=precision, d_prefilter=dp)
            b = self._buffers(nq, k, st["dq"].device)
            st["hout"].copy_(b["out"], non_blocking=True)   # one packed D2H: rows then scores
            torch.cuda.current_stream().synchronize()
            raw = st["hout"].numpy()
            n_out = nq * k
            rows = raw[: n_out * 8].view("<i8").reshape(nq, k).copy()
            scores = raw[n_out * 8 : n_out * 12].view("<f4").reshape(nq, k).copy()
            return scores, rows
        # process group without GPUs (tests): same packing / gather / merge flow on host tensors
        if self._host_merge is None:
            raise RuntimeError("ShardedSearch needs CUDA (no CPU compute path)")
        s_loc, r_loc = self.local.search(q, k, prefilter, precision=precision)
        r_loc = np.where(r_loc >= 0, r_loc + self._host_row_base(), r_loc)
        return self.merge_local(s_loc, r_loc, k)

    def merge_local(self, s_loc: np.ndarray, r_loc: np.ndarray, k: int) -> tuple[np.ndarray, np.ndarray]:
        """All-gather + merge of per-rank results that are already on the host (rows global, -1 padded):
        the exchange step for searches that return through a host call (``search_where``)."""
        nq = s_loc.shape[0]
        n_out = nq * k
        if self._host_merge is None and torch.cuda.is_available():
            from .engine import merge_topk_dev

            dev = torch.device("cuda", torch.cuda.current_device())
            b = self._buffers(nq, k, dev)
            key = ("merge", nq, k)
            st = self._bufs.get(key)
            if st is None:
                st = {"hloc": torch.empty(packed_result_bytes(nq, k), dtype=torch.uint8).pin_memory(),
                      "hout": torch.empty(packed_result_bytes(nq, k), dtype=torch.uint8).pin_memory()}
                self._bufs[key] = st
            raw = st["hloc"].numpy()
            raw[: n_out * 8] = np.ascontiguousarray(r_loc, dtype="<i8").view(np.uint8).ravel()
            raw[n_out * 8: n_out * 12] = np.ascontiguousarray(s_loc, dtype="<f4").view(np.uint8).ravel()
            b["local"].copy_(st["hloc"], non_blocking=True)
            if self.world == 1:
                gathered = b["local"]
            else:
                dist.all_gather_into_tensor(b["all"], b["local"], group=self.group)
                gathered = b["all"]
            block = packed_result_bytes(nq, k)
            merge_topk_dev(dev.index or 0, gathered.data_ptr() + n_out * 8, gathered.data_ptr(), self.world, nq, k,
                           b["scores"].data_ptr(), b["rows"].data_ptr(), stream=torch.cuda.current_stream().cuda_stream,
                           scores_stride=block // 4, rows_stride=block // 8)
            st["hout"].copy_(b["out"], non_blocking=True)
            torch.cuda.current_stream().synchronize()
            out = st["hout"].numpy()
            return (out[n_out * 8: n_out * 12].view("<f4").reshape(nq, k).copy(),
                    out[: n_out * 8].view("<i8").reshape(nq, k).copy())
        if self._host_merge is None:
            raise RuntimeError("ShardedSearch needs CUDA (no CPU compute path)")
        packed = np.zeros(packed_result_bytes(nq, k), dtype=np.uint8)
        packed[: nq * k * 8] = r_loc.astype("<i8").view(np.uint8).ravel()
        packed[nq * k * 8 : nq * k * 12] = s_loc.astype("<f4").view(np.uint8).ravel()
        loc = torch.from_numpy(packed)
        if self.world == 1:
            gathered = loc
        else:
            gathered = torch.empty(self.world * loc.numel(), dtype=torch.uint8)
            dist.all_gather_into_tensor(gathered, loc, group=self.group)
        blocks = gathered.numpy().reshape(self.world, -1)
        rows = [blk[: n_out * 8].view("<i8").reshape(nq, k) for blk in blocks]
        scores = [blk[n_out * 8 : n_out * 12].view("<f4").reshape(nq, k) for blk in blocks]
        return self._host_merge(scores, rows, k)

    def close(self) -> None:
        """Collective when an exchange is connected: unmap the peers' mailboxes, barrier, free."""
        if self._ex is not None:
            torch.cuda.synchronize()
            self._ex.disconnect()
            dist.barrier(group=self.group)
            torch.cuda.synchronize()
            self._ex.close()
            self._ex = None

    def _staging(self, nq: int, k: int, dim: int) -> dict[str, torch.Tensor]:
        """Pinned host + device staging buffers for the host-buffer path, cached per (nq, k)."""
        key = ("staging", nq, k)
        st = self._bufs.get(key)
        if st is None:
            dev = torch.device("cuda", torch.cuda.current_device())
            st = {
                "hq": torch.empty((nq, dim), dtype=torch.float32).pin_memory(),
                "dq": torch.empty((nq, dim), dtype=torch.float32, device=dev),
                "hout": torch.empty(packed_result_bytes(nq, k), dtype=torch.uint8).pin_memory(),
            }
            self._bufs[key] = st
        return st

    def _host_row_base(self) -> int:
        # a test engine that does not implement row_base itself gets the offset added here
        return 0 if getattr(self.local, "applies_row_base", False) else self.row0


class ShardedStore:
    """A store whose rows are split over the ranks of a process group, with ``DeviceStore``'s
    interface -- so ``PicoVectorDB`` can sit on top of it unchanged (SPMD: every rank makes the same
    calls with the same arguments and keeps the same ids / documents; only the vectors are sharded).

    SURVEY.md 8(e): global row r lives on rank ``r // rows_per_rank`` (contiguous blocks, so a saved
    matrix is the concatenation of the shards); upserts and deletes touch the owning shard only;
    a search is every rank's local scan + one all-gather + the merge kernel (``ShardedSearch``).
    The partition needs a fixed row capacity (``reserve_rows`` / PicoVectorDB's ``capacity=``).
    Most methods contain a collective: call them in the same order on every rank, from one thread.

    ``local_factory`` / ``merge`` are test hooks (gloo process groups without GPUs).
    """

    def __init__(self, dim: int, device: Optional[int] = None, reserve_rows: int = 0, keep_f32: bool = True,
                 bf16_mirror: bool = False, fixed_capacity: bool = False, *, group=None,
                 local_factory: Optional[Callable] = None, merge: Optional[Callable] = None) -> None:
        if reserve_rows <= 0:
            raise ValueError("ShardedStore needs the total row capacity (reserve_rows / capacity=) to place rows")
        self.dim = int(dim)
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.capacity = int(reserve_rows)
        self.row0, self.row1 = shard_range(self.capacity, self.world, self.rank)
        self.per = shard_range(self.capacity, self.world, 0)[1]
        if device is None:
            device = int(torch.cuda.current_device()) if torch.cuda.is_available() else 0
        if local_factory is None:
            from .engine import DeviceStore as local_factory  # noqa: N813
        self.local = local_factory(self.dim, device=device, reserve_rows=max(self.row1 - self.row0, 1),
                                   keep_f32=keep_f32, bf16_mirror=bf16_mirror, fixed_capacity=fixed_capacity)
        self.bf16_only = bool(getattr(self.local, "bf16_only", False))
        self._search = ShardedSearch(self.local, self.row0, group=group, merge=merge)
        self._host_only = merge is not None
        self._rows = 0  # global high-water mark (the same on every rank)
        self._cols: set[int] = set()  # metadata columns this rank holds codes for

    # ------------------------------------------------------------------ plumbing
    @property
    def rows(self) -> int:
        return self._rows

    @property
    def is_writer(self) -> bool:
        """True on the one rank that writes shared files (ids, documents, the .npy header)."""
        return self.rank == 0

    def owned_rows(self) -> tuple[int, int]:
        return self.row0, self.row1

    def free_order(self, cap: int) -> list[int]:
        """Free-slot list for PicoVectorDB (it pops from the END): rows are handed out round-robin
        over the shards -- shard 0's first row, shard 1's first row, ... -- so a store that is only
        partly filled still spreads its rows (and the scan work) evenly."""
        seq = [(j % self.world) * self.per + j // self.world for j in range(self.per * self.world)]
        seq = [r for r in seq if r < cap]
        return seq[::-1]

    def barrier(self) -> None:
        """Host-level barrier (used around shared files): an NCCL barrier is only stream ordered, so
        the host also waits for it."""
        if self.world > 1:
            dist.barrier(group=self.group)
            if not self._host_only and torch.cuda.is_available():
                torch.cuda.synchronize()

    def _sum_over_ranks(self, arr: np.ndarray) -> np.ndarray:
        """Element-wise sum of a same-shaped array over the ranks (every rank gets the result)."""
        if self.world == 1:
            return arr
        t = torch.from_numpy(np.ascontiguousarray(arr))
        if not self._host_only and torch.cuda.is_available():
            t = t.cuda()
        dist.all_reduce(t, group=self.group)
        return t.cpu().numpy()

    def _mine(self, rows: np.ndarray) -> np.ndarray:
        return (rows >= self.row0) & (rows < self.row1)

    # ------------------------------------------------------------------ write side
    def reserve(self, rows: int) -> None:
        if rows > self.capacity:
            raise ValueError(f"ShardedStore capacity is fixed at {self.capacity} rows")

    def upsert_rows(self, vecs: np.ndarray, rows: np.ndarray) -> None:
        rows = np.asarray(rows, dtype=np.int64)
        if rows.size and (rows.min() < 0 or rows.max() >= self.capacity):
            raise ValueError("row outside the sharded store's capacity")
        m = self._mine(rows)
        if m.any():
            self.local.upsert_rows(np.asarray(vecs)[m], rows[m] - self.row0)
        if rows.size:
            self._rows = max(self._rows, int(rows.max()) + 1)

    def upsert_range(self, vecs: np.ndarray, row0: int) -> None:
        n = len(vecs)
        if row0 < 0 or row0 + n > self.capacity:
            raise ValueError("row range outside the sharded store's capacity")
        lo, hi = max(row0, self.row0), min(row0 + n, self.row1)
        if hi > lo:
            self.local.upsert_range(np.asarray(vecs)[lo - row0: hi - row0], lo - self.row0)
        self._rows = max(self._rows, row0 + n)

    def delete_rows(self, rows: np.ndarray) -> None:
        rows = np.asarray(rows, dtype=np.int64)
        m = self._mine(rows)
        if m.any():
            self.local.delete_rows(rows[m] - self.row0)

    def upload(self, vecs: np.ndarray, row0: int = 0, active: Optional[np.ndarray] = None) -> None:
        """Raw load of already-normalised rows [row0, row0 + len(vecs)); each rank keeps its part."""
        n = len(vecs)
        lo, hi = max(row0, self.row0), min(row0 + n, self.row1)
        if hi > lo:
            a = None if active is None else np.asarray(active, dtype=bool)[lo - row0: hi - row0]
            self.local.upload(np.ascontiguousarray(vecs[lo - row0: hi - row0], dtype=np.float32), lo - self.row0, a)
        self._rows = max(self._rows, row0 + n)

    def upload_bf16(self, vecs16: np.ndarray, row0: int = 0, active: Optional[np.ndarray] = None) -> None:
        n = len(vecs16)
        lo, hi = max(row0, self.row0), min(row0 + n, self.row1)
        if hi > lo:
            a = None if active is None else np.asarray(active, dtype=bool)[lo - row0: hi - row0]
            self.local.upload_bf16(np.ascontiguousarray(vecs16[lo - row0: hi - row0]), lo - self.row0, a)
        self._rows = max(self._rows, row0 + n)

    def download_bf16(self, row0: int = 0, n: Optional[int] = None) -> np.ndarray:
        """This rank's part of rows [row0, row0 + n) of the mirror; rows of other ranks read as zeros
        (save() asks every rank for its own rows only)."""
        if n is None:
            n = self._rows - row0
        out = np.zeros((n, self.dim), dtype=np.uint16)
        lo, hi = max(row0, self.row0), min(row0 + n, self.row1, self.row0 + self.local.rows)
        if hi > lo:
            out[lo - row0: hi - row0] = self.local.download_bf16(lo - self.row0, hi - lo)
        return out

    def write_file(self, path: str, file_offset: int, row0: int, n: int, as_bf16: bool = False) -> None:
        """This rank's written rows of [row0, row0 + n) into the file (``file_offset`` = where row0 goes)."""
        lo, hi = max(row0, self.row0), min(row0 + n, self.row1, self.row0 + self.local.rows)
        if hi > lo and hasattr(self.local, "write_file"):
            row_bytes = self.dim * (2 if as_bf16 else 4)
            self.local.write_file(path, file_offset + (lo - row0) * row_bytes, lo - self.row0, hi - lo, as_bf16)
        elif hi > lo:  # test engines: through a map
            mm = np.memmap(path, dtype=np.uint16 if as_bf16 else np.float32, mode="r+", offset=file_offset,
                           shape=(n, self.dim))
            mm[lo - row0: hi - row0] = (self.local.download_bf16 if as_bf16 else self.local.download)(lo - self.row0, hi - lo)
            mm.flush()

    def compact(self, keep_rows: np.ndarray) -> None:
        """Global compaction: new row j <- old row keep[j] (keep ascending, so keep[j] >= j and moving
        block by block in ascending order never overwrites a row that is still to be read)."""
        keep = np.asarray(keep_rows, dtype=np.int64)
        step = max(32, ((32 << 20) // (self.dim * 4)) // 32 * 32)
        for a in range(0, keep.size, step):
            b = min(keep.size, a + step)
            self.upload(self.fetch_rows(keep[a:b]), a, np.ones(b - a, dtype=bool))
        n_local = int(np.clip(keep.size - self.row0, 0, self.row1 - self.row0))
        self.local.compact(np.arange(n_local, dtype=np.int64))  # drops the shard's rows past the new end
        self._cols.clear()                                      # ... and its metadata columns
        self._rows = int(keep.size)

    # ------------------------------------------------------------------ read side
    def fetch_rows(self, rows: np.ndarray) -> np.ndarray:
        rows = np.asarray(rows, dtype=np.int64)
        out = np.zeros((rows.size, self.dim), dtype=np.float32)
        m = self._mine(rows) & (rows - self.row0 < self.local.rows)  # rows never written read as zeros
        if m.any():
            out[m] = self.local.fetch_rows(rows[m] - self.row0)
        return self._sum_over_ranks(out)

    def download(self, row0: int = 0, n: Optional[int] = None) -> np.ndarray:
        if n is None:
            n = self._rows - row0
        lo, hi = max(row0, self.row0), min(row0 + n, self.row1)
        out = np.zeros((n, self.dim), dtype=np.float32)
        used = min(hi, self.row0 + self.local.rows)  # the shard's rows past its high-water mark are zeros
        if used > lo:
            out[lo - row0: used - row0] = self.local.download(lo - self.row0, used - lo)
        if lo == row0 and hi == row0 + n:
            return out  # entirely this rank's rows: no collective
        return self._sum_over_ranks(out)

    def active_mask(self) -> np.ndarray:
        out = np.zeros(self._rows, dtype=np.int32)
        hi = min(self._rows, self.row1)
        if hi > self.row0:
            out[self.row0: hi] = np.asarray(self.local.active_mask(), dtype=bool)[: hi - self.row0]
        return self._sum_over_ranks(out).astype(bool)

    def search(self, queries: np.ndarray, k: int, prefilter: Optional[np.ndarray] = None, precision: str = "auto",
               normalized: bool = False, **_) -> tuple[np.ndarray, np.ndarray]:
        """``prefilter`` is the GLOBAL row mask (length >= rows); each rank applies its slice."""
        pf = None
        if prefilter is not None:
            pf = np.zeros(max(self.row1 - self.row0, 0), dtype=bool)
            hi = min(len(prefilter), self.row1)
            if hi > self.row0:
                pf[: hi - self.row0] = np.asarray(prefilter, dtype=bool)[self.row0: hi]
        return self._search.search(queries, k, prefilter=pf, precision=precision)

    # ------------------------------------------------------------------ metadata columns / dict filters
    MAX_COLUMNS = 16

    def column_write(self, column: int, codes: np.ndarray, rows: Optional[np.ndarray] = None, row0: int = 0) -> None:
        """Codes of the given global rows (or of the consecutive rows from row0); each rank keeps the
        codes of the rows it owns."""
        codes = np.asarray(codes, dtype=np.int32)
        if rows is not None:
            rows = np.asarray(rows, dtype=np.int64)
            m = self._mine(rows) & (rows - self.row0 < self.local.rows)
            if m.any():
                self.local.column_write(column, codes[m], rows=rows[m] - self.row0)
                self._cols.add(column)
            return
        lo, hi = max(row0, self.row0), min(row0 + codes.size, self.row0 + self.local.rows)
        if hi > lo:
            self.local.column_write(column, codes[lo - row0: hi - row0], row0=lo - self.row0)
            self._cols.add(column)

    def column_drop(self, column: int) -> None:
        if column in self._cols:
            self.local.column_drop(column)
            self._cols.discard(column)

    def search_where(self, queries: np.ndarray, k: int, column: int, wanted_codes, extra: Optional[np.ndarray] = None,
                     precision: str = "auto") -> tuple[np.ndarray, np.ndarray, int]:
        """Every rank filters and scans its own rows on its GPU; results are merged as in ``search`` and
        the candidate counts are summed.  ``extra`` is the GLOBAL row mask of an ``ids=`` restriction."""
        q = np.ascontiguousarray(queries, dtype=np.float32)
        nq = q.shape[0]
        if column in self._cols and self.local.rows > 0:
            ex = None
            if extra is not None:
                ex = np.zeros(self.local.rows, dtype=bool)
                hi = min(len(extra), self.row0 + self.local.rows)
                if hi > self.row0:
                    ex[: hi - self.row0] = np.asarray(extra, dtype=bool)[self.row0: hi]
            s_loc, r_loc, cand = self.local.search_where(q, k, column, wanted_codes, ex, precision=precision)
            if not getattr(self.local, "applies_row_base", False):
                r_loc = np.where(r_loc >= 0, r_loc + self.row0, r_loc)
        else:  # this rank holds no row of the column
            s_loc = np.full((nq, k), -np.inf, dtype=np.float32)
            r_loc = np.full((nq, k), -1, dtype=np.int64)
            cand = 0
        scores, rows = self._search.merge_local(s_loc, r_loc, k)
        total = int(self._sum_over_ranks(np.array([cand], dtype=np.int64))[0])
        return scores, rows, total

    def close(self) -> None:
        self._search.close()
        self.local.close()


def _sharded_engine_factory(dim: int, **kw):
    return ShardedStore(dim, **kw)


def _make_sharded_db_class():
    from .db import PicoVectorDB

    class ShardedPicoVectorDB(PicoVectorDB):
        """``PicoVectorDB`` whose vectors are row-sharded over the ranks of the default process group.

        SPMD: construct it and call it identically on every rank (one process per GPU under
        ``torchrun``, ``device=LOCAL_RANK``); ``capacity=`` is required.  ids and documents are
        replicated, vectors / scans / dict filters are per shard, ``save()`` writes ONE store in the
        reference's file format (rank 0: ids + documents, every rank: its rows of the matrix).
        """

        _engine_factory = staticmethod(_sharded_engine_factory)

    return ShardedPicoVectorDB


def __getattr__(name: str):
    # lazily built so importing this module never imports db.py (and the CUDA library) by itself
    if name == "ShardedPicoVectorDB":
        cls = _make_sharded_db_class()
        globals()[name] = cls
        return cls
    raise AttributeError(name)
