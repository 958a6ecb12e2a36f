"""``GroupStore``: a row-sharded store over several GPUs of ONE process (``pvdb_group_t``).

This is the ``devices=[...]`` form of SURVEY.md 8(b): a library user with a single Python process
gets the row partition of 8(e) without ``torchrun``.  The handle owns one shard store and one
peer-memory exchange end per device (csrc/group.cu, csrc/exchange.cuh); a search is ONE C call --
every shard scans its rows on its own GPU, the per-GPU top-k lists travel over NVLink inside the
kernels, shard 0's copy of the merged result comes back.  Writes are routed to the owning shard's
store handle here (contiguous blocks of rows, the same partition as ``sharded.shard_range``).

The class has ``DeviceStore``'s interface, so ``PicoVectorDB(devices=[...], capacity=N)`` sits on top
of it unchanged.  Dict ``where`` filters take the host column index (db.py) and arrive as a row mask.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional, Sequence

import numpy as np

from . import _native as N
from .engine import DeviceStore, _f32c, _ptr, pack_row_mask


class GroupStore:
    applies_row_base = True  # result rows are global already

    def __init__(self, dim: int, devices: Sequence[int], reserve_rows: int = 0, keep_f32: bool = True,
                 bf16_mirror: bool = False, fixed_capacity: bool = False, slot_keys: int = 0, **_) -> None:
        if reserve_rows <= 0:
            raise ValueError("a device group needs the total row capacity (capacity=) to place rows")
        devices = [int(d) for d in devices]
        if len(set(devices)) != len(devices) or not devices:
            raise ValueError("devices must be a non-empty list of distinct CUDA ordinals")
        self._lib = N.load()
        flags = (N.STORE_F32 if keep_f32 else 0) | (N.STORE_BF16 if bf16_mirror else 0)
        if fixed_capacity:
            flags |= N.STORE_FIXED_CAPACITY
        h = C.c_void_p()
        arr = (C.c_int * len(devices))(*devices)
        N.check(self._lib.pvdb_group_create(C.byref(h), arr, len(devices), int(dim), int(reserve_rows), flags,
                                            int(slot_keys)))
        self._h = h
        self.dim = int(dim)
        self.devices = devices
        self.world = len(devices)
        self.capacity = int(reserve_rows)
        per = C.c_int64(0)
        N.check(self._lib.pvdb_group_size(h, None, C.byref(per)))
        self.per = int(per.value)
        self.slot_keys = int(slot_keys) if slot_keys > 0 else 65536
        self.shards = [DeviceStore.from_handle(C.c_void_p(self._lib.pvdb_group_store(h, i)), dim, d)
                       for i, d in enumerate(devices)]
        self._rows = 0
        self.bf16_only = bool(bf16_mirror and not keep_f32)

    # ------------------------------------------------------------------ plumbing
    @property
    def handle(self) -> C.c_void_p:
        if self._h is None:
            raise RuntimeError("GroupStore is closed")
        return self._h

    def close(self) -> None:
        h, self._h = getattr(self, "_h", None), None
        if h is not None and h.value:
            self._lib.pvdb_group_destroy(h)

    def __del__(self) -> None:  # pragma: no cover
        try:
            self.close()
        except Exception:
            pass

    @property
    def rows(self) -> int:
        return self._rows

    def shard_bounds(self, i: int) -> tuple[int, int]:
        lo = min(self.capacity, i * self.per)
        return lo, min(self.capacity, lo + self.per)

    def free_order(self, cap: int) -> list[int]:
        """Free-slot list for PicoVectorDB (popped from the END): rows are dealt round-robin over the
        shards so a partly filled store keeps every GPU equally busy."""
        seq = [(j % self.world) * self.per + j // self.world for j in range(self.per * self.world)]
        return [r for r in seq if r < cap][::-1]

    def set_row_base(self, base: int) -> None:
        if base:
            raise ValueError("a device group is the whole database: row_base stays 0")

    def reserve(self, rows: int) -> None:
        if rows > self.capacity:
            raise ValueError(f"the group's capacity is fixed at {self.capacity} rows")

    def _split(self, rows: np.ndarray):
        owner = np.minimum(rows // self.per, self.world - 1)
        for i in range(self.world):
            m = owner == i
            if m.any():
                yield i, m

    # ------------------------------------------------------------------ writes
    def upsert_rows(self, vecs: np.ndarray, rows: np.ndarray) -> None:
        rows = np.asarray(rows, dtype=np.int64)
        vecs = _f32c(vecs)
        if rows.size and (rows.min() < 0 or rows.max() >= self.capacity):
            raise ValueError("row outside the group's capacity")
        for i, m in self._split(rows):
            self.shards[i].upsert_rows(vecs[m], rows[m] - i * self.per)
        if rows.size:
            self._rows = max(self._rows, int(rows.max()) + 1)

    def upsert_range(self, vecs: np.ndarray, row0: int) -> None:
        n = len(vecs)
        if row0 < 0 or row0 + n > self.capacity:
            raise ValueError("row range outside the group's capacity")
        for i in range(self.world):
            lo, hi = self.shard_bounds(i)
            a, b = max(row0, lo), min(row0 + n, hi)
            if b > a:
                self.shards[i].upsert_range(vecs[a - row0: b - row0], a - lo)
        self._rows = max(self._rows, row0 + n)

    def delete_rows(self, rows) -> None:
        rows = np.asarray(rows, dtype=np.int64)
        for i, m in self._split(rows):
            self.shards[i].delete_rows(rows[m] - i * self.per)

    def upload(self, vecs: np.ndarray, row0: int = 0, active: Optional[np.ndarray] = None) -> None:
        n = len(vecs)
        for i in range(self.world):
            lo, hi = self.shard_bounds(i)
            a, b = max(row0, lo), min(row0 + n, hi)
            if b > a:
                act = None if active is None else np.asarray(active, dtype=bool)[a - row0: b - row0]
                self.shards[i].upload(np.ascontiguousarray(vecs[a - row0: b - row0], dtype=np.float32), a - lo, act)
        self._rows = max(self._rows, row0 + n)

    def compact(self, keep_rows) -> None:
        """Global compaction: new row j <- old row keep[j] (ascending, so keep[j] >= j: moving block by
        block in ascending order never overwrites a row that is still to be read)."""
        keep = np.asarray(keep_rows, dtype=np.int64)
        step = max(32, ((32 << 20) // (self.dim * 4)) // 32 * 32)
        for a in range(0, keep.size, step):
            b = min(keep.size, a + step)
            self.upload(self.fetch_rows(keep[a:b]), a, np.ones(b - a, dtype=bool))
        for i in range(self.world):
            lo, hi = self.shard_bounds(i)
            n_local = int(np.clip(keep.size - lo, 0, hi - lo))
            self.shards[i].compact(np.arange(n_local, dtype=np.int64))  # drops the shard's rows past the new end
        self._rows = int(keep.size)

    # ------------------------------------------------------------------ reads
    def fetch_rows(self, rows) -> np.ndarray:
        rows = np.asarray(rows, dtype=np.int64)
        out = np.zeros((rows.size, self.dim), dtype=np.float32)
        for i, m in self._split(rows):
            local = rows[m] - i * self.per
            have = local < self.shards[i].rows       # rows never written read as zeros
            if have.any():
                idx = np.flatnonzero(m)[have]
                out[idx] = self.shards[i].fetch_rows(local[have])
        return out

    def download(self, row0: int = 0, n: Optional[int] = None) -> np.ndarray:
        if n is None:
            n = self._rows - row0
        out = np.zeros((n, self.dim), dtype=np.float32)
        for i in range(self.world):
            lo, hi = self.shard_bounds(i)
            a, b = max(row0, lo), min(row0 + n, hi, lo + self.shards[i].rows)
            if b > a:
                out[a - row0: b - row0] = self.shards[i].download(a - lo, b - a)
        return out

    def upload_bf16(self, vecs16: np.ndarray, row0: int = 0, active: Optional[np.ndarray] = None) -> None:
        n = len(vecs16)
        for i in range(self.world):
            lo, hi = self.shard_bounds(i)
            a, b = max(row0, lo), min(row0 + n, hi)
            if b > a:
                act = None if active is None else np.asarray(active, dtype=bool)[a - row0: b - row0]
                self.shards[i].upload_bf16(np.ascontiguousarray(vecs16[a - row0: b - row0]), a - lo, act)
        self._rows = max(self._rows, row0 + n)

    def download_bf16(self, row0: int = 0, n: Optional[int] = None) -> np.ndarray:
        if n is None:
            n = self._rows - row0
        out = np.zeros((n, self.dim), dtype=np.uint16)
        for i in range(self.world):
            lo, hi = self.shard_bounds(i)
            a, b = max(row0, lo), min(row0 + n, hi, lo + self.shards[i].rows)
            if b > a:
                out[a - row0: b - row0] = self.shards[i].download_bf16(a - lo, b - a)
        return out

    def write_file(self, path: str, file_offset: int, row0: int, n: int, as_bf16: bool = False) -> None:
        row_bytes = self.dim * (2 if as_bf16 else 4)
        for i in range(self.world):
            lo, hi = self.shard_bounds(i)
            a, b = max(row0, lo), min(row0 + n, hi, lo + self.shards[i].rows)
            if b > a:
                self.shards[i].write_file(path, file_offset + (a - row0) * row_bytes, a - lo, b - a, as_bf16)

    def active_mask(self) -> np.ndarray:
        out = np.zeros(self._rows, dtype=bool)
        for i in range(self.world):
            lo, hi = self.shard_bounds(i)
            hi = min(hi, self._rows)
            if hi > lo:
                m = np.asarray(self.shards[i].active_mask(), dtype=bool)[: hi - lo]
                out[lo: lo + m.size] = m
        return out

    def search(self, queries: np.ndarray, k: int, prefilter: Optional[np.ndarray] = None, precision: str = "auto",
               normalized: bool = False, rescore: bool = True, scan_only: bool = False,
               guard: bool = True) -> tuple[np.ndarray, np.ndarray]:
        """(Q, dim) queries -> merged (scores, global rows) over all shards; ``prefilter`` is the GLOBAL
        bool row mask."""
        q = _f32c(queries)
        if q.ndim != 2 or q.shape[1] != self.dim:
            raise ValueError(f"search expects (Q, {self.dim}) queries")
        nq, k = q.shape[0], int(k)
        if k < 1:
            raise ValueError("k must be >= 1")
        bits = None
        if prefilter is not None:
            full = np.zeros(self.per * self.world, dtype=bool)   # every shard finds its words in bounds
            pf = np.asarray(prefilter, dtype=bool)[: full.size]
            full[: pf.size] = pf
            bits = pack_row_mask(full)
        flags = self.shards[0]._search_flags(precision, normalized, rescore, scan_only, guard)
        scores = np.empty((nq, k), dtype=np.float32)
        rows = np.empty((nq, k), dtype=np.int64)
        if k > self.MAX_EXCHANGE_K:
            return self._search_large_k(q, k, prefilter, precision, normalized)
        step = max(1, self.slot_keys // k)            # nq * k of one call must fit the exchange slot
        for q0 in range(0, nq, step):
            q1 = min(nq, q0 + step)
            N.check(self._lib.pvdb_group_search(self.handle, _ptr(q[q0:q1]), q1 - q0, k, _ptr(bits), flags,
                                                _ptr(scores[q0:q1]), _ptr(rows[q0:q1])))
        return scores, rows

    MAX_EXCHANGE_K = 128

    def _search_large_k(self, q, k, prefilter, precision, normalized):
        """k beyond the fused exchange (paged scans): per-shard searches, merged on the host by
        (score descending, row ascending) -- the key order of the device merges."""
        parts_s, parts_r = [], []
        for i, sh in enumerate(self.shards):
            lo, hi = self.shard_bounds(i)
            pf = None
            if prefilter is not None:
                pf = np.zeros(max(sh.rows, 1), dtype=bool)
                src = np.asarray(prefilter, dtype=bool)[lo: lo + sh.rows]
                pf[: src.size] = src
            s, r = sh.search(q, k, prefilter=pf, precision=precision, normalized=normalized)
            parts_s.append(s)
            parts_r.append(r)
        s = np.concatenate(parts_s, axis=1)
        r = np.concatenate(parts_r, axis=1)
        out_s = np.full((q.shape[0], k), -np.inf, dtype=np.float32)
        out_r = np.full((q.shape[0], k), -1, dtype=np.int64)
        for qi in range(q.shape[0]):
            ok = r[qi] >= 0
            order = np.lexsort((r[qi][ok], -s[qi][ok].astype(np.float64)))[:k]
            out_s[qi, : order.size] = s[qi][ok][order]
            out_r[qi, : order.size] = r[qi][ok][order]
        return out_s, out_r

    def guard_stats(self) -> tuple[int, int]:
        stats = [s.guard_stats() for s in self.shards]
        return sum(a for a, _ in stats), sum(b for _, b in stats)
