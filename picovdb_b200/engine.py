"""``DeviceStore``: numpy-facing wrapper of one ``pvdb_store_t`` handle.

This is the device-resident half of the reference's store (``_vectors`` + ``_active_indices``,
picovdb/pico_vdb.py:136,143) and the array-level search entry point the metric times:
normalised-or-raw queries in, ``(scores, rows)`` out.  All arithmetic happens in the CUDA
library; this file only marshals pointers.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional

import numpy as np

from . import _native as N


def _f32c(a) -> np.ndarray:
    return np.ascontiguousarray(a, dtype=np.float32)


def _i64c(a) -> np.ndarray:
    return np.ascontiguousarray(a, dtype=np.int64)


def _ptr(a: Optional[np.ndarray]):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def pack_row_mask(mask: np.ndarray) -> np.ndarray:
    """bool (n,) -> uint32 words, bit (r & 31) of word (r >> 5) (the C-ABI bitmap layout)."""
    mask = np.asarray(mask, dtype=bool)
    n = mask.shape[0]
    nwords = (n + 31) // 32
    packed = np.packbits(mask, bitorder="little")
    buf = np.zeros(nwords * 4, dtype=np.uint8)
    buf[: packed.shape[0]] = packed
    return buf.view("<u4")


def unpack_row_mask(words: np.ndarray, n: int) -> np.ndarray:
    bits = np.unpackbits(np.ascontiguousarray(words, dtype="<u4").view(np.uint8), bitorder="little")
    return bits[:n].astype(bool)


class DeviceStore:
    """Owns one native store on one GPU."""

    download_into = True  # download(..., out=) streams into a caller-provided (mapped) array

    def __init__(
        self,
        dim: int,
        device: int = 0,
        reserve_rows: int = 0,
        keep_f32: bool = True,
        bf16_mirror: bool = False,
        fixed_capacity: bool = False,
    ) -> None:
        self._lib = N.load()
        flags = (N.STORE_F32 if keep_f32 else 0) | (N.STORE_BF16 if bf16_mirror else 0)
        if fixed_capacity:
            flags |= N.STORE_FIXED_CAPACITY
        h = C.c_void_p()
        N.check(self._lib.pvdb_store_create(C.byref(h), int(device), int(dim), int(reserve_rows), flags))
        self._h = h
        self._owned = True
        self.bf16_only = bool(bf16_mirror and not keep_f32)
        self.dim = int(dim)
        self.device = int(device)

    @classmethod
    def from_handle(cls, handle: C.c_void_p, dim: int, device: int) -> "DeviceStore":
        """Non-owning view of a store that belongs to a ``pvdb_group_t`` (closing it is a no-op)."""
        self = cls.__new__(cls)
        self._lib = N.load()
        self._h = handle
        self._owned = False
        self.dim = int(dim)
        self.device = int(device)
        info = self.info()
        self.bf16_only = bool(info.flags & N.STORE_BF16) and not (info.flags & N.STORE_F32)
        return self

    # -- lifecycle ---------------------------------------------------------------------------
    def close(self) -> None:
        h, self._h = getattr(self, "_h", None), None
        if h is not None and h.value and getattr(self, "_owned", True):
            self._lib.pvdb_store_destroy(h)

    def __del__(self) -> None:  # pragma: no cover - interpreter shutdown ordering
        try:
            self.close()
        except Exception:
            pass

    @property
    def handle(self) -> C.c_void_p:
        if self._h is None:
            raise RuntimeError("DeviceStore is closed")
        return self._h

    def info(self) -> N.StoreInfo:
        out = N.StoreInfo()
        N.check(self._lib.pvdb_store_info(self.handle, C.byref(out)))
        return out

    @property
    def rows(self) -> int:
        return int(self.info().rows)

    def reserve(self, rows: int) -> None:
        N.check(self._lib.pvdb_store_reserve(self.handle, int(rows)))

    def set_row_base(self, base: int) -> None:
        N.check(self._lib.pvdb_store_set_row_base(self.handle, int(base)))

    # -- writes ------------------------------------------------------------------------------
    def upsert_rows(self, vecs: np.ndarray, rows: np.ndarray) -> None:
        """rows[i] := normalise(vecs[i]); rows must be unique within one call."""
        vecs = _f32c(vecs)
        rows = _i64c(rows)
        if vecs.ndim != 2 or vecs.shape[1] != self.dim or rows.shape != (vecs.shape[0],):
            raise ValueError(f"upsert_rows expects (n, {self.dim}) vectors and (n,) rows")
        N.check(self._lib.pvdb_store_upsert(self.handle, _ptr(vecs), _ptr(rows), vecs.shape[0]))

    def upsert_range(self, vecs: np.ndarray, row0: int) -> None:
        vecs = _f32c(vecs)
        if vecs.ndim != 2 or vecs.shape[1] != self.dim:
            raise ValueError(f"upsert_range expects (n, {self.dim}) vectors")
        N.check(self._lib.pvdb_store_upsert_range(self.handle, _ptr(vecs), int(row0), vecs.shape[0]))

    def upsert_range_dev(self, dev_ptr: int, row0: int, n: int, stream: int = 0) -> None:
        """Device-resident source (n x dim fp32 at ``dev_ptr``), enqueued on ``stream``."""
        N.check(
            self._lib.pvdb_store_upsert_range_dev(
                self.handle, C.c_void_p(dev_ptr), int(row0), int(n), C.c_void_p(stream or None)
            )
        )

    def upsert_rows_dev(self, dev_vecs: int, dev_rows: int, n: int, max_row: int, stream: int = 0) -> None:
        N.check(
            self._lib.pvdb_store_upsert_dev(
                self.handle, C.c_void_p(dev_vecs), C.c_void_p(dev_rows), int(n), int(max_row),
                C.c_void_p(stream or None),
            )
        )

    def delete_rows(self, rows) -> None:
        rows = _i64c(rows)
        N.check(self._lib.pvdb_store_delete(self.handle, _ptr(rows), rows.shape[0]))

    def upload(self, vecs: np.ndarray, row0: int = 0, active: Optional[np.ndarray] = None) -> None:
        """Raw load of already-normalised rows; ``active`` is a bool mask over those rows."""
        vecs = _f32c(vecs)
        if vecs.ndim != 2 or vecs.shape[1] != self.dim:
            raise ValueError(f"upload expects (n, {self.dim}) vectors")
        bits = None if active is None else pack_row_mask(active)
        N.check(self._lib.pvdb_store_upload(self.handle, int(row0), vecs.shape[0], _ptr(vecs), _ptr(bits)))

    def compact(self, keep_rows) -> None:
        keep = _i64c(keep_rows)
        N.check(self._lib.pvdb_store_compact(self.handle, _ptr(keep), keep.shape[0]))

    # -- reads -------------------------------------------------------------------------------
    def fetch_rows(self, rows) -> np.ndarray:
        rows = _i64c(rows)
        out = np.empty((rows.shape[0], self.dim), dtype=np.float32)
        N.check(self._lib.pvdb_store_fetch(self.handle, _ptr(rows), rows.shape[0], _ptr(out)))
        return out

    def download(self, row0: int = 0, n: Optional[int] = None, out: Optional[np.ndarray] = None) -> np.ndarray:
        """Rows as fp32.  ``out``: a C-contiguous (n, dim) float32 destination (e.g. a slice of a
        memory-mapped .npy file) that the rows are streamed into directly."""
        if n is None:
            n = self.rows - row0
        if out is None:
            out = np.empty((n, self.dim), dtype=np.float32)
        elif out.shape != (n, self.dim) or out.dtype != np.float32 or not out.flags["C_CONTIGUOUS"]:
            raise ValueError("download(out=...) needs a C-contiguous float32 (n, dim) array")
        N.check(self._lib.pvdb_store_download(self.handle, int(row0), int(n), _ptr(out)))
        return out

    def download_bf16(self, row0: int = 0, n: Optional[int] = None, out: Optional[np.ndarray] = None) -> np.ndarray:
        """Rows of the bf16 mirror as stored: (n, dim) uint16 bit patterns."""
        if n is None:
            n = self.rows - row0
        if out is None:
            out = np.empty((n, self.dim), dtype=np.uint16)
        elif out.shape != (n, self.dim) or out.dtype != np.uint16 or not out.flags["C_CONTIGUOUS"]:
            raise ValueError("download_bf16(out=...) needs a C-contiguous uint16 (n, dim) array")
        N.check(self._lib.pvdb_store_download_bf16(self.handle, int(row0), int(n), _ptr(out)))
        return out

    def upload_bf16(self, vecs16: np.ndarray, row0: int = 0, active: Optional[np.ndarray] = None) -> None:
        """Raw load of bf16 bit patterns into a bf16-only store."""
        v = np.ascontiguousarray(vecs16, dtype=np.uint16)
        if v.ndim != 2 or v.shape[1] != self.dim:
            raise ValueError(f"upload_bf16 expects (n, {self.dim}) uint16 rows")
        bits = None if active is None else pack_row_mask(active)
        N.check(self._lib.pvdb_store_upload_bf16(self.handle, int(row0), v.shape[0], _ptr(v), _ptr(bits)))

    def write_file(self, path: str, file_offset: int, row0: int, n: int, as_bf16: bool = False) -> None:
        """Rows [row0, row0 + n) into the existing file at ``file_offset`` (see pvdb_store_write_file)."""
        N.check(self._lib.pvdb_store_write_file(self.handle, os.fsencode(path), int(file_offset), int(row0), int(n),
                                                1 if as_bf16 else 0))

    def active_mask(self) -> np.ndarray:
        n = self.rows
        words = np.zeros((n + 31) // 32, dtype="<u4")
        N.check(self._lib.pvdb_store_active_bits(self.handle, _ptr(words)))
        return unpack_row_mask(words, n)

    def search(
        self,
        queries: np.ndarray,
        k: int,
        prefilter: Optional[np.ndarray] = None,
        precision: str = "auto",
        normalized: bool = False,
        rescore: bool = True,
        scan_only: bool = False,
        guard: bool = True,
    ) -> tuple[np.ndarray, np.ndarray]:
        """(Q, dim) fp32 queries -> (scores (Q, k) f32 descending, rows (Q, k) int64).

        ``prefilter`` is a bool row mask or already-packed uint32 words; rows must have both
        their active bit and their prefilter bit set to be scored.  Short results are padded with
        -inf / -1.  ``guard=False`` skips the exactness guard of the tensor-core precisions (see
        ``pvdb_store_guard_stats`` in the header).
        """
        q = _f32c(queries)
        if q.ndim != 2 or q.shape[1] != self.dim:
            raise ValueError(f"search expects (Q, {self.dim}) queries")
        nq = q.shape[0]
        k = int(k)
        if k < 1:
            raise ValueError("k must be >= 1")
        bits = None
        if prefilter is not None:
            pf = np.asarray(prefilter)
            bits = pack_row_mask(pf) if pf.dtype == np.bool_ else np.ascontiguousarray(pf, dtype="<u4")
        flags = N.PRECISIONS[precision]
        if normalized:
            flags |= N.SEARCH_QUERIES_NORMALIZED
        if not rescore:
            flags |= N.SEARCH_NO_RESCORE
        if scan_only:
            flags |= N.SEARCH_SCAN_ONLY
        if not guard:
            flags |= N.SEARCH_NO_GUARD
        scores = np.empty((nq, k), dtype=np.float32)
        rows = np.empty((nq, k), dtype=np.int64)
        N.check(self._lib.pvdb_search(self.handle, _ptr(q), nq, k, _ptr(bits), flags, _ptr(scores), _ptr(rows)))
        return scores, rows

    def _search_flags(self, precision, normalized, rescore, scan_only, guard) -> int:
        flags = N.PRECISIONS[precision]
        if normalized:
            flags |= N.SEARCH_QUERIES_NORMALIZED
        if not rescore:
            flags |= N.SEARCH_NO_RESCORE
        if scan_only:
            flags |= N.SEARCH_SCAN_ONLY
        if not guard:
            flags |= N.SEARCH_NO_GUARD
        return flags

    def search_exchange(self, ex: "Exchange", queries: np.ndarray, k: int, prefilter: Optional[np.ndarray] = None,
                        precision: str = "auto", normalized: bool = False, rescore: bool = True,
                        scan_only: bool = False, guard: bool = True) -> tuple[np.ndarray, np.ndarray]:
        """``search`` on this SHARD with the cross-GPU exchange fused in: returns the merged top k over
        all shards (global rows).  Every rank must make the same call.  ``prefilter``: this shard's rows."""
        q = _f32c(queries)
        if q.ndim != 2 or q.shape[1] != self.dim:
            raise ValueError(f"search expects (Q, {self.dim}) queries")
        nq, k = q.shape[0], int(k)
        bits = None
        if prefilter is not None:
            pf = np.asarray(prefilter)
            bits = pack_row_mask(pf) if pf.dtype == np.bool_ else np.ascontiguousarray(pf, dtype="<u4")
        scores = np.empty((nq, k), dtype=np.float32)
        rows = np.empty((nq, k), dtype=np.int64)
        N.check(self._lib.pvdb_search_exchange(self.handle, ex.handle, _ptr(q), nq, k, _ptr(bits),
                                               self._search_flags(precision, normalized, rescore, scan_only, guard),
                                               _ptr(scores), _ptr(rows)))
        return scores, rows

    def search_exchange_dev(self, ex: "Exchange", d_queries: int, nq: int, k: int, d_scores: int, d_rows: int,
                            d_prefilter: int = 0, precision: str = "auto", normalized: bool = False,
                            rescore: bool = True, stream: int = 0, scan_only: bool = False, guard: bool = True) -> None:
        N.check(
            self._lib.pvdb_search_exchange_dev(
                self.handle, ex.handle, C.c_void_p(d_queries), int(nq), int(k), C.c_void_p(d_prefilter or None),
                self._search_flags(precision, normalized, rescore, scan_only, guard),
                C.c_void_p(d_scores), C.c_void_p(d_rows), C.c_void_p(stream or None),
            )
        )

    def guard_stats(self) -> tuple[int, int]:
        """(queries of the last search, of all searches) that the tensor-core path could not prove
        exact and answered again with the exact scan."""
        last, total = C.c_int64(0), C.c_int64(0)
        N.check(self._lib.pvdb_store_guard_stats(self.handle, C.byref(last), C.byref(total)))
        return int(last.value), int(total.value)

    # -- metadata columns (on-device dict `where` filters) ---------------------------------------
    MAX_COLUMNS = 16  # pvdb_store::kMaxColumns
    applies_row_base = True  # result rows already include set_row_base()'s offset

    def column_write(self, column: int, codes: np.ndarray, rows: Optional[np.ndarray] = None, row0: int = 0) -> None:
        """codes[i] (int32 >= 0, -1 = absent) for rows[i], or for the consecutive rows from row0."""
        codes = np.ascontiguousarray(codes, dtype=np.int32)
        r = None if rows is None else _i64c(rows)
        N.check(self._lib.pvdb_store_column_write(self.handle, int(column), _ptr(r), int(row0), _ptr(codes),
                                                  codes.shape[0]))

    def column_drop(self, column: int) -> None:
        N.check(self._lib.pvdb_store_column_drop(self.handle, int(column)))

    def search_where(self, queries: np.ndarray, k: int, column: int, wanted_codes, extra: Optional[np.ndarray] = None,
                     precision: str = "auto") -> tuple[np.ndarray, np.ndarray, int]:
        """Search restricted to rows whose `column` code is in `wanted_codes` (and `extra` mask, if
        given).  The bitmap is built on the device.  Returns (scores, rows, number of eligible rows)."""
        q = _f32c(queries)
        nq = q.shape[0]
        wanted = np.unique(np.asarray(list(wanted_codes), dtype=np.int32))  # sorted: the kernel bisects
        bits = None if extra is None else pack_row_mask(extra)
        scores = np.empty((nq, k), dtype=np.float32)
        rows = np.empty((nq, k), dtype=np.int64)
        ncand = C.c_int64(0)
        N.check(self._lib.pvdb_search_where(self.handle, _ptr(q), nq, int(k), int(column), _ptr(wanted),
                                            wanted.shape[0], _ptr(bits), N.PRECISIONS[precision], _ptr(scores),
                                            _ptr(rows), C.byref(ncand)))
        return scores, rows, int(ncand.value)

    def search_dev(self, d_queries: int, nq: int, k: int, d_scores: int, d_rows: int, d_prefilter: int = 0,
                   precision: str = "auto", normalized: bool = False, rescore: bool = True, stream: int = 0,
                   scan_only: bool = False, guard: bool = True) -> None:
        """Device pointers in / out; only enqueues work on ``stream`` (a ``cudaStream_t`` as int;
        0 = CUDA's legacy default stream, which is also torch's default stream)."""
        flags = N.PRECISIONS[precision]
        if normalized:
            flags |= N.SEARCH_QUERIES_NORMALIZED
        if not rescore:
            flags |= N.SEARCH_NO_RESCORE
        if scan_only:
            flags |= N.SEARCH_SCAN_ONLY
        if not guard:
            flags |= N.SEARCH_NO_GUARD
        N.check(
            self._lib.pvdb_search_dev(
                self.handle, C.c_void_p(d_queries), int(nq), int(k), C.c_void_p(d_prefilter or None), flags,
                C.c_void_p(d_scores), C.c_void_p(d_rows), C.c_void_p(stream or None),
            )
        )


class Exchange:
    """One GPU's end of the peer-memory top-k exchange (``pvdb_exchange_t``, include/picovdb_b200.h)."""

    def __init__(self, device: int, world: int, rank: int, slot_keys: int) -> None:
        self._lib = N.load()
        h = C.c_void_p()
        N.check(self._lib.pvdb_exchange_create(C.byref(h), int(device), int(world), int(rank), int(slot_keys)))
        self._h = h
        self.world, self.rank, self.slot_keys = int(world), int(rank), int(slot_keys)

    @property
    def handle(self) -> C.c_void_p:
        if self._h is None:
            raise RuntimeError("Exchange is closed")
        return self._h

    def ipc_handle(self) -> bytes:
        buf = C.create_string_buffer(N.IPC_HANDLE_BYTES)
        N.check(self._lib.pvdb_exchange_ipc_handle(self.handle, buf))
        return buf.raw

    def connect_ipc(self, handles: bytes) -> None:
        """``handles``: the ``world`` IPC handles concatenated in rank order."""
        if len(handles) != self.world * N.IPC_HANDLE_BYTES:
            raise ValueError("expected one IPC handle per rank")
        N.check(self._lib.pvdb_exchange_connect_ipc(self.handle, C.c_char_p(handles)))

    def disconnect(self) -> None:
        N.check(self._lib.pvdb_exchange_disconnect(self.handle))

    def launches(self) -> int:
        n = C.c_int64(0)
        N.check(self._lib.pvdb_exchange_info(self.handle, None, None, None, C.byref(n)))
        return int(n.value)

    def close(self) -> None:
        h, self._h = getattr(self, "_h", None), None
        if h is not None and h.value:
            self._lib.pvdb_exchange_destroy(h)

    def __del__(self) -> None:  # pragma: no cover
        try:
            self.close()
        except Exception:
            pass


def merge_topk_dev(device: int, d_scores: int, d_rows: int, nlists: int, nq: int, k: int, d_out_scores: int,
                   d_out_rows: int, stream: int = 0, scores_stride: int = 0, rows_stride: int = 0) -> None:
    """Merge ``nlists`` per-shard (nq, k) results; strides are in elements (0 = contiguous)."""
    lib = N.load()
    N.check(
        lib.pvdb_merge_topk_dev(
            int(device), C.c_void_p(d_scores), C.c_void_p(d_rows), int(nlists), int(nq), int(k),
            int(scores_stride), int(rows_stride),
            C.c_void_p(d_out_scores), C.c_void_p(d_out_rows), C.c_void_p(stream or None),
        )
    )
