"""TEST-ONLY engine: lets the host logic of ``picovdb_b200.db.PicoVectorDB`` run without a GPU.

It implements the ``DeviceStore`` interface on top of ``oracle/picovdb_oracle.py``.  It lives in
``tests/`` and is injected by the ``host_db`` fixture; the product package never imports it and
has no way to select it.  What it covers: id / slot bookkeeping, candidate-mask construction,
result assembly, persistence format, locking -- i.e. everything in db.py that is not arithmetic.
"""
from __future__ import annotations

from typing import Optional

import numpy as np

from oracle import picovdb_oracle as O


class HostEngine:
    def __init__(self, dim, device=0, reserve_rows=0, keep_f32=True, bf16_mirror=False, fixed_capacity=False):
        self.dim = dim
        self.device = device
        self.vectors = np.zeros((0, dim), dtype=np.float32)
        self.active = np.zeros(0, dtype=bool)
        self.fixed_capacity = fixed_capacity
        self.reserved = int(reserve_rows)
        self.calls: list[str] = []
        self.columns: dict[int, np.ndarray] = {}

    @property
    def rows(self) -> int:
        return self.vectors.shape[0]

    def _grow(self, n: int) -> None:
        if n > self.rows:
            if self.fixed_capacity and n > self.reserved:
                raise RuntimeError("Database capacity exceeded")
            pad = n - self.rows
            self.vectors = np.vstack([self.vectors, np.zeros((pad, self.dim), np.float32)])
            self.active = np.concatenate([self.active, np.zeros(pad, bool)])

    def reserve(self, rows: int) -> None:
        self.reserved = max(self.reserved, rows)

    def upsert_rows(self, vecs, rows) -> None:
        self.calls.append("upsert_rows")
        rows = np.asarray(rows, dtype=np.int64)
        assert len(set(rows.tolist())) == rows.size, "rows must be unique within one call"
        self._grow(int(rows.max()) + 1)
        self.vectors[rows] = O.normalize_rows(vecs)
        self.active[rows] = True

    def upsert_range(self, vecs, row0) -> None:
        self.calls.append("upsert_range")
        n = vecs.shape[0]
        self._grow(row0 + n)
        self.vectors[row0 : row0 + n] = O.normalize_rows(vecs)
        self.active[row0 : row0 + n] = True

    def delete_rows(self, rows) -> None:
        self.calls.append("delete_rows")
        rows = np.asarray(rows, dtype=np.int64)
        self.vectors[rows] = 0
        self.active[rows] = False

    def upload(self, vecs, row0=0, active=None) -> None:
        self.calls.append("upload")
        n = vecs.shape[0]
        self._grow(row0 + n)
        self.vectors[row0 : row0 + n] = vecs
        self.active[row0 : row0 + n] = True if active is None else active

    def compact(self, keep_rows) -> None:
        self.calls.append("compact")
        keep = np.asarray(keep_rows, dtype=np.int64)
        self.vectors = np.ascontiguousarray(self.vectors[keep])
        self.active = np.ones(keep.size, dtype=bool)
        self.columns.clear()  # like the device store: compaction drops the columns

    def fetch_rows(self, rows) -> np.ndarray:
        return self.vectors[np.asarray(rows, dtype=np.int64)].copy()

    def download(self, row0=0, n=None) -> np.ndarray:
        n = self.rows - row0 if n is None else n
        return np.ascontiguousarray(self.vectors[row0 : row0 + n]).copy()

    def active_mask(self) -> np.ndarray:
        return self.active.copy()

    def search(self, queries, k, prefilter: Optional[np.ndarray] = None, precision="auto", normalized=False,
               rescore=True, scan_only=False, guard=True):
        self.calls.append("search")
        qn = np.ascontiguousarray(queries, np.float32) if normalized else O.prepare_queries(queries, self.dim)[0]
        pf = None
        if prefilter is not None:
            pf = np.zeros(self.rows, dtype=bool)
            m = np.asarray(prefilter, dtype=bool)
            pf[: m.size] = m[: self.rows]
        return O.search(self.vectors, qn, k, self.active, pf)

    def close(self) -> None:
        pass

    def set_row_base(self, base: int) -> None:
        self.row_base = int(base)

    # -- metadata columns (mirror of DeviceStore.column_write / search_where) ----------------
    MAX_COLUMNS = 16

    def column_write(self, column, codes, rows=None, row0=0) -> None:
        self.calls.append("column_write")
        col = self.columns.setdefault(column, np.full(0, -1, dtype=np.int32))
        codes = np.asarray(codes, dtype=np.int32)
        idx = np.asarray(rows, dtype=np.int64) if rows is not None else np.arange(row0, row0 + codes.size)
        assert idx.size == 0 or idx.max() < self.rows, "column rows must exist in the store"
        if col.size < self.rows:
            col = np.concatenate([col, np.full(self.rows - col.size, -1, dtype=np.int32)])
        col[idx] = codes
        self.columns[column] = col

    def column_drop(self, column) -> None:
        self.columns.pop(column, None)

    def search_where(self, queries, k, column, wanted_codes, extra=None, precision="auto"):
        self.calls.append("search_where")
        col = self.columns[column]
        if col.size < self.rows:
            col = np.concatenate([col, np.full(self.rows - col.size, -1, dtype=np.int32)])
        pf = np.isin(col[: self.rows], np.asarray(list(wanted_codes), dtype=np.int32)) & (col[: self.rows] >= 0)
        if extra is not None:
            pf &= np.asarray(extra, dtype=bool)[: self.rows]
        n_cand = int((pf & self.active).sum())
        qn = O.prepare_queries(queries, self.dim)[0]
        s, r = O.search(self.vectors, qn, k, self.active, pf)
        return s, r, n_cand
