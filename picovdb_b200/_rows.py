"""Host bookkeeping that scales to 10^7 - 10^8 rows.

The reference keeps one Python object per row in three containers -- ``_ids`` (list), ``_docs`` (list of
dicts) and ``_id2idx`` (dict), picovdb/pico_vdb.py:137-140 -- which is ~200 bytes per row: 20 GB for
the 100M-row north-star store, before the first query.  Bulk-ingested rows (``upsert_array`` without
ids / docs) do not need any of it: their id IS an integer that continues a range, their document is
``{"_id_": id}``.  The two classes below keep the reference's container semantics (the drop-in class
and the reference's tests index, iterate and compare them like lists / dicts) while storing such rows
as *implicit ranges*: O(1) memory per bulk call, documents materialised only for the <= Q*k rows a query
returns (pico_vdb.py:753-775).  Everything else -- explicit ids, metadata, updates, deletes -- lives in
ordinary Python lists / dicts exactly as before.
"""
from __future__ import annotations

import bisect
from typing import Any, Callable, Iterable, Iterator, Optional

import numpy as np


def _is_plain_int(x: Any) -> bool:
    return isinstance(x, (int, np.integer)) and not isinstance(x, (bool, np.bool_))


class RowSeq:
    """List of per-row values: explicit chunks (Python lists) and implicit chunks (row r of the chunk
    holds ``make(id0 + r - start)``), plus sparse overrides inside implicit chunks."""

    def __init__(self, make: Callable[[int], Any], items: Optional[Iterable[Any]] = None) -> None:
        self._make = make
        self._starts: list[int] = []       # first row of every chunk (ascending)
        self._chunks: list[Any] = []       # list (explicit) or (id0, n) tuple (implicit)
        self._n = 0
        self._over: dict[int, Any] = {}    # row -> value, rows of implicit chunks only
        if items is not None:
            self.extend(items)

    # ---- structure ------------------------------------------------------------------------
    def _locate(self, row: int) -> tuple[int, int]:
        c = bisect.bisect_right(self._starts, row) - 1
        return c, row - self._starts[c]

    def _norm(self, row: int) -> int:
        if row < 0:
            row += self._n
        if not 0 <= row < self._n:
            raise IndexError("row index out of range")
        return row

    @property
    def implicit_rows(self) -> int:
        return sum(c[1] for c in self._chunks if isinstance(c, tuple))

    def implicit_ranges(self) -> list[tuple[int, int, int]]:
        """(first row, rows, first id) of every implicit chunk."""
        return [(s, c[1], c[0]) for s, c in zip(self._starts, self._chunks) if isinstance(c, tuple)]

    # ---- list protocol --------------------------------------------------------------------
    def __len__(self) -> int:
        return self._n

    def __getitem__(self, row):
        if isinstance(row, slice):
            return [self[i] for i in range(*row.indices(self._n))]
        row = self._norm(int(row))
        c, off = self._locate(row)
        chunk = self._chunks[c]
        if isinstance(chunk, list):
            return chunk[off]
        if self._over and row in self._over:
            return self._over[row]
        return self._make(chunk[0] + off)

    def getter(self) -> Callable[[int], Any]:
        """``get(row)`` for rows already known to be in [0, len): what result assembly calls Q * k times per
        query (pico_vdb.py:753-775).  A store of explicit rows only is ONE list, so this is ``list.__getitem__``;
        otherwise a closure without the negative-index / slice handling of ``__getitem__``.  Valid until the
        sequence is next extended or replaced (queries hold the read lock)."""
        if len(self._chunks) == 1 and isinstance(self._chunks[0], list):
            return self._chunks[0].__getitem__
        starts, chunks, over, make = self._starts, self._chunks, self._over, self._make
        locate = bisect.bisect_right

        def get(row: int) -> Any:
            c = locate(starts, row) - 1
            chunk = chunks[c]
            if type(chunk) is list:
                return chunk[row - starts[c]]
            if over and row in over:
                return over[row]
            return make(chunk[0] + row - starts[c])

        return get

    def __setitem__(self, row: int, value: Any) -> None:
        row = self._norm(int(row))
        c, off = self._locate(row)
        chunk = self._chunks[c]
        if isinstance(chunk, list):
            chunk[off] = value
        else:
            self._over[row] = value

    def append(self, value: Any) -> None:
        self.extend((value,))

    def extend(self, items: Iterable[Any]) -> None:
        items = list(items)
        if not items:
            return
        if self._chunks and isinstance(self._chunks[-1], list):
            self._chunks[-1].extend(items)
        else:
            self._starts.append(self._n)
            self._chunks.append(items)
        self._n += len(items)

    def extend_range(self, id0: int, n: int) -> None:
        """n rows whose values are make(id0), make(id0 + 1), ..."""
        if n <= 0:
            return
        last = self._chunks[-1] if self._chunks else None
        if isinstance(last, tuple) and last[0] + last[1] == id0:
            self._chunks[-1] = (last[0], last[1] + n)      # continues the previous range
        else:
            self._starts.append(self._n)
            self._chunks.append((int(id0), int(n)))
        self._n += n

    def __iter__(self) -> Iterator[Any]:
        for start, chunk in zip(self._starts, self._chunks):
            if isinstance(chunk, list):
                yield from chunk
            else:
                id0, n = chunk
                over, make = self._over, self._make
                if over:
                    for off in range(n):
                        r = start + off
                        yield over[r] if r in over else make(id0 + off)
                else:
                    for off in range(n):
                        yield make(id0 + off)

    def __eq__(self, other) -> bool:
        if isinstance(other, (list, RowSeq)):
            return len(other) == self._n and all(a == b for a, b in zip(self, other))
        return NotImplemented

    def __repr__(self) -> str:
        return f"RowSeq(rows={self._n}, chunks={len(self._chunks)}, implicit={self.implicit_rows})"

    # ---- bulk helpers ---------------------------------------------------------------------
    def explicit_items(self) -> Iterator[tuple[int, Any]]:
        """(row, value) of every row that is NOT a pristine implicit row."""
        for start, chunk in zip(self._starts, self._chunks):
            if isinstance(chunk, list):
                for off, v in enumerate(chunk):
                    yield start + off, v
        yield from sorted(self._over.items())

    def take_sorted(self, keep: np.ndarray) -> "RowSeq":
        """New sequence whose row j is this sequence's row keep[j] (keep strictly ascending): runs of
        consecutive kept rows inside an implicit chunk stay implicit."""
        out = RowSeq(self._make)
        keep = np.asarray(keep, dtype=np.int64)
        pos = 0
        for start, chunk in zip(self._starts, self._chunks):
            n = len(chunk) if isinstance(chunk, list) else chunk[1]
            end = int(np.searchsorted(keep, start + n, side="left"))
            rows = keep[pos:end]
            pos = end
            if rows.size == 0:
                continue
            if isinstance(chunk, list):
                out.extend([chunk[int(r) - start] for r in rows])
                continue
            id0 = chunk[0]
            breaks = np.flatnonzero(np.diff(rows) != 1) + 1
            for run in np.split(rows, breaks):
                a, m = int(run[0]), int(run.size)
                touched = [r for r in range(a, a + m) if r in self._over] if self._over else []
                if not touched:
                    out.extend_range(id0 + a - start, m)
                else:  # a run with overridden rows is cut around them
                    cur = a
                    for r in touched:
                        out.extend_range(id0 + cur - start, r - cur)
                        out.append(self._over[r])
                        cur = r + 1
                    out.extend_range(id0 + cur - start, a + m - cur)
        return out

    # ---- persistence ----------------------------------------------------------------------
    def to_compact(self) -> dict:
        chunks = []
        for chunk in self._chunks:
            chunks.append(["L", chunk] if isinstance(chunk, list) else ["R", chunk[0], chunk[1]])
        return {"picovdb_b200_rows": 1, "n": self._n, "chunks": chunks,
                "overrides": [[r, v] for r, v in sorted(self._over.items())]}

    @classmethod
    def from_compact(cls, make: Callable[[int], Any], obj: dict) -> "RowSeq":
        out = cls(make)
        for ch in obj["chunks"]:
            if ch[0] == "L":
                out.extend(ch[1])
            else:
                out._starts.append(out._n)          # keep chunk boundaries as saved (no merging)
                out._chunks.append((int(ch[1]), int(ch[2])))
                out._n += int(ch[2])
        for r, v in obj.get("overrides", []):
            out._over[int(r)] = v
        if out._n != obj["n"]:
            raise ValueError("corrupt compact row list")
        return out


class IdMap:
    """``_id2idx``: id -> row.  Explicit ids live in a dict; bulk rows are ranges of consecutive integer
    ids mapped to consecutive rows, minus the ids deleted since."""

    def __init__(self) -> None:
        self._d: dict[Any, int] = {}
        self._id0: list[int] = []      # sorted first ids of the ranges
        self._rng: list[tuple[int, int, int]] = []   # (id0, n, row0), parallel to _id0
        self._gone: set[int] = set()   # ids inside a range that were deleted / re-pointed
        self._n_implicit = 0

    # ---- ranges ---------------------------------------------------------------------------
    def add_range(self, id0: int, n: int, row0: int) -> None:
        if n <= 0:
            return
        if self.overlaps(id0, n):
            raise ValueError("ids of the new range are already present")
        i = bisect.bisect_left(self._id0, id0)
        self._id0.insert(i, int(id0))
        self._rng.insert(i, (int(id0), int(n), int(row0)))
        self._n_implicit += n

    def overlaps(self, id0: int, n: int) -> bool:
        """Is any id of [id0, id0 + n) present (explicitly or in a range, deleted ones excepted)?"""
        for a, m, _ in self._rng:
            lo, hi = max(a, id0), min(a + m, id0 + n)
            if lo < hi:
                if hi - lo > len(self._gone) or any(i not in self._gone for i in range(lo, hi)):
                    return True
        if self._d:
            if n < len(self._d):
                return any(i in self._d for i in range(id0, id0 + n))
            return any(_is_plain_int(i) and id0 <= i < id0 + n for i in self._d)
        return False

    def _range_row(self, key: Any) -> Optional[int]:
        if not self._rng or not _is_plain_int(key):
            return None
        key = int(key)
        i = bisect.bisect_right(self._id0, key) - 1
        if i < 0:
            return None
        a, m, row0 = self._rng[i]
        if key >= a + m or key in self._gone:
            return None
        return row0 + (key - a)

    @property
    def implicit_count(self) -> int:
        return self._n_implicit - len(self._gone)

    def implicit_rows(self) -> np.ndarray:
        """Rows of all live range ids, ascending per range."""
        if not self._rng:
            return np.empty(0, dtype=np.int64)
        parts = []
        for a, m, row0 in self._rng:
            rows = np.arange(row0, row0 + m, dtype=np.int64)
            parts.append(rows)
        out = np.concatenate(parts)
        if self._gone:
            gone_rows = np.fromiter((self._range_row_raw(g) for g in self._gone), dtype=np.int64, count=len(self._gone))
            out = out[~np.isin(out, gone_rows)]
        return out

    def _range_row_raw(self, key: int) -> int:
        i = bisect.bisect_right(self._id0, key) - 1
        a, _, row0 = self._rng[i]
        return row0 + (key - a)

    def sorted_rows(self) -> np.ndarray:
        rows = self.implicit_rows()
        if self._d:
            rows = np.concatenate([rows, np.fromiter(self._d.values(), dtype=np.int64, count=len(self._d))])
        rows.sort()
        return rows

    # ---- dict protocol --------------------------------------------------------------------
    def __len__(self) -> int:
        return len(self._d) + self.implicit_count

    def __bool__(self) -> bool:
        return len(self) > 0

    def __contains__(self, key: Any) -> bool:
        try:
            if key in self._d:
                return True
        except TypeError:
            return False
        return self._range_row(key) is not None

    def get(self, key: Any, default: Any = None) -> Any:
        try:
            row = self._d.get(key)
        except TypeError:
            return default
        if row is not None:
            return row
        row = self._range_row(key)
        return default if row is None else row

    def __getitem__(self, key: Any) -> int:
        row = self.get(key)
        if row is None:
            raise KeyError(key)
        return row

    def __setitem__(self, key: Any, row: int) -> None:
        if self._range_row(key) is not None:
            self._gone.add(int(key))       # the id now points somewhere else: the explicit entry wins
        self._d[key] = row

    def pop(self, key: Any, default: Any = None) -> Any:
        try:
            if key in self._d:
                return self._d.pop(key)
        except TypeError:
            return default
        row = self._range_row(key)
        if row is None:
            return default
        self._gone.add(int(key))
        return row

    def update(self, pairs) -> None:
        for k, v in (pairs.items() if hasattr(pairs, "items") else pairs):
            self[k] = v

    def keys(self) -> Iterator[Any]:
        yield from self._d.keys()
        for a, m, _ in self._rng:
            for i in range(a, a + m):
                if i not in self._gone:
                    yield i

    def values(self) -> Iterator[int]:
        yield from self._d.values()
        for r in self.implicit_rows().tolist():
            yield r

    def items(self) -> Iterator[tuple[Any, int]]:
        yield from self._d.items()
        for a, m, row0 in self._rng:
            for i in range(a, a + m):
                if i not in self._gone:
                    yield i, row0 + (i - a)

    def __iter__(self) -> Iterator[Any]:
        return self.keys()

    def __eq__(self, other) -> bool:
        if isinstance(other, (dict, IdMap)):
            return len(other) == len(self) and all(other.get(k) == v for k, v in self.items())
        return NotImplemented

    def __repr__(self) -> str:
        return f"IdMap(explicit={len(self._d)}, ranges={len(self._rng)}, implicit={self.implicit_count})"
