"""``PicoVectorDB`` -- drop-in for wensheng/picovdb's class with the exact-search hot path on a B200.

Level-1 boundary of SURVEY.md 8(b): same constructor, methods, record schema (``_id_`` /
``_vector_`` / ``_metrics_``), error messages and storage files as the reference class
(picovdb/pico_vdb.py:97-1011), so user code, the reference's bench scripts and its tests run
against it unchanged.  What differs is where the numbers live and who computes them:

* the ``_vectors`` matrix and the active-row set are device resident (``DeviceStore``); ``upsert``
  stages the raw vectors once and a fused CUDA kernel normalises + scatters them
  (replaces pico_vdb.py:58-68, 422-472);
* ``query`` builds a row bitmap from ``ids`` / ``where`` (the reference's candidate builder,
  pico_vdb.py:604-658), hands raw queries + bitmap to ONE C-ABI call and assembles dict results
  from the returned ``(scores, rows)`` (pico_vdb.py:753-775).  Normalisation, scoring and top-k
  (pico_vdb.py:584-591, 683-714) all run on the GPU;
* host bookkeeping (``_ids`` / ``_docs`` / ``_id2idx`` / ``_free`` / ``_active_indices``), the RW
  lock, persistence format and the quirks listed in SURVEY.md (Q2, Q6, Q7, Q8) are kept.

There is no CPU compute path: constructing a DB without the CUDA extension or without a GPU
raises.  The FAISS/HNSW keyword arguments are accepted and ignored (approximate search is out of
scope; the class behaves like the reference with ``no_faiss=True``).
"""
from __future__ import annotations

import hashlib
import json
import logging
import os
import threading
import time
import warnings
from contextlib import contextmanager
from typing import Any, Callable, Literal, Optional, Union

import numpy as np

from ._rows import IdMap, RowSeq

Float = np.float32
_UPSERT_BLOCK_BYTES = 32 << 20   # staged vectors per device call of upsert() (tests shrink it)
ADAPTIVE_BUFFER = 32
ARGSORT_THRESHOLD = 0.2
K_ID = "_id_"
K_VECTOR = "_vector_"
K_METRICS = "_metrics_"
_HAS_FAISS = False  # the exact path never uses FAISS

logger = logging.getLogger("picovdb")

WhereT = Union[dict[str, Any], Callable[[dict[str, Any]], bool]]


# --------------------------------------------------------------------------- small helpers
def _ids_path(base: str) -> str:
    return f"{base}.ids.json"


def _meta_path(base: str) -> str:
    return f"{base}.meta.json"


def _vecs_path(base: str) -> str:
    return f"{base}.vecs.npy"


def _vecs16_path(base: str) -> str:
    """bf16 bit patterns of a bf16-only store, (rows, dim) uint16 in .npy framing -- half the size of
    the fp32 expansion; written instead of ``.vecs.npy`` unless ``save_dtype="f32"``."""
    return f"{base}.vecs.bf16.npy"


def _hash_vec(v: np.ndarray) -> str:
    return hashlib.md5(v.tobytes()).hexdigest()


def _normalize(v: np.ndarray) -> np.ndarray:
    """Host-side L2 normalisation (zero -> e0).  Used only to derive the md5 auto-id of a record
    without ``_id_`` so ids stay bit-compatible with the reference; the stored row is
    normalised by the device kernel."""
    vec = np.asarray(v, dtype=Float)
    n = float(np.linalg.norm(vec))
    if n == 0.0:
        unit = np.zeros_like(vec, dtype=Float)
        if unit.size:
            unit.flat[0] = Float(1.0)
        return unit
    return (vec / n).astype(Float, copy=False)


def _to_c_f32(a: np.ndarray) -> np.ndarray:
    return np.ascontiguousarray(a, dtype=Float)


def _timed(name: str):
    """DEBUG-level wall-clock logging, same message shape as the reference's decorator."""

    def decorator(func):
        def wrapper(*args, **kwargs):
            t0 = time.perf_counter()
            try:
                return func(*args, **kwargs)
            finally:
                logger.debug("%s took %.4f ms", name, (time.perf_counter() - t0) * 1000)

        wrapper.__name__ = getattr(func, "__name__", name)
        wrapper.__doc__ = func.__doc__
        return wrapper

    return decorator


class _RWLock:
    """Writer-exclusive / multi-reader lock (same contract as pico_vdb.py:1019-1063)."""

    def __init__(self) -> None:
        self._cond = threading.Condition(threading.Lock())
        self._readers = 0
        self._writer = False

    def acquire_read(self) -> None:
        with self._cond:
            self._cond.wait_for(lambda: not self._writer)
            self._readers += 1

    def release_read(self) -> None:
        with self._cond:
            self._readers -= 1
            if self._readers == 0:
                self._cond.notify_all()

    def acquire_write(self) -> None:
        with self._cond:
            self._cond.wait_for(lambda: not self._writer and self._readers == 0)
            self._writer = True

    def release_write(self) -> None:
        with self._cond:
            self._writer = False
            self._cond.notify_all()

    @contextmanager
    def read_lock(self):
        self.acquire_read()
        try:
            yield
        finally:
            self.release_read()

    @contextmanager
    def write_lock(self):
        self.acquire_write()
        try:
            yield
        finally:
            self.release_write()


class _ColumnIndex:
    """Columnar mirror of one metadata key: ``codes[row]`` is the dictionary code of
    ``docs[row].get(key)`` (-1 for deleted slots).  It turns the reference's per-query Python loop
    over candidate docs for ``{key: value}`` / ``{key: {"$in": [...]}}`` filters
    (pico_vdb.py:615-638; 21 ms at 100k rows) into one vectorised compare.  Built lazily on the
    first filter that names the key and maintained incrementally by upsert / delete / vacuum.
    Python equality semantics are kept because codes come from a dict lookup (1 == 1.0 == True)."""

    def __init__(self, key: str, docs: list) -> None:
        self.key = key
        self.vocab: dict[Any, int] = {}
        self.codes = np.full(max(len(docs), 16), -1, dtype=np.int32)
        self.n = len(docs)
        self.ok = True
        # device mirror (engine.column_write): slot number, rows uploaded so far, rows changed since
        self.dev_slot: Optional[int] = None
        self.dev_rows = 0
        self.dev_dirty: set[int] = set()
        if isinstance(docs, RowSeq):
            # bulk rows carry no metadata besides their id: only the explicit rows need a look
            if key == K_ID and docs.implicit_rows:
                self.ok = False
            for row, doc in docs.explicit_items():
                if doc is not None:
                    self.set(row, doc)
        else:
            for row, doc in enumerate(docs):
                if doc is not None:
                    self.set(row, doc)

    def _code(self, value: Any, create: bool) -> int:
        """Dictionary code of a STORED value (create=True) or of a filter value (create=False).  An
        unhashable stored value disables the column for good (the key falls back to the Python loop);
        an unhashable FILTER value only makes that one query unsupported (TypeError propagates to the
        caller) -- it says nothing about the stored data."""
        if not create:
            code = self.vocab.get(value)  # raises TypeError for an unhashable filter value
            return -1 if code is None else code
        try:
            code = self.vocab.get(value)
            if code is None:
                code = self.vocab[value] = len(self.vocab)
        except TypeError:
            self.ok = False
            return -1
        return code

    def set(self, row: int, doc: Optional[dict]) -> None:
        if row >= self.codes.shape[0]:
            grown = np.full(max(row + 1, 2 * self.codes.shape[0]), -1, dtype=np.int32)
            grown[: self.n] = self.codes[: self.n]
            self.codes = grown
        self.n = max(self.n, row + 1)
        self.codes[row] = -1 if doc is None else self._code(doc.get(self.key), True)
        if self.dev_slot is not None and row < self.dev_rows:
            self.dev_dirty.add(row)

    def wanted_codes(self, values) -> Optional[list[int]]:
        """Codes of the filter values that occur in the column (None if the key is unsupported)."""
        if not self.ok:
            return None
        out = []
        try:
            for v in values:
                c = self._code(v, False)
                if c >= 0:
                    out.append(c)
        except TypeError:  # unhashable filter value: unsupported for THIS query only
            return None
        return out

    def sync_device(self, engine, n: int) -> None:
        """Bring the device copy of the codes up to date for rows [0, n)."""
        self.resize(n)
        if self.dev_dirty and len(self.dev_dirty) * 4 > max(n, 1):
            self.dev_rows = 0          # cheaper to re-send everything
            self.dev_dirty.clear()
        if self.dev_dirty:
            rows = np.fromiter((r for r in self.dev_dirty if r < n), dtype=np.int64)
            if rows.size:
                engine.column_write(self.dev_slot, self.codes[rows], rows=rows)
            self.dev_dirty.clear()
        if self.dev_rows < n:
            engine.column_write(self.dev_slot, self.codes[self.dev_rows:n], row0=self.dev_rows)
            self.dev_rows = n

    def resize(self, n: int) -> None:
        if n > self.n:
            self.set(n - 1, None)
        self.n = n

    def match(self, values, n: int) -> Optional[np.ndarray]:
        """bool mask over rows [0, n) whose value equals one of ``values``; None if unsupported."""
        self.resize(n)
        wanted = self.wanted_codes(values)
        if wanted is None:
            return None
        codes = self.codes[:n]
        if not wanted:
            return np.zeros(n, dtype=bool)
        if len(wanted) == 1:
            return codes == wanted[0]
        return np.isin(codes, np.asarray(wanted, dtype=np.int32))

    def reindex(self, keep) -> None:
        keep = np.asarray(keep, dtype=np.int64)
        self.resize(int(keep.max()) + 1 if keep.size else 0)
        kept = self.codes[keep] if keep.size else np.empty(0, np.int32)
        self.codes = np.full(max(len(keep), 16), -1, dtype=np.int32)
        self.codes[: len(keep)] = kept
        self.n = len(keep)
        self.dev_slot, self.dev_rows = None, 0   # compaction drops the device columns
        self.dev_dirty.clear()


def _implicit_id(i: int) -> int:
    return i


def _implicit_doc(i: int) -> dict[str, Any]:
    return {K_ID: i}


# bulk stores above this many implicit rows are saved in the compact row-list form (see save())
COMPACT_ROWS_THRESHOLD = 1_000_000


def _rows_to_json(seq: RowSeq):
    """What save() writes for ``_ids`` / the documents: the reference's plain JSON list
    (pico_vdb.py:351-371) -- unless the store holds more than COMPACT_ROWS_THRESHOLD bulk rows, whose
    10^7-10^8 implicit entries are written as ranges (a form only this class reads back)."""
    if seq.implicit_rows > COMPACT_ROWS_THRESHOLD:
        return seq.to_compact()
    return seq[:]


def _rows_from_json(obj, make) -> RowSeq:
    if isinstance(obj, dict) and "picovdb_b200_rows" in obj:
        return RowSeq.from_compact(make, obj)
    return RowSeq(make, obj)


def _default_engine_factory(dim: int, **kw):
    """The product engine: CUDA or nothing."""
    from .engine import DeviceStore

    return DeviceStore(dim, **kw)


# --------------------------------------------------------------------------- the class
class PicoVectorDB:
    """Cosine-only vector DB with metadata persistence; exact search runs on a B200."""

    # Tests that exercise only the host logic replace this with a test-local engine
    # (tests/_host_engine.py).  The product never does.
    _engine_factory = staticmethod(_default_engine_factory)

    def __init__(
        self,
        embedding_dim: int = 1024,
        metric: Literal["cosine"] = "cosine",
        storage_file: str = "picovdb",
        use_memmap: bool = False,
        capacity: Optional[int] = None,
        no_faiss: bool = False,
        faiss_threads: Optional[int] = None,
        hnsw_m: Optional[int] = None,
        hnsw_ef_construction: Optional[int] = None,
        ef_search_default: Optional[int] = None,
        hnsw_ef_search_default: Optional[int] = None,
        faiss_incremental_threshold_ratio: float = 0.2,
        adaptive_buffer: Optional[int] = None,
        argsort_threshold: Optional[float] = None,
        # ---- B200 engine options (no reference counterpart) ----
        device: Optional[int] = None,
        devices: Optional[list[int]] = None,
        bf16_mirror: bool = False,
        keep_f32: bool = True,
        precision: str = "auto",
        save_dtype: Optional[str] = None,
        max_rows: Optional[int] = None,
    ) -> None:
        self._rwlock = _RWLock()
        self.dim = int(embedding_dim)
        self.metric = metric
        self._path = storage_file
        self._use_memmap = use_memmap
        self._capacity = capacity
        self._precision = precision
        if save_dtype not in (None, "f32", "bf16"):
            raise ValueError("save_dtype must be None, 'f32' or 'bf16'")
        self._save_dtype = save_dtype

        # list / dict semantics of the reference's containers (pico_vdb.py:137-143); bulk-ingested rows
        # are stored as implicit ranges (see _rows.py), so 10^8 rows do not cost 10^8 Python objects
        self._ids: RowSeq = RowSeq(_implicit_id)
        self._docs: RowSeq = RowSeq(_implicit_doc)
        self._free: list[int] = []
        self._id2idx: IdMap = IdMap()
        self._additional: dict[str, Any] = {}
        self._active_explicit: np.ndarray = np.empty(0, dtype=np.int64)  # rows of explicit ids, insertion order
        self._auto_hi = 0  # 1 + the largest integer id seen: default ids of upsert_array start no lower

        ab_env = os.getenv("PICOVDB_ADAPTIVE_BUFFER")
        thr_env = os.getenv("PICOVDB_ARGSORT_THRESHOLD")
        if adaptive_buffer is not None:
            self._adaptive_buffer = int(adaptive_buffer)
        else:
            self._adaptive_buffer = int(ab_env) if ab_env is not None else ADAPTIVE_BUFFER
        if argsort_threshold is not None:
            self._argsort_threshold = float(argsort_threshold)
        else:
            self._argsort_threshold = float(thr_env) if thr_env is not None else ARGSORT_THRESHOLD
        self._last_topk_strategy: Optional[str] = None
        self._last_k_eff: Optional[int] = None

        # HNSW knobs are accepted for signature compatibility and recorded, nothing else
        self._hnsw_m = int(hnsw_m) if hnsw_m is not None else 32
        self._hnsw_efc = int(hnsw_ef_construction) if hnsw_ef_construction is not None else 40
        if hnsw_ef_search_default is not None:
            self._faiss_ef_search = int(hnsw_ef_search_default)
        elif ef_search_default is not None:
            self._faiss_ef_search = int(ef_search_default)
        else:
            self._faiss_ef_search = 32
        self._faiss_incr_threshold_ratio = float(faiss_incremental_threshold_ratio)
        self._faiss = None
        self._dirty = False

        if device is None:
            device = int(os.getenv("PICOVDB_DEVICE", "0"))
        self._device = device
        # the reference ignores capacity= when it loads existing files (pico_vdb.py:227-284): only a
        # fresh DB is pinned to its pre-allocation
        loading = os.path.exists(_ids_path(storage_file)) and (
            os.path.exists(_vecs_path(storage_file)) or os.path.exists(_vecs16_path(storage_file)))
        factory = type(self)._engine_factory
        # max_rows: size of the engine's row space WITHOUT the reference's capacity= semantics (no
        # pre-filled slot lists, no free-slot dealing): rows are appended 0, 1, 2, ... as in a plain DB.
        # This is how a row-sharded engine (devices=[...] / ShardedPicoVectorDB), whose partition must be
        # fixed up front, takes a 10^8-row bulk load without 10^8 host-side slot entries.
        if max_rows is not None and capacity is not None:
            raise ValueError("give either capacity= (pre-allocated slots, as in the reference) or max_rows=")
        if devices is not None and len(devices) > 1:
            # one process, several GPUs: rows sharded over `devices`, searches merged over NVLink
            # inside the kernels (group.py); needs the total row count to lay out the partition
            if capacity is None and max_rows is None:
                raise ValueError("devices=[...] needs capacity= or max_rows= (the row partition over the GPUs is fixed)")
            from .group import GroupStore

            def factory(dim, **kw):  # noqa: E306
                kw.pop("device", None)
                return GroupStore(dim, devices, **kw)
        elif devices:
            device = int(devices[0])
            self._device = device
        self._engine = factory(
            self.dim,
            device=device,
            reserve_rows=int(capacity) if capacity else (int(max_rows) if max_rows else 0),
            keep_f32=keep_f32,
            bf16_mirror=bf16_mirror,
            fixed_capacity=capacity is not None and not loading,
        )
        self._host_cache: Optional[np.ndarray] = None  # lazily downloaded copy behind `_vectors`
        self._columns: dict[str, _ColumnIndex] = {}    # metadata key -> columnar index (lazy)
        self._col_lock = threading.RLock()             # guards _columns and the device mirrors among readers
        self._load_or_init()

    # ------------------------------------------------------------------ host mirror of the matrix
    @property
    def _vectors(self) -> np.ndarray:
        """Host copy of the device matrix, (rows, dim) C-contiguous fp32 (downloaded on demand)."""
        if self._host_cache is None:
            n = len(self._ids)
            self._host_cache = self._download(0, n) if n else np.empty((0, self.dim), dtype=Float)
        return self._host_cache

    def _download(self, row0: int, n: int) -> np.ndarray:
        """Rows [row0, row0 + n) as fp32; slots the engine has never written (a fresh ``capacity=``
        DB, pico_vdb.py:286-296) read as zeros, as in the reference's pre-allocated matrix."""
        have = max(0, min(n, int(self._engine.rows) - row0))
        if have == n:
            return self._engine.download(row0, n)
        out = np.zeros((n, self.dim), dtype=Float)
        if have:
            out[:have] = self._engine.download(row0, have)
        return out

    def _invalidate(self) -> None:
        self._host_cache = None

    def _note_id(self, _id: Any) -> None:
        if isinstance(_id, (int, np.integer)) and not isinstance(_id, (bool, np.bool_)) and _id >= self._auto_hi:
            self._auto_hi = int(_id) + 1

    def _default_ids(self, n: int) -> range:
        """Default ids of a bulk call: n consecutive integers from the slot count upwards, moved past
        every integer id already seen so that they never collide."""
        id0 = max(len(self._ids), self._auto_hi)
        return range(id0, id0 + n)

    @property
    def _active_indices(self) -> np.ndarray:
        """Rows of all live records (pico_vdb.py:143).  Explicit rows in insertion order, then the
        rows of bulk ranges; materialised on demand (O(rows) for a bulk store)."""
        if self._id2idx.implicit_count == 0:
            return self._active_explicit
        return np.concatenate([self._active_explicit, self._id2idx.implicit_rows()])

    @_active_indices.setter
    def _active_indices(self, value: np.ndarray) -> None:
        self._active_explicit = np.asarray(value, dtype=np.int64)

    # ------------------------------------------------------------------ persistence
    @_timed("load")
    def _load_or_init(self) -> None:
        ids_file, vecs_file, meta_file = _ids_path(self._path), _vecs_path(self._path), _meta_path(self._path)
        vecs16_file = _vecs16_path(self._path)
        # a bf16-only store reads its own bf16 file when there is one (no fp32 detour, half the bytes)
        use16 = os.path.exists(vecs16_file) and getattr(self._engine, "bf16_only", False) and (
            not os.path.exists(vecs_file) or os.path.getmtime(vecs16_file) >= os.path.getmtime(vecs_file))
        if os.path.exists(ids_file) and (use16 or os.path.exists(vecs_file)):
            logger.info("Loading existing DB …")
            with open(ids_file, "r", encoding="utf-8") as f:
                self._ids = _rows_from_json(json.load(f), _implicit_id)
            count = len(self._ids)
            if use16:
                vectors = np.load(vecs16_file, mmap_mode="r")
                if vectors.shape != (count, self.dim) or vectors.dtype != np.uint16:
                    raise ValueError(f"stored bf16 matrix has shape {vectors.shape}, expected ({count}, {self.dim})")
            else:
                vectors = self._read_vectors(vecs_file, count)  # memory-mapped; streamed to the device below
            if os.path.exists(meta_file):
                with open(meta_file, "r", encoding="utf-8") as f:
                    meta_json = json.load(f)
                self._docs = _rows_from_json(meta_json.get("data", [None] * count), _implicit_doc)
                self._additional = meta_json.get("additional_data", {})
            else:
                self._docs = RowSeq(_implicit_doc, [None] * count)
            active = np.zeros(count, dtype=bool)
            explicit_rows: list[int] = []
            implicit = {r0: (n, id0) for r0, n, id0 in self._ids.implicit_ranges()}
            for r0, n, _ in self._docs.implicit_ranges():
                # bulk range (compact form only): every row is live unless a later delete overrode it
                if implicit.get(r0, (None,))[0] != n:
                    raise ValueError("compact id / document ranges do not line up")
                self._id2idx.add_range(implicit[r0][1], n, r0)
                active[r0:r0 + n] = True
            for row, doc in self._docs.explicit_items():
                _id = self._ids[row]
                in_range = bool(active[row])
                if doc is None:
                    self._free.append(row)
                    if in_range:
                        self._id2idx.pop(_id, None)
                        active[row] = False
                elif _id is not None and not in_range:
                    self._id2idx[_id] = row
                    explicit_rows.append(row)
                    active[row] = True
            self._active_explicit = np.asarray(explicit_rows, dtype=np.int64)
            for _, n, id0 in self._ids.implicit_ranges():
                self._auto_hi = max(self._auto_hi, id0 + n)
            for _, _id in self._ids.explicit_items():
                self._note_id(_id)
            # stream the matrix to the device in row blocks (multiples of 32 rows so every block's
            # slice of the active bitmap starts on a word boundary); the file is only mapped, so a
            # store larger than host RAM still loads
            # (inside one call the library streams through pinned double buffers: the host copy of block
            # i+1 -- the page faults of the mapped file -- overlaps the DMA of block i)
            step = max(32, ((512 << 20) // (self.dim * 4)) // 32 * 32)
            put = self._engine.upload_bf16 if use16 else self._engine.upload
            for r0 in range(0, count, step):
                r1 = min(count, r0 + step)
                # the slice stays a lazy view of the mapped file: a sharded engine only reads its own rows
                put(vectors[r0:r1], r0, active[r0:r1])
            logger.info("Loaded %d active / %d total vectors", len(self._id2idx), count)
        else:
            if self._capacity is not None:
                cap = int(self._capacity)
                self._ids = RowSeq(_implicit_id, [None] * cap)
                self._docs = RowSeq(_implicit_doc, [None] * cap)
                # popped from the end, as in the reference; a row-sharded engine supplies an order that
                # deals the rows out over its shards so a partly filled store is balanced
                order = getattr(self._engine, "free_order", None)
                self._free = order(cap) if order is not None else list(range(cap))
                if self._use_memmap:
                    # keep the reference's observable side effect: a pre-sized raw file
                    np.memmap(vecs_file, dtype=Float, mode="w+", shape=(cap, self.dim)).flush()
            logger.info("No persisted data – fresh DB")

    def _read_vectors(self, vecs_file: str, count: int) -> np.ndarray:
        try:
            arr = np.load(vecs_file, mmap_mode="r")
        except (ValueError, OSError):
            # headerless pre-allocated file written by `capacity=` + `use_memmap=True`
            arr = np.fromfile(vecs_file, dtype=Float)
            if arr.size != count * self.dim:
                raise
            arr = arr.reshape(count, self.dim)
        if arr.dtype != Float:
            arr = _to_c_f32(arr)
        if arr.shape != (count, self.dim):
            raise ValueError(
                f"stored matrix has shape {arr.shape}, expected ({count}, {self.dim})"
            )
        return arr

    @_timed("save")
    def save(self) -> None:
        """Persist atomically: temp files first, then ``os.replace`` (pico_vdb.py:330-393)."""
        with self._rwlock.write_lock():
            ids_file, vecs_file, meta_file = _ids_path(self._path), _vecs_path(self._path), _meta_path(self._path)
            # a bf16-only store persists its mirror as it is unless the reference's fp32 file is asked for
            as16 = self._save_dtype == "bf16" or (self._save_dtype is None and getattr(self._engine, "bf16_only", False))
            if as16 and not hasattr(self._engine, "download_bf16"):
                as16 = False
            stale_vecs = vecs_file if as16 else _vecs16_path(self._path)
            if as16:
                vecs_file = _vecs16_path(self._path)
            tmp_ids = f"{ids_file}.tmp"
            tmp_vecs_base = f"{self._path}.vecs.tmp"
            tmp_vecs = f"{tmp_vecs_base}.npy"
            tmp_meta = f"{meta_file}.tmp"
            # a row-sharded engine (sharded.ShardedStore): every rank makes this call, ONE rank writes
            # the shared files, all ranks write their own rows of the matrix
            writer = getattr(self._engine, "is_writer", True)
            try:
                if writer:
                    with open(tmp_ids, "w", encoding="utf-8") as f:
                        json.dump(_rows_to_json(self._ids), f, ensure_ascii=False)
                self._write_vectors(tmp_vecs, as16)
                if writer:
                    with open(tmp_meta, "w", encoding="utf-8") as f:
                        json.dump(
                            {"embedding_dim": self.dim, "data": _rows_to_json(self._docs),
                             "additional_data": self._additional},
                            f,
                            ensure_ascii=False,
                        )
                    os.replace(tmp_ids, ids_file)
                    os.replace(tmp_vecs, vecs_file)
                    os.replace(tmp_meta, meta_file)
                    if os.path.exists(stale_vecs):   # the other precision's file would describe an older state
                        os.remove(stale_vecs)
                    logger.info("Saved %d vectors", len(self._ids))
                self._engine_barrier()
            finally:
                for tmp in (tmp_ids, tmp_vecs, tmp_meta) if writer else ():
                    if os.path.exists(tmp):
                        try:
                            os.remove(tmp)
                        except OSError:
                            pass

    def _write_vectors(self, path: str, as16: bool = False) -> None:
        """Write the ``.npy`` file (same header as ``np.save``) in row blocks straight from the
        device into the mapped file, so saving never needs a second full copy of the matrix in host
        memory; inside one block the library overlaps the device->host DMA with the copy into the file
        (pinned double buffers).  ``as16``: the bf16 mirror's bit patterns (uint16) instead of fp32."""
        n = len(self._ids)
        eng = self._engine
        writer = getattr(eng, "is_writer", True)
        sharded = hasattr(eng, "owned_rows")
        if not as16 and (n == 0 or (self._host_cache is not None and not sharded)):
            with open(path, "wb") as f:
                np.save(f, self._vectors)
            return
        from numpy.lib.format import open_memmap

        dtype = np.uint16 if as16 else Float
        if n == 0:
            with open(path, "wb") as f:
                np.save(f, np.empty((0, self.dim), dtype=dtype))
            return
        # the writer creates the file (header + size); with a sharded engine the other ranks then map
        # it and every rank stores the rows it owns
        out = open_memmap(path, mode="w+", dtype=dtype, shape=(n, self.dim)) if writer else None
        self._engine_barrier()
        if out is None:
            out = open_memmap(path, mode="r+")
        lo, hi = eng.owned_rows() if sharded else (0, n)
        lo, hi = max(lo, 0), min(hi, n)
        if hasattr(eng, "write_file"):
            # the library streams device -> pinned -> pwrite() into the file (the map above only created
            # header + size); slots the engine never wrote stay the zeros of the sparse file
            row_bytes = self.dim * out.dtype.itemsize
            offset = int(out.offset)
            del out
            hi = min(hi, int(eng.rows))
            if hi > lo:
                eng.write_file(path, offset + lo * row_bytes, lo, hi - lo, as16)
            self._engine_barrier()
            return
        step = max(1, (512 << 20) // (self.dim * out.dtype.itemsize))
        for r0 in range(lo, hi, step):
            r1 = min(hi, r0 + step)
            if as16:
                out[r0:r1] = eng.download_bf16(r0, r1 - r0)
            else:
                out[r0:r1] = self._download(r0, r1 - r0)
        out.flush()
        del out
        self._engine_barrier()

    def _engine_barrier(self) -> None:
        barrier = getattr(self._engine, "barrier", None)
        if barrier is not None:
            barrier()

    def flush(self) -> None:
        """No-op: the store is device resident; ``save()`` writes the files."""
        with self._rwlock.read_lock():
            return None

    # ------------------------------------------------------------------ counters
    def size(self) -> int:
        warnings.warn(
            "size() is deprecated: use count() for active items; capacity() will be added in a future release.",
            DeprecationWarning,
            stacklevel=2,
        )
        with self._rwlock.read_lock():
            return len(self._ids)

    def capacity(self) -> int:
        with self._rwlock.read_lock():
            return len(self._ids)

    def count(self) -> int:
        with self._rwlock.read_lock():
            return len(self._id2idx)

    def __len__(self) -> int:
        with self._rwlock.read_lock():
            return len(self._id2idx)

    # ------------------------------------------------------------------ mutators
    def upsert(self, items: list[dict[str, Any]]) -> dict[str, list[Any]]:
        """Insert or update records.  Host side: validation, id / slot bookkeeping, staging of the
        raw vectors.  Device side (one call): normalise, scatter, set active bits."""
        with self._rwlock.write_lock():
            report: dict[str, list[Any]] = {"update": [], "insert": []}
            staged: list[np.ndarray] = []
            staged_rows: list[int] = []
            slot_of_row: dict[int, int] = {}  # a row written twice in one call keeps the last vector
            appended_ids: list[Any] = []
            appended_docs: list[dict[str, Any]] = []
            new_active: list[int] = []
            # Vectors go to the device in blocks of ~32 MiB through ONE reused staging array: stacking 100k x 1024
            # rows into a fresh 400 MB array spent 0.46 s on first-touch page faults alone (80 % of the host
            # side of such a call); a block that stays cache- and TLB-warm costs a fifth of that.
            block_rows = max(256, _UPSERT_BLOCK_BYTES // (self.dim * 4))
            n_hint = len(items) if hasattr(items, "__len__") else block_rows
            stage_buf: Optional[np.ndarray] = None

            def flush() -> None:
                nonlocal stage_buf
                if not staged:
                    return
                m = len(staged)
                if stage_buf is None or stage_buf.shape[0] < m:
                    stage_buf = np.empty((max(m, min(block_rows, n_hint)), self.dim), dtype=Float)
                np.stack(staged, out=stage_buf[:m])
                rows_arr = np.asarray(staged_rows, dtype=np.int64)
                staged.clear()
                staged_rows.clear()
                slot_of_row.clear()
                self._engine.upsert_rows(stage_buf[:m], rows_arr)
                self._invalidate()

            try:
                for item in items:
                    raw = np.ascontiguousarray(item[K_VECTOR], dtype=Float)
                    if raw.ndim != 1:
                        raise ValueError(
                            f"upsert vector must be 1D with length {self.dim}; got shape {tuple(raw.shape)}"
                        )
                    if raw.shape[0] != self.dim:
                        raise ValueError(
                            f"upsert vector dim mismatch: expected {self.dim}, got {raw.shape[0]}"
                        )
                    meta = {k: v for k, v in item.items() if k != K_VECTOR}
                    item_id = meta.get(K_ID)
                    if item_id is None:
                        item_id = _hash_vec(_normalize(raw))
                    meta[K_ID] = item_id
                    self._note_id(item_id)
                    if item_id in self._id2idx:
                        row = self._id2idx[item_id]
                        if row < len(self._docs):
                            self._docs[row] = meta
                        else:
                            appended_docs[row - len(self._docs)] = meta
                        report["update"].append(item_id)
                    else:
                        if self._free:
                            row = self._free.pop()
                            self._ids[row] = item_id
                            self._docs[row] = meta
                        else:
                            if self._capacity is not None:
                                raise ValueError("Database capacity exceeded")
                            appended_ids.append(item_id)
                            appended_docs.append(meta)
                            row = len(self._ids) + len(appended_ids) - 1
                        new_active.append(row)
                        self._id2idx[item_id] = row
                        report["insert"].append(item_id)
                    for col in self._columns.values():
                        col.set(row, meta)
                    pos = slot_of_row.get(row)
                    if pos is None:
                        slot_of_row[row] = len(staged)
                        staged.append(raw)
                        staged_rows.append(row)
                        if len(staged) >= block_rows:
                            flush()    # (a later write to one of these rows lands after this block: last one wins)
                    else:
                        staged[pos] = raw
            finally:
                # whatever was accepted before a validation error is committed, as in the reference
                # (its row writes happen item by item, pico_vdb.py:428-449)
                if appended_ids:
                    self._ids.extend(appended_ids)
                    self._docs.extend(appended_docs)
                flush()
                if new_active:
                    add = np.asarray(new_active, dtype=np.int64)
                    self._active_explicit = (
                        np.append(self._active_explicit, add) if self._active_explicit.size else add
                    )
            return report

    def upsert_array(
        self,
        vectors: np.ndarray,
        ids: Optional[list[Any]] = None,
        docs: Optional[list[dict[str, Any]]] = None,
    ) -> list[Any]:
        """Bulk ingest of NEW records from an (n, dim) array: one device call, no per-item work.

        ``ids`` default to consecutive integers continuing from the current slot count; every id must be
        absent from the DB.  Without ``ids`` and ``docs`` the rows are *implicit*: no Python object is
        created per row (``_rows.py``) -- the id of row r is an integer of a range, its document
        ``{"_id_": id}`` is materialised when a query returns it -- which is what makes 10^7-10^8-row
        stores usable through this class.  A DB with free slots or ``capacity=`` fills those slots
        first (explicit bookkeeping).  Returns the ids (a ``range`` for implicit rows)."""
        vecs = _to_c_f32(vectors)
        if vecs.ndim != 2 or vecs.shape[1] != self.dim:
            raise ValueError(f"upsert_array expects shape (n, {self.dim}); got {tuple(vecs.shape)}")
        n = vecs.shape[0]
        with self._rwlock.write_lock():
            if (ids is not None and len(ids) != n) or (docs is not None and len(docs) != n):
                raise ValueError("ids / docs length does not match the number of vectors")
            row0 = len(self._ids)
            if self._capacity is not None or self._free:
                # slots come from the free list (popped from its end, as upsert does)
                if self._capacity is not None and n > len(self._free):
                    raise ValueError("Database capacity exceeded")
                n_slots = min(n, len(self._free))
                rows = [self._free.pop() for _ in range(n_slots)] + list(range(row0, row0 + n - n_slots))
                new_ids = list(self._default_ids(n)) if ids is None else list(ids)
                if len(set(new_ids)) != n or any(i in self._id2idx for i in new_ids):
                    self._free.extend(reversed(rows[:n_slots]))
                    raise ValueError("upsert_array ids must be unique and not present in the DB")
                self._engine.upsert_rows(vecs, np.asarray(rows, dtype=np.int64))
                self._invalidate()
                self._ids.extend([None] * (n - n_slots))
                self._docs.extend([None] * (n - n_slots))
                for j, (row, _id) in enumerate(zip(rows, new_ids)):
                    self._note_id(_id)
                    self._ids[row] = _id
                    self._docs[row] = {K_ID: _id} if docs is None else {**docs[j], K_ID: _id}
                    self._id2idx[_id] = row
                    for col in self._columns.values():
                        col.set(row, self._docs[row])
                add = np.asarray(rows, dtype=np.int64)
                self._active_explicit = np.append(self._active_explicit, add) if self._active_explicit.size else add
                return new_ids
            if ids is None and docs is None:
                # implicit rows: n consecutive integer ids, no per-row object
                auto = self._default_ids(n)
                if self._id2idx.overlaps(auto.start, n):
                    raise ValueError("upsert_array ids must be unique and not present in the DB")
                self._engine.upsert_range(vecs, row0)
                self._invalidate()
                self._ids.extend_range(auto.start, n)
                self._docs.extend_range(auto.start, n)
                self._id2idx.add_range(auto.start, n, row0)
                self._auto_hi = auto.stop
                for col in self._columns.values():
                    col.resize(row0 + n)      # bulk rows carry no metadata: every key reads "absent"
                return auto  # type: ignore[return-value]
            new_ids = list(self._default_ids(n)) if ids is None else list(ids)
            if len(set(new_ids)) != n or any(i in self._id2idx for i in new_ids):
                raise ValueError("upsert_array ids must be unique and not present in the DB")
            self._engine.upsert_range(vecs, row0)
            self._invalidate()
            self._ids.extend(new_ids)
            if docs is None:
                self._docs.extend({K_ID: i} for i in new_ids)
            else:
                self._docs.extend({**d, K_ID: i} for d, i in zip(docs, new_ids))
            self._id2idx.update(zip(new_ids, range(row0, row0 + n)))
            for _id in new_ids:
                self._note_id(_id)
            for col in self._columns.values():
                for row in range(row0, row0 + n):
                    col.set(row, self._docs[row])
            add = np.arange(row0, row0 + n, dtype=np.int64)
            self._active_explicit = np.append(self._active_explicit, add) if self._active_explicit.size else add
            return new_ids

    def store_additional_data(self, **kwargs) -> None:
        with self._rwlock.write_lock():
            self._additional.update(kwargs)

    def get_additional_data(self) -> dict[str, Any]:
        with self._rwlock.read_lock():
            return self._additional

    def delete(self, ids: list[Any]) -> list[Any]:
        """Delete by id; returns the ids that existed.  Device: clear bits + zero rows."""
        with self._rwlock.write_lock():
            removed: list[Any] = []
            rows: list[int] = []
            for _id in ids:
                row = self._id2idx.pop(_id, None)
                if row is not None:
                    self._docs[row] = None
                    for col in self._columns.values():
                        col.set(row, None)
                    self._free.append(row)
                    rows.append(row)
                    removed.append(_id)
            if rows:
                self._engine.delete_rows(np.asarray(rows, dtype=np.int64))
                self._invalidate()
                if self._active_explicit.size:
                    gone = np.asarray(rows, dtype=np.int64)
                    self._active_explicit = self._active_explicit[~np.isin(self._active_explicit, gone)]
            return removed

    # ------------------------------------------------------------------ search
    def _validate_queries(self, query_vecs) -> tuple[np.ndarray, bool]:
        raw = np.ascontiguousarray(query_vecs, dtype=Float)
        if raw.ndim == 1:
            if raw.shape[0] != self.dim:
                raise ValueError(f"query vector dim mismatch: expected {self.dim}, got {raw.shape[0]}")
            return raw[None, :], True
        if raw.ndim == 2:
            if raw.shape[1] != self.dim:
                raise ValueError(
                    f"query vectors dim mismatch: expected last dim {self.dim}, got {raw.shape[1]}"
                )
            return raw, False
        raise ValueError(
            f"query expects 1D or 2D array with last dim {self.dim}; got shape {tuple(raw.shape)}"
        )

    def _candidate_mask(self, where: Optional[WhereT], ids: Optional[list[Any]]) -> Optional[np.ndarray]:
        """Row mask for ``ids`` / ``where`` (None = every active row).  Same selection rules as the
        reference's candidate builder (pico_vdb.py:604-658): ``ids`` -> mapped rows; a one-key dict is
        equality or ``{"$in": [...]}`` over the candidate docs; anything else is called as a
        predicate over the active docs and intersected."""
        if ids is None and where is None:
            return None
        n = len(self._ids)
        mask = np.zeros(n, dtype=bool)
        docs = self._docs
        if ids is not None:
            for s in ids:
                row = self._id2idx.get(s)
                if row is not None:
                    mask[row] = True
            base_rows = np.flatnonzero(mask)
        else:
            base_rows = self._active_indices
        if where is None:
            return mask
        if isinstance(where, dict) and len(where) == 1:
            ((key, val),) = where.items()
            is_in = isinstance(val, dict) and set(val.keys()) == {"$in"}
            with self._col_lock:  # concurrent readers share (and lazily build) the column index
                col = self._columns.get(key) if isinstance(key, str) else None
                if col is None and isinstance(key, str):
                    col = self._columns[key] = _ColumnIndex(key, docs)
                hit = None
                if col is not None and col.ok:
                    try:
                        hit = col.match(set(val["$in"]) if is_in else (val,), n)
                    except TypeError:  # unhashable filter value
                        hit = None
            if hit is not None:
                return hit if ids is None else (hit & mask)
            if isinstance(val, dict) and set(val.keys()) == {"$in"}:
                wanted = set(val["$in"])
                keep = [i for i in base_rows if docs[i] is not None and docs[i].get(key) in wanted]
            else:
                keep = [i for i in base_rows if docs[i] is not None and docs[i].get(key) == val]
            out = np.zeros(n, dtype=bool)
            if keep:
                out[np.asarray(keep, dtype=np.int64)] = True
            return out
        passed = np.zeros(n, dtype=bool)
        for i in self._active_indices:
            d = docs[i]
            if d is not None and where(d):  # type: ignore[operator]
                passed[i] = True
        return passed if ids is None else (mask & passed)

    def _device_where(self, where: Optional[WhereT], ids: Optional[list[Any]]):
        """(column slot, wanted codes, optional ids mask) when a one-key dict filter can be evaluated
        on the device, else None (callable filters, unhashable values, engines without columns)."""
        eng = self._engine
        if not (isinstance(where, dict) and len(where) == 1 and hasattr(eng, "search_where")):
            return None
        ((key, val),) = where.items()
        if not isinstance(key, str):
            return None
        is_in = isinstance(val, dict) and set(val.keys()) == {"$in"}
        col = self._columns.get(key)
        if col is None:
            col = self._columns[key] = _ColumnIndex(key, self._docs)
        if not col.ok:
            return None
        try:
            wanted = col.wanted_codes(set(val["$in"]) if is_in else (val,))
        except TypeError:
            return None
        if wanted is None:
            return None
        if col.dev_slot is None:
            used = {c.dev_slot for c in self._columns.values() if c.dev_slot is not None}
            free = [i for i in range(getattr(eng, "MAX_COLUMNS", 16)) if i not in used]
            if not free:  # all device columns taken: evict one (it is re-sent when used again)
                victim = next(c for c in self._columns.values() if c.dev_slot is not None)
                free = [victim.dev_slot]
                victim.dev_slot, victim.dev_rows = None, 0
                victim.dev_dirty.clear()
            col.dev_slot, col.dev_rows = free[0], 0
            col.dev_dirty.clear()
        n = len(self._ids)
        col.sync_device(eng, n)
        extra = None
        if ids is not None:
            extra = np.zeros(n, dtype=bool)
            for s in ids:
                row = self._id2idx.get(s)
                if row is not None:
                    extra[row] = True
        return col.dev_slot, wanted, extra

    def search(
        self,
        query_vecs: np.ndarray,
        top_k: int = 10,
        prefilter: Optional[np.ndarray] = None,
        precision: Optional[str] = None,
    ) -> tuple[np.ndarray, np.ndarray]:
        """Array-level search: (Q, dim) or (dim,) queries -> (scores (Q, k) f32, rows (Q, k) int64).

        This is the entry point the throughput metric times: no dict assembly, no Python loops.
        ``prefilter`` is an optional bool mask over row slots.  Rows index ``_ids`` / ``_docs``.
        """
        raw, _ = self._validate_queries(query_vecs)
        with self._rwlock.read_lock():
            return self._engine.search(raw, int(top_k), prefilter, precision=precision or self._precision)

    @_timed("query")
    def query(
        self,
        query_vecs: np.ndarray,
        top_k: int = 10,
        better_than: Optional[float] = None,
        where: Optional[WhereT] = None,
        ids: Optional[list[Any]] = None,
        ef_search: Optional[int] = None,
        hnsw_ef_search: Optional[int] = None,
    ) -> Union[list[list[dict[str, Any]]], list[dict[str, Any]]]:
        """Exact cosine top-k.  Same results contract as the reference's NumPy path."""
        raw, is_single = self._validate_queries(query_vecs)
        num_q = raw.shape[0]
        with self._rwlock.read_lock():
            if not self._id2idx:
                return [[] for _ in range(num_q)]  # also for a single query (reference quirk Q2)
            filtered = ids is not None or where is not None
            base = top_k + self._adaptive_buffer if filtered else top_k
            dev = None
            if isinstance(where, dict) and base >= 1:
                # queries run under the READ lock, concurrently: the column index and its device mirror
                # (slot numbers, rows sent so far) are shared state, so the bookkeeping and the search
                # that relies on it are serialised among filtered queries
                with self._col_lock:
                    dev = self._device_where(where, ids)
                    if dev is not None:
                        # dict filter evaluated on the device from the code column: no host mask, no upload
                        slot, wanted, extra = dev
                        if not wanted:
                            return [[] for _ in range(num_q)]
                        scores, rows, n_cand = self._engine.search_where(
                            raw, int(base), slot, wanted, extra, precision=self._precision
                        )
            if dev is not None:
                if n_cand == 0:
                    return [[] for _ in range(num_q)]
                k_eff = min(base, n_cand)
            else:
                mask = self._candidate_mask(where, ids)
                n_cand = len(self._id2idx) if mask is None else int(np.count_nonzero(mask))
                if n_cand == 0:
                    return [[] for _ in range(num_q)]
                k_eff = min(base, n_cand)
                if k_eff < 1:
                    self._last_k_eff = int(k_eff)
                    return [[] for _ in range(num_q)]
                scores, rows = self._engine.search(raw, int(k_eff), mask, precision=self._precision)
            self._last_k_eff = int(k_eff)
            # which numpy strategy the reference would have used; the device does one fused select
            self._last_topk_strategy = (
                "argsort" if (k_eff / n_cand) > self._argsort_threshold else "argpartition"
            )
            get_doc = self._docs.getter()
            n_slots = len(self._ids)
            recheck = callable(where)
            out: list[list[dict[str, Any]]] = []
            # one bulk conversion to Python numbers for the whole batch (tolist() of a float32 array yields
            # Python floats: the `float(score)` of the reference's records, pico_vdb.py:770)
            for q_rows, q_scores in zip(rows.tolist(), scores.tolist()):
                hits: list[dict[str, Any]] = []
                for row, score in zip(q_rows, q_scores):
                    if row < 0 or row >= n_slots:
                        continue
                    doc = get_doc(row)
                    if doc is None:
                        continue
                    if better_than is not None and score < better_than:
                        continue
                    if recheck and not where(doc):  # type: ignore[operator]
                        continue
                    hits.append({**doc, K_METRICS: score})
                    if len(hits) == top_k:
                        break
                out.append(hits)
        return out[0] if is_single else out

    def query_one(
        self,
        query_vec: np.ndarray,
        top_k: int = 10,
        better_than: Optional[float] = None,
        where: Optional[Callable[[dict[str, Any]], bool]] = None,
        ids: Optional[list[Any]] = None,
        ef_search: Optional[int] = None,
        hnsw_ef_search: Optional[int] = None,
    ) -> list[dict[str, Any]]:
        return self.query(  # type: ignore[return-value]
            query_vec, top_k=top_k, better_than=better_than, where=where, ids=ids,
            ef_search=ef_search, hnsw_ef_search=hnsw_ef_search,
        )

    # ------------------------------------------------------------------ maintenance
    def stats(self) -> dict[str, Any]:
        with self._rwlock.read_lock():
            active = len(self._id2idx)
            total = len(self._ids)
            sizes = {}
            for fn in (_ids_path, _meta_path, _vecs_path, _vecs16_path):
                p = fn(self._path)
                try:
                    if os.path.exists(p):
                        sizes[os.path.basename(p)] = os.path.getsize(p)
                except OSError:
                    pass
            return {
                "active": active,
                "deleted": total - active,
                "total": total,
                "dim": self.dim,
                "faiss": False,
                "memmap": self._use_memmap,
                "file_sizes": sizes,
                "device": self._device,
                "precision": self._precision,
            }

    def vacuum(self) -> None:
        """Drop deleted slots: device-side row compaction + remapped host lists."""
        with self._rwlock.write_lock():
            if not self._free:
                return
            keep = self._id2idx.sorted_rows()
            self._engine.compact(keep)
            self._invalidate()
            self._ids = self._ids.take_sorted(keep)      # runs of bulk rows stay implicit ranges
            self._docs = self._docs.take_sorted(keep)
            for col in self._columns.values():
                col.reindex(keep)
            self._id2idx = IdMap()
            explicit_rows = []
            for r0, n, id0 in self._ids.implicit_ranges():
                self._id2idx.add_range(id0, n, r0)
            for row, _id in self._ids.explicit_items():
                self._id2idx[_id] = row
                explicit_rows.append(row)
            self._active_explicit = np.asarray(explicit_rows, dtype=np.int64)
            self._free = []

    def rebuild_index(self) -> None:
        """There is no ANN index to rebuild on the exact path; kept for API compatibility."""
        with self._rwlock.write_lock():
            return None

    # ------------------------------------------------------------------ getters
    def _record(self, row: int, _id: Any, include_vector: bool) -> dict[str, Any]:
        rec = dict(self._docs[row] or {K_ID: _id})
        if include_vector:
            rec[K_VECTOR] = self._engine.fetch_rows(np.asarray([row], dtype=np.int64))[0].copy()
        return rec

    def get(self, ids: Union[Any, list[Any]], include_vector: bool = False):
        with self._rwlock.read_lock():
            if isinstance(ids, str):
                row = self._id2idx.get(ids)
                return None if row is None else self._record(row, ids, include_vector)
            found = [(i, self._id2idx.get(i)) for i in ids]
            found = [(i, r) for i, r in found if r is not None]
            recs = [dict(self._docs[r] or {K_ID: i}) for i, r in found]
            if include_vector and found:
                vecs = self._engine.fetch_rows(np.asarray([r for _, r in found], dtype=np.int64))
                for rec, v in zip(recs, vecs):
                    rec[K_VECTOR] = v.copy()
            return recs

    def get_by_id(self, sid: Any, include_vector: bool = False) -> Optional[dict[str, Any]]:
        warnings.warn(
            "get_by_id() is deprecated: use get(id) or get([ids])", DeprecationWarning, stacklevel=2
        )
        return self.get(sid, include_vector=include_vector)  # type: ignore[return-value]

    def get_all(self, include_vector: bool = False, include_deleted: bool = False) -> list[dict[str, Any]]:
        with self._rwlock.read_lock():
            vecs = self._vectors if include_vector else None
            out: list[dict[str, Any]] = []
            if include_deleted:
                for row, (_id, doc) in enumerate(zip(self._ids, self._docs)):
                    if doc is None:
                        out.append({K_ID: _id})
                        continue
                    rec = dict(doc)
                    rec[K_ID] = _id
                    if vecs is not None:
                        rec[K_VECTOR] = vecs[row].copy()
                    out.append(rec)
                return out
            for row in self._active_indices.tolist():
                _id, doc = self._ids[row], self._docs[row]
                if _id is None or doc is None:
                    continue
                rec = dict(doc)
                rec[K_ID] = _id
                if vecs is not None:
                    rec[K_VECTOR] = vecs[row].copy()
                out.append(rec)
            return out

    def close(self) -> None:
        """Release the device store (also happens on garbage collection)."""
        eng = getattr(self, "_engine", None)
        if eng is not None and hasattr(eng, "close"):
            eng.close()
