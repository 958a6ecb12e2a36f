"""CPU oracle for the exact brute-force cosine top-k path of wensheng/picovdb.

TEST INFRASTRUCTURE ONLY.  This module restates, in plain numpy, the algorithm of the
reference's ``PicoVectorDB.query()`` NumPy path and the store it scans.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs may
import it -- and there only as the checker or the timed CPU baseline, never as the product.
``picovdb_b200`` never imports this package: the product path is CUDA-only and raises when its
extension or a GPU is missing.

Parity status: PINNED.  ``tests/test_oracle_golden.py`` checks every function below against
  * the reference's own result-pinning tests, restated with the same seeds
    (tests/test_task20_argsort_vs_argpartition.py:12-36, tests/test_more.py:133-155,
     tests/test_task5_zero_vector_normalization.py:7-41,
     tests/test_task2_numpy_query_active_indices.py:6-41), and
  * fixtures produced by importing the unmodified reference class in the build container
    (tests/golden/make_golden.py -> tests/golden/*.npz).

The arithmetic of the path lives in a third-party dependency that is not vendored under the
reference tree: numpy (pyproject.toml:19, unpinned "*"; 2.3.5 installed here) -> OpenBLAS
sgemv/sgemm, ``np.argpartition``, ``np.argsort``.  The call sites restated here are
picovdb/pico_vdb.py:58-68, 584-591, 683-714 and 753-775.

Deliberate divergence: the reference maps fast-path columns through
``_active_indices`` in *insertion* order (pico_vdb.py:686 vs :714) which returns wrong ids when
that array is not the identity.  The oracle always scores ``V[candidates]`` with ``candidates``
sorted ascending, i.e. it returns the true row owner.  All golden fixtures use stores whose
``_active_indices`` is sorted, where both agree.
"""

from __future__ import annotations

from typing import Any, Callable, Optional, Sequence, Union

import numpy as np

Float = np.float32
ADAPTIVE_BUFFER = 32  # pico_vdb.py:30
ARGSORT_THRESHOLD = 0.2  # pico_vdb.py:163
K_ID = "_id_"
K_VECTOR = "_vector_"
K_METRICS = "_metrics_"


# ----------------------------------------------------------------------------- store side
def normalize(v: np.ndarray) -> np.ndarray:
    """L2-normalise one vector in f32; the zero vector maps to e0 (pico_vdb.py:58-68)."""
    vec = np.asarray(v, dtype=Float)
    n = float(np.linalg.norm(vec))
    if n == 0.0:
        out = np.zeros_like(vec, dtype=Float)
        if out.size:
            out.flat[0] = Float(1.0)
        return out
    return (vec / n).astype(Float, copy=False)


def normalize_rows(mat: np.ndarray) -> np.ndarray:
    """Row-wise :func:`normalize` (what upsert does item by item, pico_vdb.py:413-422)."""
    mat = np.ascontiguousarray(mat, dtype=Float)
    out = np.empty_like(mat)
    for i in range(mat.shape[0]):
        out[i] = normalize(mat[i])
    return out


def normalize_rows_fast(mat: np.ndarray) -> np.ndarray:
    """Vectorised row normalisation for large synthetic stores (same zero rule).

    Not bit-identical to :func:`normalize` (different summation order inside numpy); used only
    to build big synthetic matrices where the tolerance, not bit equality, is the bar.
    """
    mat = np.ascontiguousarray(mat, dtype=Float)
    norms = np.sqrt(np.einsum("ij,ij->i", mat, mat, dtype=np.float32))[:, None]
    zero = norms[:, 0] == 0
    out = mat / np.where(zero[:, None], Float(1.0), norms)
    if zero.any():
        out[zero] = 0
        out[zero, 0] = 1.0
    return out.astype(Float, copy=False)


# ----------------------------------------------------------------------------- query side
def prepare_queries(query_vecs: np.ndarray, dim: int) -> tuple[np.ndarray, bool]:
    """Validate + normalise queries (pico_vdb.py:564-591).  Returns ((Q, dim) f32, is_single)."""
    raw = np.ascontiguousarray(query_vecs, dtype=Float)
    if raw.ndim == 1:
        if raw.shape[0] != dim:
            raise ValueError(
                f"query vector dim mismatch: expected {dim}, got {raw.shape[0]}"
            )
        is_single = True
    elif raw.ndim == 2:
        if raw.shape[1] != dim:
            raise ValueError(
                f"query vectors dim mismatch: expected last dim {dim}, got {raw.shape[1]}"
            )
        is_single = False
    else:
        raise ValueError(
            f"query expects 1D or 2D array with last dim {dim}; got shape {tuple(raw.shape)}"
        )
    vecs = raw[None, :] if is_single else raw
    norms = np.linalg.norm(vecs, axis=1, keepdims=True)
    zero_mask = norms.squeeze(-1) == 0
    if np.any(zero_mask):
        vecs = vecs.copy()
        vecs[zero_mask] = 0
        vecs[zero_mask, 0] = 1.0
        norms = np.where(zero_mask[:, None], 1.0, norms)
    vecs = (vecs / norms).astype(Float, copy=False)
    return vecs, is_single


def scores(qn: np.ndarray, vectors: np.ndarray, candidates: Optional[np.ndarray]) -> np.ndarray:
    """S = Qn . V^T (fast path, pico_vdb.py:686) or Qn . V[cand]^T (gather path, :688-689)."""
    if candidates is None:
        return qn @ vectors.T
    return qn @ vectors[candidates].T


def topk_desc(
    scores_act: np.ndarray,
    k_eff: int,
    argsort_threshold: float = ARGSORT_THRESHOLD,
) -> tuple[np.ndarray, np.ndarray, str]:
    """Per row: the k_eff largest scores, descending, with their local column index.

    Restates pico_vdb.py:698-713: full ``argsort`` when k_eff is a large fraction of the
    candidates, else ``argpartition`` + sort of the k slice.  Returns (scores, local_idx, strategy).
    """
    ncand = scores_act.shape[1]
    frac = k_eff / ncand if ncand > 0 else 0.0
    if frac > argsort_threshold:
        order_full = np.argsort(-scores_act, axis=1)[:, :k_eff]
        return np.take_along_axis(scores_act, order_full, axis=1), order_full, "argsort"
    idxs_part = np.argpartition(scores_act, -k_eff, axis=1)[:, -k_eff:]
    scores_part = np.take_along_axis(scores_act, idxs_part, axis=1)
    order = np.argsort(-scores_part, axis=1)
    return (
        np.take_along_axis(scores_part, order, axis=1),
        np.take_along_axis(idxs_part, order, axis=1),
        "argpartition",
    )


def search(
    vectors: np.ndarray,
    queries_normalised: np.ndarray,
    k: int,
    active: Optional[np.ndarray] = None,
    prefilter: Optional[np.ndarray] = None,
    argsort_threshold: float = ARGSORT_THRESHOLD,
) -> tuple[np.ndarray, np.ndarray]:
    """Array-level restatement of the hot path (pico_vdb.py:683-714).

    vectors            (N, dim) f32, rows unit-norm (deleted rows all-zero)
    queries_normalised (Q, dim) f32 from :func:`prepare_queries`
    active / prefilter optional boolean (N,) row masks; candidates = rows where both are true
    Returns (scores (Q, k) f32 descending, rows (Q, k) int64); short rows are padded with
    -inf / -1 when fewer than k candidates exist (the C-ABI's output convention).
    """
    n = vectors.shape[0]
    nq = queries_normalised.shape[0]
    mask = None
    if active is not None:
        mask = np.asarray(active, dtype=bool)
    if prefilter is not None:
        pf = np.asarray(prefilter, dtype=bool)
        mask = pf if mask is None else (mask & pf)
    out_s = np.full((nq, k), -np.inf, dtype=Float)
    out_r = np.full((nq, k), -1, dtype=np.int64)
    if mask is None or bool(mask.all()):
        cand = None
        ncand = n
    else:
        cand = np.flatnonzero(mask).astype(np.int64)
        ncand = cand.size
    if ncand == 0 or k <= 0:
        return out_s, out_r
    s = scores(queries_normalised, vectors, cand)
    k_eff = min(k, ncand)
    top_s, top_local, _ = topk_desc(s, k_eff, argsort_threshold)
    rows = top_local.astype(np.int64) if cand is None else cand[top_local]
    out_s[:, :k_eff] = top_s
    out_r[:, :k_eff] = rows
    return out_s, out_r


def search_chunked(
    vectors: np.ndarray,
    queries_normalised: np.ndarray,
    k: int,
    active: Optional[np.ndarray] = None,
    prefilter: Optional[np.ndarray] = None,
    chunk_rows: int = 262144,
) -> tuple[np.ndarray, np.ndarray]:
    """Same result as :func:`search` without materialising the (Q, N) score matrix.

    Used for large synthetic stores where the reference itself cannot run (SURVEY.md hard part
    e).  Exact: per-chunk top-k lists are merged with a final descending sort.
    """
    n = vectors.shape[0]
    nq = queries_normalised.shape[0]
    best_s = np.full((nq, 0), -np.inf, dtype=Float)
    best_r = np.full((nq, 0), -1, dtype=np.int64)
    for r0 in range(0, n, chunk_rows):
        r1 = min(n, r0 + chunk_rows)
        a = None if active is None else active[r0:r1]
        p = None if prefilter is None else prefilter[r0:r1]
        s, r = search(vectors[r0:r1], queries_normalised, k, a, p)
        r = np.where(r >= 0, r + r0, -1)
        cs = np.concatenate([best_s, s], axis=1)
        cr = np.concatenate([best_r, r], axis=1)
        order = np.argsort(-cs, axis=1, kind="stable")[:, :k]
        best_s = np.take_along_axis(cs, order, axis=1)
        best_r = np.take_along_axis(cr, order, axis=1)
    if best_s.shape[1] < k:
        pad = k - best_s.shape[1]
        best_s = np.concatenate([best_s, np.full((nq, pad), -np.inf, Float)], axis=1)
        best_r = np.concatenate([best_r, np.full((nq, pad), -1, np.int64)], axis=1)
    best_r = np.where(np.isfinite(best_s), best_r, -1)
    return best_s, best_r


def merge_topk(
    shard_scores: Sequence[np.ndarray], shard_rows: Sequence[np.ndarray], k: int
) -> tuple[np.ndarray, np.ndarray]:
    """k-way merge of per-shard (Q, k) lists (rows already global). Ties: lower row first."""
    cs = np.concatenate(list(shard_scores), axis=1)
    cr = np.concatenate(list(shard_rows), axis=1)
    nq = cs.shape[0]
    out_s = np.full((nq, k), -np.inf, dtype=Float)
    out_r = np.full((nq, k), -1, dtype=np.int64)
    for qi in range(nq):
        valid = cr[qi] >= 0
        s, r = cs[qi][valid], cr[qi][valid]
        order = np.lexsort((r, -s))[:k]
        out_s[qi, : order.size] = s[order]
        out_r[qi, : order.size] = r[order]
    return out_s, out_r


# ----------------------------------------------------------------------------- comparator
def compare_topk(
    got_scores: np.ndarray,
    got_rows: np.ndarray,
    ref_scores: np.ndarray,
    ref_rows: np.ndarray,
    rtol: float,
    atol: float = 0.0,
) -> dict[str, float]:
    """The north-star parity rule.

    * every returned score within ``rtol`` (relative, plus ``atol``) of the reference score at
      the same rank;
    * ids identical at every rank whose neighbouring score gaps exceed the tolerance (ranks that
      sit inside a near-tie may swap), and for the last rank only if the gap to the first
      excluded candidate is unknown -> it is checked as a set membership instead;
    * recall@k of the id sets is reported.
    Raises AssertionError on violation; returns summary statistics otherwise.
    """
    got_scores = np.asarray(got_scores, dtype=np.float64)
    ref_scores = np.asarray(ref_scores, dtype=np.float64)
    assert got_scores.shape == ref_scores.shape, (got_scores.shape, ref_scores.shape)
    assert got_rows.shape == ref_rows.shape
    nq, k = ref_scores.shape
    fin = np.isfinite(ref_scores)
    assert np.array_equal(fin, np.isfinite(got_scores)), "padding pattern differs"
    tol = atol + rtol * np.abs(ref_scores[fin])
    err = np.abs(got_scores[fin] - ref_scores[fin])
    assert np.all(err <= tol), f"score error {err.max():.3e} exceeds tolerance (rtol={rtol})"
    mismatches = 0
    hits = 0
    total = 0
    for qi in range(nq):
        kk = int(fin[qi].sum())
        total += kk
        ref_set = set(ref_rows[qi, :kk].tolist())
        hits += len(ref_set.intersection(got_rows[qi, :kk].tolist()))
        for j in range(kk):
            if got_rows[qi, j] == ref_rows[qi, j]:
                continue
            # a differing id is only legal inside a near-tie: some reference score within the
            # tolerance band of this rank's score must belong to the id we returned, or (last
            # ranks) the returned score must itself be within tolerance of the k-th ref score.
            s = ref_scores[qi, j]
            band = atol + 2 * rtol * max(abs(s), 1e-30)
            near = np.abs(ref_scores[qi, :kk] - s) <= band
            ok = got_rows[qi, j] in set(ref_rows[qi, :kk][near].tolist())
            if not ok:
                ok = abs(got_scores[qi, j] - ref_scores[qi, kk - 1]) <= band and kk == k
            assert ok, (
                f"query {qi} rank {j}: row {got_rows[qi, j]} != {ref_rows[qi, j]} "
                f"outside a near-tie (ref score {s}, got {got_scores[qi, j]})"
            )
            mismatches += 1
    return {
        "max_abs_err": float(err.max()) if err.size else 0.0,
        "recall": hits / total if total else 1.0,
        "rank_swaps": float(mismatches),
    }


def recall_at_k(got_rows: np.ndarray, ref_rows: np.ndarray) -> float:
    hits = 0
    total = 0
    for g, r in zip(got_rows, ref_rows):
        rs = set(int(x) for x in r if x >= 0)
        total += len(rs)
        hits += len(rs.intersection(int(x) for x in g if x >= 0))
    return hits / total if total else 1.0


# ----------------------------------------------------------------------------- record level
class OracleDB:
    """Minimal record-level restatement of the reference class (numpy path only).

    Mirrors upsert (pico_vdb.py:403-472), delete (:514-531), the candidate builder (:604-658),
    ``k_eff`` (:691-697) and result assembly (:753-775), including quirk Q2 (a single query with
    no candidates returns ``[[]]``).  Free-slot reuse follows ``_free.pop()`` (:434-439).
    """

    def __init__(self, dim: int, adaptive_buffer: int = ADAPTIVE_BUFFER,
                 argsort_threshold: float = ARGSORT_THRESHOLD) -> None:
        self.dim = dim
        self.vectors = np.empty((0, dim), dtype=Float)
        self.ids: list[Any] = []
        self.docs: list[Optional[dict[str, Any]]] = []
        self.free: list[int] = []
        self.id2idx: dict[Any, int] = {}
        self.adaptive_buffer = adaptive_buffer
        self.argsort_threshold = argsort_threshold
        self.last_k_eff: Optional[int] = None
        self.last_strategy: Optional[str] = None

    def upsert(self, items: list[dict[str, Any]]) -> dict[str, list[Any]]:
        import hashlib

        report: dict[str, list[Any]] = {"update": [], "insert": []}
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
            vec = normalize(raw)
            meta = {k: v for k, v in item.items() if k != K_VECTOR}
            item_id = meta.get(K_ID)
            if item_id is None:
                item_id = hashlib.md5(vec.tobytes()).hexdigest()
            meta[K_ID] = item_id
            if item_id in self.id2idx:
                idx = self.id2idx[item_id]
                self.vectors[idx] = vec
                self.docs[idx] = meta
                report["update"].append(item_id)
                continue
            if self.free:
                idx = self.free.pop()
                self.vectors[idx] = vec
                self.ids[idx] = item_id
                self.docs[idx] = meta
            else:
                idx = len(self.ids)
                self.vectors = np.ascontiguousarray(
                    np.vstack([self.vectors, vec[None, :]]), dtype=Float
                )
                self.ids.append(item_id)
                self.docs.append(meta)
            self.id2idx[item_id] = idx
            report["insert"].append(item_id)
        return report

    def delete(self, ids: list[Any]) -> list[Any]:
        removed = []
        for _id in ids:
            idx = self.id2idx.pop(_id, None)
            if idx is not None:
                self.docs[idx] = None
                self.vectors[idx].fill(0)
                self.free.append(idx)
                removed.append(_id)
        return removed

    def active_mask(self) -> np.ndarray:
        m = np.zeros(len(self.ids), dtype=bool)
        if self.id2idx:
            m[np.fromiter(self.id2idx.values(), dtype=np.int64)] = True
        return m

    def candidate_mask(
        self,
        where: Optional[Union[dict[str, Any], Callable[[dict[str, Any]], bool]]] = None,
        ids: Optional[list[Any]] = None,
    ) -> np.ndarray:
        """Row mask equivalent of the candidate builder (pico_vdb.py:604-658)."""
        mask = self.active_mask()
        if ids is not None:
            sel = np.zeros_like(mask)
            for s in ids:
                m = self.id2idx.get(s)
                if m is not None:
                    sel[m] = True
            mask &= sel
        if where is not None:
            if isinstance(where, dict) and len(where) == 1:
                ((k, v),) = where.items()
                if isinstance(v, dict) and set(v.keys()) == {"$in"}:
                    values = set(v["$in"])
                    test = lambda d: d.get(k) in values  # noqa: E731
                else:
                    test = lambda d: d.get(k) == v  # noqa: E731
            else:
                test = where  # generic callable (pico_vdb.py:643-654)
            for i in np.flatnonzero(mask):
                if not test(self.docs[i]):
                    mask[i] = False
        return mask

    def query(
        self,
        query_vecs: np.ndarray,
        top_k: int = 10,
        better_than: Optional[float] = None,
        where: Optional[Union[dict[str, Any], Callable[[dict[str, Any]], bool]]] = None,
        ids: Optional[list[Any]] = None,
    ):
        qn, is_single = prepare_queries(query_vecs, self.dim)
        nq = qn.shape[0]
        if not self.id2idx:
            return [[] for _ in range(nq)]
        cand_mask = self.candidate_mask(where, ids)
        ncand = int(cand_mask.sum())
        if ncand == 0:
            return [[] for _ in range(nq)]
        base = top_k + self.adaptive_buffer if (ids is not None or where is not None) else top_k
        k_eff = min(base, ncand)
        self.last_k_eff = int(k_eff)
        frac = k_eff / ncand
        self.last_strategy = "argsort" if frac > self.argsort_threshold else "argpartition"
        s, r = search(self.vectors, qn, k_eff, None, cand_mask, self.argsort_threshold)
        where_callable = callable(where)
        out = []
        for qi in range(nq):
            results: list[dict[str, Any]] = []
            for idx, score in zip(r[qi], s[qi]):
                if idx < 0 or idx >= len(self.ids):
                    continue
                doc = self.docs[idx]
                if doc is None:
                    continue
                if better_than is not None and score < better_than:
                    continue
                if where_callable and not where(doc):  # type: ignore[misc]
                    continue
                results.append({**doc, K_METRICS: float(score)})
                if len(results) == top_k:
                    break
            out.append(results)
        return out[0] if is_single else out
