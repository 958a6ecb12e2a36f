/*
 * picovdb_b200 -- C ABI of the B200-native exact cosine top-k engine.
 *
 * This header is the drop-in boundary for the one hot path of wensheng/picovdb that this
 * repository rebuilds: the NumPy branch of PicoVectorDB.query() and the store it scans.
 * The reference has NO FFI of its own (it is a single pure-Python module), so every entry
 * point below cites the reference statement(s) it replaces (file:line, relative to the
 * reference tree); INTEGRATION.md shows the ctypes stub a reference maintainer would add.
 *
 * Conventions
 *   - plain pointers and sizes only; no torch / numpy types.
 *   - every function returns PVDB_OK (0) or a negative PVDB_ERR_* code; the thread-local
 *     message of the last failure is available from pvdb_last_error().
 *   - "host" entry points take host buffers, run on the store's own stream and return when the
 *     results are in the caller's buffers.  "_dev" entry points take device pointers plus a
 *     cudaStream_t (passed as void*; NULL is CUDA's legacy default stream) and only enqueue work;
 *     the library orders them against earlier calls on other streams with an event.
 *   - matrices are row-major; row indices are int64; outputs are sorted by (score descending,
 *     row ascending) and padded with score = -inf, row = -1 when fewer than k candidates exist.
 *   - bitmaps are uint32 words, bit (r & 31) of word (r >> 5) describes row r.
 *   - a handle serialises its own calls with an internal mutex; device memory is owned by the
 *     library.  There is no CPU fallback: every call fails with PVDB_ERR_CUDA without a GPU.
 */
#ifndef PICOVDB_B200_H
#define PICOVDB_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PVDB_ABI_VERSION 1

#define PVDB_OK 0
#define PVDB_ERR_INVALID (-1)     /* bad argument */
#define PVDB_ERR_CUDA (-2)        /* CUDA runtime / driver failure (incl. no device) */
#define PVDB_ERR_OOM (-3)         /* device or pinned allocation failed */
#define PVDB_ERR_UNSUPPORTED (-4) /* requested precision / layout not available on this store */
#define PVDB_ERR_CAPACITY (-5)    /* fixed-capacity store is full (pico_vdb.py:441-442) */

/* store flags (pvdb_store_create) */
#define PVDB_STORE_F32 0x1            /* keep the fp32 matrix (reference layout, pico_vdb.py:136) */
#define PVDB_STORE_BF16 0x2           /* keep a bf16 mirror of every row */
#define PVDB_STORE_FIXED_CAPACITY 0x4 /* never grow past reserve_rows (capacity=, pico_vdb.py:286-296) */

/* scoring precision (low byte of the search flags) */
#define PVDB_PREC_AUTO 0 /* F32 scan for few queries, tensor-core batch path otherwise */
#define PVDB_PREC_F32 1  /* fp32 CUDA-core scan of the fp32 matrix (exact path) */
#define PVDB_PREC_TF32 2 /* tcgen05 kind::tf32 over the fp32 matrix, fp32 re-scoring of candidates */
#define PVDB_PREC_BF16 3 /* bf16 mirror: fp32-accumulated scan (1 query) / tcgen05 kind::f16 (batch) */
#define PVDB_PREC_MASK 0xff
/* search flags */
#define PVDB_SEARCH_QUERIES_NORMALIZED 0x100 /* skip the query L2-normalisation (pico_vdb.py:584-591) */
#define PVDB_SEARCH_NO_RESCORE 0x200         /* tensor-core paths: return the low-precision scores */
#define PVDB_SEARCH_SCAN_ONLY 0x400          /* answer every query with its own HBM scan pass (no batching) */
#define PVDB_SEARCH_NO_GUARD 0x800           /* tensor-core paths: skip the exactness guard (see pvdb_search) */

typedef struct pvdb_store pvdb_store_t;
typedef struct pvdb_exchange pvdb_exchange_t; /* one GPU's end of the cross-GPU top-k exchange */
typedef struct pvdb_group pvdb_group_t;       /* a row-sharded store over several GPUs of one process */
#define PVDB_IPC_HANDLE_BYTES 64

typedef struct pvdb_store_info {
  int32_t dim;          /* embedding dimension */
  int32_t ld_f32;       /* row stride of the fp32 matrix, in elements (dim rounded up to 4) */
  int32_t ld_bf16;      /* row stride of the bf16 mirror, in elements (dim rounded up to 8) */
  int32_t flags;        /* PVDB_STORE_* */
  int32_t device;       /* CUDA ordinal */
  int32_t reserved;
  int64_t rows;         /* slots in use (high-water mark) == len(_ids) in the reference */
  int64_t capacity;     /* slots allocated on the device */
  int64_t active;       /* rows whose active bit is set == len(_id2idx) */
  int64_t row_base;     /* added to every row index written by a search (shard offset) */
  uint64_t device_bytes; /* bytes of HBM held by this store */
} pvdb_store_info_t;

/* ---- library ---------------------------------------------------------------------------- */
int pvdb_abi_version(void);
/* message of the last failing call on this thread ("" if none) */
const char* pvdb_last_error(void);
/* number of visible CUDA devices; PVDB_ERR_CUDA if the runtime cannot initialise */
int pvdb_device_count(int* out_count);

/* ---- store: replaces the `_vectors` matrix + `_active_indices` (pico_vdb.py:136,143) ----- */
/* Allocate a store on `device` for `dim`-float rows with room for `reserve_rows` rows. */
int pvdb_store_create(pvdb_store_t** out, int device, int dim, int64_t reserve_rows, int flags);
int pvdb_store_destroy(pvdb_store_t* s);
/* Grow the allocation to at least `rows` slots (replaces the np.vstack realloc, pico_vdb.py:451-462). */
int pvdb_store_reserve(pvdb_store_t* s, int64_t rows);
int pvdb_store_info(pvdb_store_t* s, pvdb_store_info_t* out);
/* Shard offset added to result rows (multi-GPU row sharding; no reference counterpart). */
int pvdb_store_set_row_base(pvdb_store_t* s, int64_t row_base);

/* Fused L2-normalise + scatter + set-active: row rows[i] := vecs[i] / ||vecs[i]||, the zero vector
 * becomes e0 (replaces _normalize + the row write of upsert, pico_vdb.py:58-68, 422, 430, 436,
 * 451-462, 466-472).  vecs is n x dim dense fp32; rows must be unique within one call; rows past
 * the current high-water mark extend the store (slots in between stay inactive and zero). */
int pvdb_store_upsert(pvdb_store_t* s, const float* vecs, const int64_t* rows, int64_t n);
int pvdb_store_upsert_dev(pvdb_store_t* s, const float* d_vecs, const int64_t* d_rows, int64_t n,
                          int64_t max_row, void* stream);
/* Same, for n consecutive rows row0 .. row0+n-1 (bulk ingest; no row array needed). */
int pvdb_store_upsert_range(pvdb_store_t* s, const float* vecs, int64_t row0, int64_t n);
int pvdb_store_upsert_range_dev(pvdb_store_t* s, const float* d_vecs, int64_t row0, int64_t n,
                                void* stream);

/* Clear the active bit and zero-fill the rows (replaces pico_vdb.py:523 + :528-531). */
int pvdb_store_delete(pvdb_store_t* s, const int64_t* rows, int64_t n);

/* Copy rows out as n x dim dense fp32 (replaces `self._vectors[idx].copy()`, pico_vdb.py:945). */
int pvdb_store_fetch(pvdb_store_t* s, const int64_t* rows, int64_t n, float* out);
/* Copy rows row0 .. row0+n-1 out as dense fp32 (feeds np.save in save(), pico_vdb.py:356). */
int pvdb_store_download(pvdb_store_t* s, int64_t row0, int64_t n, float* out);
/* Raw load of already-normalised rows (the np.load of _load_or_init, pico_vdb.py:233-237): no
 * normalisation; `active_bits` has ceil(n/32) words for rows row0.. (row0 % 32 == 0), NULL = all
 * active (rebuild of _active_indices, pico_vdb.py:247-259). */
int pvdb_store_upload(pvdb_store_t* s, int64_t row0, int64_t n, const float* vecs,
                      const uint32_t* active_bits);
/* The bf16 mirror as stored: n x dim bf16 bit patterns.  Persists a bf16-only store at half the size of
 * its fp32 expansion; upload_bf16 is the matching raw load (bf16-only stores). */
int pvdb_store_download_bf16(pvdb_store_t* s, int64_t row0, int64_t n, uint16_t* out);
int pvdb_store_upload_bf16(pvdb_store_t* s, int64_t row0, int64_t n, const uint16_t* vecs,
                           const uint32_t* active_bits);
/* save() of a large store (pico_vdb.py:356): rows [row0, row0+n) are written into the EXISTING file
 * `path` at byte `file_offset` -- dense fp32 rows, or (as_bf16) the mirror's bit patterns -- device->host
 * DMA overlapped with pwrite() from pinned buffers.  The caller creates the file with its .npy header
 * and full size first; ranks of a sharded store write their own row ranges concurrently. */
int pvdb_store_write_file(pvdb_store_t* s, const char* path, int64_t file_offset, int64_t row0, int64_t n,
                          int as_bf16);
/* Copy the active bitmap out: ceil(rows/32) words. */
int pvdb_store_active_bits(pvdb_store_t* s, uint32_t* out_words);
/* Compaction: new row i := old row keep_rows[i] (ascending), all n kept rows active, rows := n
 * (replaces the fancy-index copy of vacuum(), pico_vdb.py:840-848). */
int pvdb_store_compact(pvdb_store_t* s, const int64_t* keep_rows, int64_t n);

/* ---- metadata columns: on-device evaluation of the dict `where` prefilters ---------------- */
/* The reference evaluates {key: value} / {key: {"$in": [...]}} filters with a Python loop over the
 * candidate docs on every query (pico_vdb.py:615-638).  Here the host dictionary-encodes a metadata
 * key once (value -> int32 code >= 0, -1 = absent) and keeps the codes in a device column; a filter
 * is then one kernel that turns (column, wanted codes) into the prefilter bitmap on the device.
 * Write codes for rows[i] (rows == NULL: the n consecutive rows from row0).  Columns are numbered
 * 0..15; rows never written read as "absent"; compaction drops all columns (re-upload them). */
int pvdb_store_column_write(pvdb_store_t* s, int column, const int64_t* rows, int64_t row0,
                            const int32_t* codes, int64_t n);
int pvdb_store_column_drop(pvdb_store_t* s, int column);
/* pvdb_search with the prefilter "column value in wanted[0..n_wanted)", optionally ANDed with a host
 * bitmap (extra_bits, e.g. an `ids=` restriction).  *out_candidates receives the number of rows that
 * were eligible (active, matching, in extra_bits) -- the reference's len(candidate_idx). */
int pvdb_search_where(pvdb_store_t* s, const float* queries, int64_t nq, int k, int column,
                      const int32_t* wanted, int n_wanted, const uint32_t* extra_bits, int flags,
                      float* out_scores, int64_t* out_rows, int64_t* out_candidates);

/* ---- search: replaces pico_vdb.py:584-591 + :683-714 -------------------------------------- */
/* For each of nq queries (nq x dim fp32): normalise (zero -> e0) unless flagged, score every row
 * whose active bit -- and prefilter bit, when prefilter_bits != NULL (ceil(rows/32) words) -- is
 * set, and write the k best as out_scores[nq*k] / out_rows[nq*k]. */
int pvdb_search(pvdb_store_t* s, const float* queries, int64_t nq, int k,
                const uint32_t* prefilter_bits, int flags, float* out_scores, int64_t* out_rows);
/* Exactness of the tensor-core precisions (TF32 / BF16 batches).  The tensor-core pass only RANKS
 * rows; the best k + slack are re-scored with the exact arithmetic of the single-query scan.  A
 * guard then proves, per query, that no row outside the kept candidates can belong to the top k
 * (re-scored k-th best > weakest kept low-precision score + input-rounding bound for unit vectors).
 * Queries that cannot be proven (dense near-ties, e.g. near-duplicate corpora) are answered again by
 * the exact scan, so a query returns the same ids alone and inside a batch -- the id rule of the
 * reference's numpy path (pico_vdb.py:699-714).  The guard costs one stream synchronisation per
 * call (also in pvdb_search_dev); PVDB_SEARCH_NO_GUARD skips it.  This call reports how many queries
 * of the last search on this store / of all searches fell back to the exact scan. */
int pvdb_store_guard_stats(pvdb_store_t* s, int64_t* out_last, int64_t* out_total);
int pvdb_search_dev(pvdb_store_t* s, const float* d_queries, int64_t nq, int k,
                    const uint32_t* d_prefilter_bits, int flags, float* d_out_scores,
                    int64_t* d_out_rows, void* stream);

/* ---- multi-GPU: row shards + peer-memory exchange (no reference counterpart) ---
 * The database rows are split into contiguous shards, one pvdb_store_t per GPU with row_base = the
 * shard's first global row.  Every GPU answers a query from its shard; the per-GPU top-k lists are
 * then exchanged and merged.  The exchange does not go through a collective library: every GPU owns
 * a mailbox in its HBM that all peers can write over NVLink (CUDA IPC between processes, peer access
 * inside one process), and the kernel that finishes a local list stores it into every peer's
 * mailbox, raises a flag, waits for the peers' flags and merges -- inside the scan kernel for single
 * queries, in one extra kernel after a tensor-core batch.  Every GPU ends with the same final top k.
 *
 * One exchange end per GPU, one GPU per process-or-thread: kernels that wait for each other must not
 * share a GPU.  All ranks must make the same pvdb_search_exchange* calls in the same order.
 * slot_keys = the largest nq * k one call may produce; k <= 128 (larger k: gather the per-shard
 * results and use pvdb_merge_topk_dev).  The wait is bounded (~4 s), then the kernel traps. */
int pvdb_exchange_create(pvdb_exchange_t** out, int device, int world, int rank, int64_t slot_keys);
int pvdb_exchange_destroy(pvdb_exchange_t* ex);
/* Between processes: export this end's mailbox (PVDB_IPC_HANDLE_BYTES bytes), all-gather the handles
 * by any means, then connect with the `world` handles in rank order. */
int pvdb_exchange_ipc_handle(pvdb_exchange_t* ex, void* out_handle);
int pvdb_exchange_connect_ipc(pvdb_exchange_t* ex, const void* handles);
/* Unmap the peers' mailboxes.  Tear-down order between processes: every rank disconnects, a barrier,
 * then every rank destroys (a mailbox must not be freed while a peer still maps it). */
int pvdb_exchange_disconnect(pvdb_exchange_t* ex);
/* Inside one process: connect the `world` ends exs[0..world) (distinct devices, peer access). */
int pvdb_exchange_connect_local(pvdb_exchange_t** exs, int world);
int pvdb_exchange_info(pvdb_exchange_t* ex, int* out_world, int* out_rank, int64_t* out_slot_keys,
                       int64_t* out_launches);
/* pvdb_search / pvdb_search_dev on one shard with the exchange fused in: the outputs are the merged
 * top k over all shards (global rows). */
int pvdb_search_exchange(pvdb_store_t* s, pvdb_exchange_t* ex, const float* queries, int64_t nq, int k,
                         const uint32_t* prefilter_bits, int flags, float* out_scores, int64_t* out_rows);
int pvdb_search_exchange_dev(pvdb_store_t* s, pvdb_exchange_t* ex, const float* d_queries, int64_t nq,
                             int k, const uint32_t* d_prefilter_bits, int flags, float* d_out_scores,
                             int64_t* d_out_rows, void* stream);

/* Single-process form (the `devices=[...]` keyword of the Python class): one handle that owns a
 * shard store + exchange end per listed device (distinct devices, peer access required) and a worker
 * thread per device.  Global row r lives in shard r / rows_per_shard (contiguous blocks of
 * ceil(capacity_rows / ndev) rows rounded up to 32).  Writes go through the per-shard store handles
 * (pvdb_group_store; rows local to the shard = global row - shard * rows_per_shard); a search is ONE
 * call: every shard scans its rows, the lists are exchanged over NVLink inside the kernels, shard 0's
 * copy of the merged result is returned.  prefilter_bits is the GLOBAL bitmap (ceil(capacity/32)
 * words).  slot_keys <= 0 selects 65536 (nq * k of one call must fit). */
int pvdb_group_create(pvdb_group_t** out, const int* devices, int ndev, int dim, int64_t capacity_rows,
                      int flags, int64_t slot_keys);
int pvdb_group_destroy(pvdb_group_t* g);
int pvdb_group_size(pvdb_group_t* g, int* out_world, int64_t* out_rows_per_shard);
pvdb_store_t* pvdb_group_store(pvdb_group_t* g, int shard);
int pvdb_group_search(pvdb_group_t* g, const float* queries, int64_t nq, int k,
                      const uint32_t* prefilter_bits, int flags, float* out_scores, int64_t* out_rows);

/* k-way merge of nlists per-shard results (what an all-gather of the per-GPU outputs produces)
 * into [nq][k]; rows are global already.  List l's scores start at d_scores + l*scores_stride
 * (elements) and its rows at d_rows + l*rows_stride; a stride <= 0 means contiguous (nq*k), so one
 * packed all-gather buffer of {rows, scores} blocks can be merged in place.  Device pointers. */
int pvdb_merge_topk_dev(int device, const float* d_scores, const int64_t* d_rows, int nlists,
                        int64_t nq, int k, int64_t scores_stride, int64_t rows_stride,
                        float* d_out_scores, int64_t* d_out_rows, void* stream);

/* Number of kernels this library has launched in the calling process (bench bookkeeping). */
int64_t pvdb_kernel_launches(void);

#ifdef __cplusplus
}
#endif
#endif /* PICOVDB_B200_H */
