// C ABI: search + merge entry points (the read side of the hot path).
#include <algorithm>
#include <cstdlib>
#include <cstring>

#include "batch.cuh"
#include "exchange.cuh"
#include "scan.cuh"
#include "store.cuh"

using namespace pvdb;

namespace pvdb {

__global__ void fill_empty_results_kernel(float* scores, int64_t* rows, int64_t n) {
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < n;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    scores[i] = -INFINITY;
    rows[i] = -1;
  }
}

static int fill_empty(float* d_scores, int64_t* d_rows, int64_t n, cudaStream_t st) {
  if (n == 0) return PVDB_OK;
  const int blocks = static_cast<int>(std::min<int64_t>((n + 255) / 256, kNumSMs * 8));
  fill_empty_results_kernel<<<blocks, 256, 0, st>>>(d_scores, d_rows, n);
  PVDB_LAUNCH_CHECK();
  return PVDB_OK;
}

// Per-query scan passes (exact fp32 accumulation, fp32 matrix or bf16 mirror).  k > kFusedK is
// served by paging: pass p only admits keys strictly below the last key of pass p-1, which the
// kernel keeps on the device, so no host round trip separates the passes.
// d_qn: normalised queries (nq x ldq) or NULL, in which case d_raw (nq x dim, raw) is normalised by
// the scan kernel itself.
// ex (optional, k <= kFusedK only): every launch exchanges its list with the peer GPUs inside the
// kernel and writes the merged, final top k (one exchange sequence number per query).
// h_sel (optional, host): serve only the queries h_sel[0 .. nq) of the arrays (the guard's flagged queries); every
// pointer still addresses the WHOLE query / result arrays.
static int search_scan(pvdb_store* s, bool bf16, const float* d_qn, const float* d_raw, int64_t nq, int k,
                       const uint32_t* d_pref, float* d_out_scores, int64_t* d_out_rows, cudaStream_t st,
                       pvdb_exchange* ex = nullptr, const int* h_sel = nullptr) {
  const int grid = scan_grid_blocks();
  const size_t list_bytes = static_cast<size_t>(grid) * kFusedK * sizeof(uint64_t);
  PVDB_TRY(s->d_partial.ensure(list_bytes + 256));   // ticket, paging bound (+ the phase stamps of trace builds)
  unsigned char* ctrl = static_cast<unsigned char*>(s->d_partial.ptr) + list_bytes;
  if (s->partial_gen_inited != s->d_partial.gen) {
    PVDB_CUDA(cudaMemsetAsync(ctrl, 0, 64, st));
    s->partial_gen_inited = s->d_partial.gen;
  }
  static const bool no_pdl = getenv("PVDB_SCAN_NO_PDL") != nullptr;
  ScanParams p{};
  p.matrix = bf16 ? s->bf16.ptr : s->f32.ptr;
  p.n_rows = s->rows;
  p.row_chunks = bf16 ? s->ld_bf16 / 8 : s->ld_f32 / 4;
  p.active = static_cast<const uint32_t*>(s->active.ptr);
  p.prefilter = d_pref;
  p.query_floats = s->ldq;
  p.partial = static_cast<uint64_t*>(s->d_partial.ptr);
  p.ticket = reinterpret_cast<unsigned int*>(ctrl);
  p.next_upper = reinterpret_cast<uint64_t*>(ctrl + 8);
  p.floor_key = reinterpret_cast<unsigned long long*>(ctrl + 16);
  p.row_base = s->row_base;
  p.dim = s->dim;
  p.nq = 1;
  // Several queries of one call share a pass over the matrix where the kernels allow it (k <= 32, no fused
  // exchange: scan_kernel.cuh, "several queries per pass"); a lone query, a remainder of one, large k and
  // exchanging launches take the single-query kernels.
  const int width = (ex == nullptr && nq > 1) ? scan_multi_width(bf16, s->ldq, k) : 1;
  int64_t q = 0;
  bool first = true;
  auto sel = [&](int64_t i) -> int64_t { return h_sel ? h_sel[i] : i; };
  for (; width > 1 && nq - q >= 2; q += width) {
    p.nq = static_cast<int>(std::min<int64_t>(width, nq - q));
    for (int j = 0; j < 4; ++j) p.qsel[j] = j < p.nq ? sel(q + j) : 0;
    p.query = d_qn;
    p.raw_query = d_qn ? nullptr : d_raw;
    p.k = k;
    p.upper = nullptr;
    p.out_scores = d_out_scores;
    p.out_rows = d_out_rows;
    p.xv = ExchangeView{};
    p.pdl = !first && !no_pdl;
    first = false;
    PVDB_TRY(launch_scan_multi(p, bf16, st));
  }
  p.nq = 1;
  for (; q < nq; ++q) {
    const int64_t qi = sel(q);
    p.query = d_qn ? d_qn + qi * s->ldq : nullptr;
    p.raw_query = d_qn ? nullptr : d_raw + qi * s->dim;
    for (int k0 = 0; k0 < k; k0 += kFusedK) {
      p.k = std::min(kFusedK, k - k0);
      p.upper = (k0 == 0) ? nullptr : p.next_upper;
      p.out_scores = d_out_scores + qi * k + k0;
      p.out_rows = d_out_rows + qi * k + k0;
      p.xv = ExchangeView{};
      if (ex != nullptr) p.xv = ex->next_view();
      // second and later launches of this call: the previous operation on the stream is a scan whose
      // inputs were complete before it started (PVDB_SCAN_NO_PDL=1 switches the overlap off)
      p.pdl = (!first || k0 > 0) && !no_pdl;
      first = false;
      PVDB_TRY(launch_scan(p, bf16, st));
    }
  }
  return PVDB_OK;
}

// Device-side search shared by both entry points.  d_queries is nq x dim fp32 (raw).
// ex (optional): this store is one shard of a row-sharded database; the results written are the
// MERGED top k over all shards (every rank makes the same call and ends with the same answer).
static int search_device(pvdb_store* s, const float* d_queries, int64_t nq, int k, const uint32_t* d_pref,
                         int flags, float* d_out_scores, int64_t* d_out_rows, cudaStream_t st,
                         pvdb_exchange* ex = nullptr) {
  if (nq == 0) return PVDB_OK;
  // a world of one has nobody to exchange with; PVDB_EXCHANGE_SELF=1 keeps the (self-)mailbox steps
  // in the kernels anyway so that a single-GPU box can test them
  if (ex != nullptr && ex->world <= 1 && getenv("PVDB_EXCHANGE_SELF") == nullptr) ex = nullptr;
  if (ex != nullptr) {
    if (!ex->connected) return fail(PVDB_ERR_INVALID, "search: the exchange is not connected to its peers");
    if (k > kFusedK) return fail(PVDB_ERR_UNSUPPORTED, "search: the fused exchange serves k <= %d", kFusedK);
    if (nq * k > ex->slot_keys)
      return fail(PVDB_ERR_INVALID, "search: %lld x %d results exceed the exchange slot (%lld keys)", (long long)nq, k,
                  (long long)ex->slot_keys);
  }
  int prec = flags & PVDB_PREC_MASK;
  const bool has_f32 = (s->flags & PVDB_STORE_F32) != 0;
  const bool has_b16 = (s->flags & PVDB_STORE_BF16) != 0;
  if (prec == PVDB_PREC_AUTO) {
    if (nq < kBatchMinQueries || !batch_path_available())
      prec = has_f32 ? PVDB_PREC_F32 : PVDB_PREC_BF16;
    else
      prec = has_f32 ? PVDB_PREC_TF32 : PVDB_PREC_BF16;
  }
  if ((prec == PVDB_PREC_F32 || prec == PVDB_PREC_TF32) && !(s->flags & PVDB_STORE_F32))
    return fail(PVDB_ERR_UNSUPPORTED, "search: this store keeps no fp32 matrix");
  if (prec == PVDB_PREC_BF16 && !has_b16) return fail(PVDB_ERR_UNSUPPORTED, "search: this store keeps no bf16 mirror");
  if (prec < PVDB_PREC_F32 || prec > PVDB_PREC_BF16) return fail(PVDB_ERR_INVALID, "search: unknown precision %d", prec);
  if (s->rows == 0 && ex == nullptr) return fill_empty(d_out_scores, d_out_rows, nq * k, st);

  const bool want_rescore = !(flags & PVDB_SEARCH_NO_RESCORE);
  const bool batch = batch_path_available() && nq >= kBatchMinQueries && !(flags & PVDB_SEARCH_SCAN_ONLY) &&
                     (prec == PVDB_PREC_TF32 || prec == PVDB_PREC_BF16) &&
                     k <= batch_max_k(prec == PVDB_PREC_BF16, want_rescore);
  const bool normalised = (flags & PVDB_SEARCH_QUERIES_NORMALIZED) != 0;
  const bool need16 = batch && (prec == PVDB_PREC_BF16);
  // exactness guard of the tensor-core paths (batch.cu): on unless the caller opts out or asks for
  // the raw low-precision scores
  const bool guard = batch && want_rescore && !(flags & PVDB_SEARCH_NO_GUARD);
  s->guard_flagged_last = 0;
  // With an exchange, the batch path (and an empty shard) first produce this shard's lists in
  // scratch; one exchange + merge launch then writes the final result.  The scan path exchanges
  // inside the scan kernel itself.  Every rank takes the same path (it depends on the arguments and
  // the store layout only), so all ranks consume the same exchange sequence numbers.
  // ... unless several queries can share a pass (scan_kernel.cuh): then the local lists of the whole call go to
  // scratch as well and ONE exchange + merge launch follows, instead of one exchanging scan per query.
  const bool grouped = ex != nullptr && !batch && nq >= 2 && scan_multi_width(prec == PVDB_PREC_BF16, s->ldq, k) > 1;
  float* d_fin_scores = d_out_scores;
  int64_t* d_fin_rows = d_out_rows;
  if (ex != nullptr && (batch || grouped || s->rows == 0)) {
    PVDB_TRY(s->d_xloc.ensure(static_cast<size_t>(nq) * k * (sizeof(int64_t) + sizeof(float))));
    d_out_rows = static_cast<int64_t*>(s->d_xloc.ptr);
    d_out_scores = reinterpret_cast<float*>(d_out_rows + nq * k);
  }
  if (s->rows == 0) {  // only reached with an exchange: publish empty lists, collect the peers'
    PVDB_TRY(fill_empty(d_out_scores, d_out_rows, nq * k, st));
    if (batch || grouped) return launch_exchange_merge(ex, d_out_scores, d_out_rows, nq, k, d_fin_scores, d_fin_rows, st);
    for (int64_t q = 0; q < nq; ++q)
      PVDB_TRY(launch_exchange_merge(ex, d_out_scores + q * k, d_out_rows + q * k, 1, k, d_fin_scores + q * k,
                                     d_fin_rows + q * k, st));
    return PVDB_OK;
  }
  const float* d_qn = nullptr;
  __nv_bfloat16* d_qn16 = nullptr;
  float* d_qeps = nullptr;
  if (guard) {
    PVDB_TRY(s->d_qeps.ensure(static_cast<size_t>(nq) * 4 * sizeof(float)));
    d_qeps = static_cast<float*>(s->d_qeps.ptr);
  }
  if (normalised && s->ldq == s->dim && !need16 && !guard && (!batch || nq == batch_query_rows(nq))) {
    d_qn = d_queries;  // already in the padded layout the kernels read: no preparation launch at all
  } else if (!batch && !normalised) {
    d_qn = nullptr;    // the scan kernel normalises the raw query itself (fused, no extra launch)
  } else {
    // The batch path's TMA boxes are 128 queries tall: keep the prepared queries padded with zero
    // rows to whole boxes, so no box is partly out of bounds (measured: a 16-query batch ran 27 %
    // slower than a 128-query one over the same rows because of the clipped boxes).
    const int64_t nq_pad = batch ? batch_query_rows(nq) : nq;
    PVDB_TRY(s->d_qn.ensure(static_cast<size_t>(nq_pad) * s->ldq * sizeof(float)));
    if (nq_pad > nq)
      PVDB_CUDA(cudaMemsetAsync(static_cast<float*>(s->d_qn.ptr) + static_cast<size_t>(nq) * s->ldq, 0,
                                static_cast<size_t>(nq_pad - nq) * s->ldq * sizeof(float), st));
    if (need16) {
      PVDB_TRY(s->d_qn16.ensure(static_cast<size_t>(nq_pad) * s->ldq * sizeof(__nv_bfloat16)));
      d_qn16 = static_cast<__nv_bfloat16*>(s->d_qn16.ptr);
      if (nq_pad > nq)
        PVDB_CUDA(cudaMemsetAsync(d_qn16 + static_cast<size_t>(nq) * s->ldq, 0,
                                  static_cast<size_t>(nq_pad - nq) * s->ldq * sizeof(__nv_bfloat16), st));
    }
    PVDB_TRY(launch_prepare_queries(d_queries, nq, s->dim, normalised, static_cast<float*>(s->d_qn.ptr), d_qn16,
                                    s->ldq, d_qeps, st));
    d_qn = static_cast<const float*>(s->d_qn.ptr);
  }

  if (batch) {
    unsigned* d_flag_count = nullptr;
    int* d_flag_list = nullptr;
    if (guard) {
      PVDB_TRY(s->d_flag.ensure((static_cast<size_t>(nq) + 4) * sizeof(int)));
      PVDB_TRY(s->h_flag.ensure((static_cast<size_t>(nq) + 4) * sizeof(int)));
      d_flag_count = static_cast<unsigned*>(s->d_flag.ptr);
      d_flag_list = static_cast<int*>(s->d_flag.ptr) + 4;
      PVDB_CUDA(cudaMemsetAsync(d_flag_count, 0, sizeof(unsigned), st));
    }
    PVDB_TRY(search_batch(s, prec == PVDB_PREC_BF16, d_qn, d_qn16, nq, k, d_pref, !want_rescore, d_qeps, d_flag_count,
                          d_flag_list, d_out_scores, d_out_rows, st));
    if (!guard) {
      if (ex != nullptr) return launch_exchange_merge(ex, d_out_scores, d_out_rows, nq, k, d_fin_scores, d_fin_rows, st);
      return PVDB_OK;
    }
    // The one host round trip of a guarded tensor-core search: how many queries could not be PROVEN
    // exact?  (Normally none; near-duplicate corpora flag many.)  Those are answered again by the
    // exact scan of the same store -- fp32 rows when the store has them, else the bf16 rows -- which
    // overwrites their slice of the output.
    int* h = static_cast<int*>(s->h_flag.ptr);
    PVDB_CUDA(cudaMemcpyAsync(h, d_flag_count, sizeof(unsigned), cudaMemcpyDeviceToHost, st));
    PVDB_CUDA(cudaStreamSynchronize(st));
    const int64_t n_flag = std::min<int64_t>(static_cast<unsigned>(h[0]), nq);
    s->guard_flagged_last = n_flag;
    s->guard_flagged_total += n_flag;
    if (n_flag > 0)
    PVDB_CUDA(cudaMemcpyAsync(h + 4, d_flag_list, static_cast<size_t>(n_flag) * sizeof(int), cudaMemcpyDeviceToHost, st));
    if (n_flag > 0) PVDB_CUDA(cudaStreamSynchronize(st));
    for (int64_t i = 0; i < n_flag; ++i)
      if (h[4 + i] < 0 || h[4 + i] >= nq) return fail(PVDB_ERR_CUDA, "guard: corrupt flag list");
    // the arithmetic a lone query gets (raw query normalised inside the scan kernel), several flagged queries
    // per pass over the matrix where the scan kernels allow it
    if (n_flag > 0) {
      if (normalised)
        PVDB_TRY(search_scan(s, !has_f32, d_qn, nullptr, n_flag, k, d_pref, d_out_scores, d_out_rows, st, nullptr, h + 4));
      else
        PVDB_TRY(search_scan(s, !has_f32, nullptr, d_queries, n_flag, k, d_pref, d_out_scores, d_out_rows, st, nullptr, h + 4));
    }
    if (ex != nullptr) return launch_exchange_merge(ex, d_out_scores, d_out_rows, nq, k, d_fin_scores, d_fin_rows, st);
    return PVDB_OK;
  }
  // scan path: TF32 requests with few queries are served by the exact fp32 scan
  if (grouped) {
    PVDB_TRY(search_scan(s, prec == PVDB_PREC_BF16, d_qn, d_queries, nq, k, d_pref, d_out_scores, d_out_rows, st));
    return launch_exchange_merge(ex, d_out_scores, d_out_rows, nq, k, d_fin_scores, d_fin_rows, st);
  }
  return search_scan(s, prec == PVDB_PREC_BF16, d_qn, d_queries, nq, k, d_pref, d_out_scores, d_out_rows, st, ex);
}

}  // namespace pvdb

#define PVDB_ENTER(s)                                                     \
  if ((s) == nullptr) return fail(PVDB_ERR_INVALID, "null store handle"); \
  std::lock_guard<std::mutex> _guard((s)->mu);                            \
  PVDB_CUDA(cudaSetDevice((s)->device))

static int search_dev_entry(pvdb_store_t* s, pvdb_exchange* ex, const float* d_queries, int64_t nq, int k,
                            const uint32_t* d_prefilter_bits, int flags, float* d_out_scores, int64_t* d_out_rows,
                            void* stream) {
  PVDB_ENTER(s);
  if (nq < 0 || k < 1 || (nq > 0 && (!d_queries || !d_out_scores || !d_out_rows)))
    return fail(PVDB_ERR_INVALID, "search_dev: bad arguments (nq=%lld, k=%d)", (long long)nq, k);
  if (ex != nullptr && ex->device != s->device) return fail(PVDB_ERR_INVALID, "search: exchange and store live on different devices");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  PVDB_TRY(s->use_stream(st));
  return search_device(s, d_queries, nq, k, d_prefilter_bits, flags, d_out_scores, d_out_rows, st, ex);
}

extern "C" int pvdb_search_dev(pvdb_store_t* s, const float* d_queries, int64_t nq, int k,
                               const uint32_t* d_prefilter_bits, int flags, float* d_out_scores,
                               int64_t* d_out_rows, void* stream) {
  return search_dev_entry(s, nullptr, d_queries, nq, k, d_prefilter_bits, flags, d_out_scores, d_out_rows, stream);
}

extern "C" int pvdb_search_exchange_dev(pvdb_store_t* s, pvdb_exchange_t* ex, const float* d_queries, int64_t nq,
                                        int k, const uint32_t* d_prefilter_bits, int flags, float* d_out_scores,
                                        int64_t* d_out_rows, void* stream) {
  if (!ex) return fail(PVDB_ERR_INVALID, "null exchange handle");
  return search_dev_entry(s, ex, d_queries, nq, k, d_prefilter_bits, flags, d_out_scores, d_out_rows, stream);
}

// Enqueue one host-buffer search on the store's stream (store mutex held by the caller); the results
// land in the slot's pinned buffer once `slot.done` has completed.
static int enqueue_host_search(pvdb_store* s, pvdb_store::IoSlot& slot, pvdb_exchange* ex, const float* queries,
                               int64_t nq, int k, const uint32_t* prefilter_bits, int flags) {
  cudaStream_t st = s->stream;
  PVDB_TRY(s->use_stream(st));
  const size_t q_bytes = static_cast<size_t>(nq) * s->dim * sizeof(float);
  const size_t n_out = static_cast<size_t>(nq) * k;
  // results: rows (int64) first, then scores (fp32) so both are naturally aligned in one block
  const size_t out_bytes = n_out * (sizeof(int64_t) + sizeof(float));
  PVDB_TRY(s->d_in.ensure(q_bytes));
  PVDB_TRY(s->d_out.ensure(out_bytes));
  PVDB_TRY(slot.h_res.ensure(out_bytes));
  PVDB_CUDA(cudaMemcpyAsync(s->d_in.ptr, queries, q_bytes, cudaMemcpyHostToDevice, st));
  const uint32_t* d_pref = nullptr;
  if (prefilter_bits && s->rows > 0) {
    const size_t nwords = static_cast<size_t>((s->rows + 31) >> 5);
    PVDB_TRY(s->d_prefilter.ensure(nwords * sizeof(uint32_t)));
    PVDB_CUDA(cudaMemcpyAsync(s->d_prefilter.ptr, prefilter_bits, nwords * sizeof(uint32_t), cudaMemcpyHostToDevice, st));
    d_pref = static_cast<const uint32_t*>(s->d_prefilter.ptr);
  }
  // Small result sets are written by the kernels straight into the pinned host buffer (pinned
  // memory is device-addressable under UVA), which removes the D2H copy from the latency of a
  // single query; large ones go through HBM and one bulk copy.
  const bool zero_copy = out_bytes <= (64u << 10);
  int64_t* d_rows = static_cast<int64_t*>(zero_copy ? slot.h_res.ptr : s->d_out.ptr);
  float* d_scores = reinterpret_cast<float*>(d_rows + n_out);
  PVDB_TRY(search_device(s, static_cast<const float*>(s->d_in.ptr), nq, k, d_pref, flags, d_scores, d_rows, st, ex));
  if (!zero_copy) PVDB_CUDA(cudaMemcpyAsync(slot.h_res.ptr, s->d_out.ptr, out_bytes, cudaMemcpyDeviceToHost, st));
  PVDB_CUDA(cudaEventRecord(slot.done, st));
  return PVDB_OK;
}

static int search_host_entry(pvdb_store_t* s, pvdb_exchange* ex, const float* queries, int64_t nq, int k,
                             const uint32_t* prefilter_bits, int flags, float* out_scores, int64_t* out_rows) {
  if (s == nullptr) return fail(PVDB_ERR_INVALID, "null store handle");
  if (ex != nullptr && ex->device != s->device) return fail(PVDB_ERR_INVALID, "search: exchange and store live on different devices");
  if (nq < 0 || k < 1 || (nq > 0 && (!queries || !out_scores || !out_rows)))
    return fail(PVDB_ERR_INVALID, "search: bad arguments (nq=%lld, k=%d)", (long long)nq, k);
  if (nq == 0) return PVDB_OK;
  const int slot_id = s->acquire_io_slot();  // ours until the results have been copied out
  pvdb_store::IoSlot& slot = s->io[slot_id];
  int rc;
  {
    std::lock_guard<std::mutex> guard(s->mu);   // held while enqueueing only
    cudaError_t e = cudaSetDevice(s->device);
    rc = e == cudaSuccess ? enqueue_host_search(s, slot, ex, queries, nq, k, prefilter_bits, flags)
                          : fail(PVDB_ERR_CUDA, "cudaSetDevice failed: %s", cudaGetErrorString(e));
    // a call that failed half way may have work in flight that still writes into the slot
    if (rc != PVDB_OK && e == cudaSuccess) (void)cudaStreamSynchronize(s->stream);
  }
  if (rc == PVDB_OK) {
    cudaError_t e = cudaEventSynchronize(slot.done);
    if (e != cudaSuccess) {
      (void)cudaGetLastError();
      rc = fail(PVDB_ERR_CUDA, "search failed on the device: %s", cudaGetErrorString(e));
    }
  }
  if (rc == PVDB_OK) {
    const size_t n_out = static_cast<size_t>(nq) * k;
    const int64_t* h_rows = static_cast<const int64_t*>(slot.h_res.ptr);
    std::memcpy(out_rows, h_rows, n_out * sizeof(int64_t));
    std::memcpy(out_scores, h_rows + n_out, n_out * sizeof(float));
  }
  s->release_io_slot(slot_id);
  return rc;
}

extern "C" int pvdb_search(pvdb_store_t* s, const float* queries, int64_t nq, int k,
                           const uint32_t* prefilter_bits, int flags, float* out_scores, int64_t* out_rows) {
  return search_host_entry(s, nullptr, queries, nq, k, prefilter_bits, flags, out_scores, out_rows);
}

extern "C" int pvdb_search_exchange(pvdb_store_t* s, pvdb_exchange_t* ex, const float* queries, int64_t nq, int k,
                                    const uint32_t* prefilter_bits, int flags, float* out_scores, int64_t* out_rows) {
  if (!ex) return fail(PVDB_ERR_INVALID, "null exchange handle");
  return search_host_entry(s, ex, queries, nq, k, prefilter_bits, flags, out_scores, out_rows);
}

extern "C" int pvdb_search_where(pvdb_store_t* s, const float* queries, int64_t nq, int k, int column,
                                 const int32_t* wanted, int n_wanted, const uint32_t* extra_bits, int flags,
                                 float* out_scores, int64_t* out_rows, int64_t* out_candidates) {
  PVDB_ENTER(s);
  if (nq < 0 || k < 1 || (nq > 0 && (!queries || !out_scores || !out_rows)))
    return fail(PVDB_ERR_INVALID, "search_where: bad arguments (nq=%lld, k=%d)", (long long)nq, k);
  if (out_candidates) *out_candidates = 0;
  if (nq == 0) return PVDB_OK;
  cudaStream_t st = s->stream;
  PVDB_TRY(s->use_stream(st));
  const size_t q_bytes = static_cast<size_t>(nq) * s->dim * sizeof(float);
  const size_t n_out = static_cast<size_t>(nq) * k;
  const size_t out_bytes = n_out * (sizeof(int64_t) + sizeof(float));
  PVDB_TRY(s->d_in.ensure(q_bytes));
  PVDB_TRY(s->d_out.ensure(out_bytes));
  PVDB_TRY(s->h_pinned.ensure(out_bytes + 16));
  PVDB_CUDA(cudaMemcpyAsync(s->d_in.ptr, queries, q_bytes, cudaMemcpyHostToDevice, st));
  const bool zero_copy = out_bytes <= (64u << 10);
  int64_t* d_rows = static_cast<int64_t*>(zero_copy ? s->h_pinned.ptr : s->d_out.ptr);
  float* d_scores = reinterpret_cast<float*>(d_rows + n_out);
  unsigned long long* h_count = reinterpret_cast<unsigned long long*>(static_cast<unsigned char*>(s->h_pinned.ptr) +
                                                                     ((out_bytes + 7) & ~size_t(7)));
  *h_count = 0;
  if (s->rows == 0) {
    PVDB_TRY(search_device(s, static_cast<const float*>(s->d_in.ptr), nq, k, nullptr, flags, d_scores, d_rows, st));
  } else {
    const uint32_t* d_bits = nullptr;
    unsigned long long* d_count = nullptr;
    PVDB_TRY(build_column_filter(s, column, wanted, n_wanted, extra_bits, &d_bits, &d_count, st));
    PVDB_TRY(search_device(s, static_cast<const float*>(s->d_in.ptr), nq, k, d_bits, flags, d_scores, d_rows, st));
    PVDB_CUDA(cudaMemcpyAsync(h_count, d_count, sizeof(unsigned long long), cudaMemcpyDeviceToHost, st));
  }
  if (!zero_copy) PVDB_CUDA(cudaMemcpyAsync(s->h_pinned.ptr, s->d_out.ptr, out_bytes, cudaMemcpyDeviceToHost, st));
  PVDB_CUDA(cudaStreamSynchronize(st));
  const int64_t* h_rows = static_cast<const int64_t*>(s->h_pinned.ptr);
  std::memcpy(out_rows, h_rows, n_out * sizeof(int64_t));
  std::memcpy(out_scores, h_rows + n_out, n_out * sizeof(float));
  if (out_candidates) *out_candidates = static_cast<int64_t>(*h_count);
  return PVDB_OK;
}

#ifdef PVDB_SCAN_TRACE
// variant builds only (tools/scan_trace.py): the phase stamps of the store's last scan launch
extern "C" int pvdb_debug_scan_trace(pvdb_store_t* s, unsigned long long* out24) {
  PVDB_ENTER(s);
  const size_t list_bytes = static_cast<size_t>(scan_grid_blocks()) * kFusedK * sizeof(uint64_t);
  PVDB_CUDA(cudaDeviceSynchronize());
  PVDB_CUDA(cudaMemcpy(out24, static_cast<unsigned char*>(s->d_partial.ptr) + list_bytes + 64, 192, cudaMemcpyDeviceToHost));
  return PVDB_OK;
}
#endif

extern "C" int pvdb_store_guard_stats(pvdb_store_t* s, int64_t* out_last, int64_t* out_total) {
  PVDB_ENTER(s);
  if (out_last) *out_last = s->guard_flagged_last;
  if (out_total) *out_total = s->guard_flagged_total;
  return PVDB_OK;
}

extern "C" int pvdb_merge_topk_dev(int device, const float* d_scores, const int64_t* d_rows, int nlists,
                                   int64_t nq, int k, int64_t scores_stride, int64_t rows_stride,
                                   float* d_out_scores, int64_t* d_out_rows, void* stream) {
  PVDB_CUDA(cudaSetDevice(device));
  if (!d_scores || !d_rows || !d_out_scores || !d_out_rows) return fail(PVDB_ERR_INVALID, "merge: null buffer");
  return launch_merge_topk(d_scores, d_rows, nlists, nq, k, scores_stride, rows_stride, d_out_scores, d_out_rows,
                           static_cast<cudaStream_t>(stream));
}
