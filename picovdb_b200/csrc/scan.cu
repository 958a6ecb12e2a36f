// Single-query masked scan with fused top-k, query preparation, and the k-way merge kernel.
//
// Replaces, for one query, the reference's
//     scores = vecs @ V.T  (or V[candidates])      picovdb/pico_vdb.py:683-689
//     argpartition / argsort / take_along_axis      picovdb/pico_vdb.py:698-714
// with ONE pass over the matrix: every row is read from HBM exactly once with 128-bit streaming
// loads, rows whose (active & prefilter) bit is clear are not read at all, the dot product is
// reduced with warp shuffles, and each warp keeps its k best candidates in registers.  Per-block
// lists are merged through shared memory, the last block to finish merges the per-block lists
// and writes the result -- the score vector never exists in memory.
//
// HBM roofline: algorithmic bytes per query = scored_rows * dim * sizeof(T) + rows/8 (bitmap).
#include <algorithm>
#include <cstdlib>

#include "scan.cuh"

namespace pvdb {

int scan_grid_blocks() { return kNumSMs * kScanBlocksPerSM; }

// defined in scan_inst_*.cu
template <bool BF16, bool SPARSE>
int launch_scan_variant(const ScanParams& p, int lpr, int ch, cudaStream_t stream);
extern template int launch_scan_variant<false, false>(const ScanParams&, int, int, cudaStream_t);
extern template int launch_scan_variant<false, true>(const ScanParams&, int, int, cudaStream_t);
extern template int launch_scan_variant<true, false>(const ScanParams&, int, int, cudaStream_t);
extern template int launch_scan_variant<true, true>(const ScanParams&, int, int, cudaStream_t);
// defined in scan_inst_bf16_mma_*.cu
template <bool SPARSE>
int launch_scan_mma_variant(const ScanParams& p, cudaStream_t stream);
extern template int launch_scan_mma_variant<false>(const ScanParams&, cudaStream_t);
extern template int launch_scan_mma_variant<true>(const ScanParams&, cudaStream_t);

// defined in scan_inst_f32_multi_*.cu / scan_inst_bf16_mma_multi_*.cu
template <bool SPARSE>
int launch_scan_multi_variant(const ScanParams& p, int lpr, int ch, cudaStream_t stream);
extern template int launch_scan_multi_variant<false>(const ScanParams&, int, int, cudaStream_t);
extern template int launch_scan_multi_variant<true>(const ScanParams&, int, int, cudaStream_t);
template <bool SPARSE>
int launch_scan_mma_multi_variant(const ScanParams& p, cudaStream_t stream);
extern template int launch_scan_mma_multi_variant<false>(const ScanParams&, cudaStream_t);
extern template int launch_scan_mma_multi_variant<true>(const ScanParams&, cudaStream_t);

// lanes per row: the widest of {32,16,8} that divides the row's 16-byte chunk count (else the
// widest that does not exceed it), then the smallest unroll in {1,2,4,8} covering the row.
static void scan_row_shape(int rc, int* lpr_out, int* ch_out) {
  int lpr = 8;
  if (rc % 32 == 0) lpr = 32;
  else if (rc % 16 == 0) lpr = 16;
  else if (rc % 8 == 0) lpr = 8;
  else if (rc >= 32) lpr = 32;
  else if (rc >= 16) lpr = 16;
  const int per_lane = (rc + lpr - 1) / lpr;
  *lpr_out = lpr;
  *ch_out = per_lane <= 1 ? 1 : per_lane <= 2 ? 2 : per_lane <= 4 ? 4 : 8;
}

// Several queries per pass (scan_kernel.cuh): 4 over fp32 rows, 2 over bf16 rows on mma.sync; k <= 32; the
// queries of one launch must fit shared memory next to the lists.  PVDB_SCAN_NO_MULTI=1 switches it off.
int scan_multi_width(bool is_bf16, int query_floats, int k) {
  const bool off = getenv("PVDB_SCAN_NO_MULTI") != nullptr;
  const bool no_mma = getenv("PVDB_SCAN_NO_MMA") != nullptr;
  if (off || k < 1 || k > 32) return 1;
  if (is_bf16 && no_mma) return 1;   // the CUDA-core bf16 kernel has no several-queries form
  const int nq = is_bf16 ? 2 : 4;
  // fp32: nq queries + 16 warps x nq lists; bf16 (wide rows): + 3 nq fragment parts of query_floats / 2 words
  const size_t smem = static_cast<size_t>(nq) * query_floats * sizeof(float) + static_cast<size_t>(nq) * 16 * k * 8 +
                      (is_bf16 ? static_cast<size_t>(3 * nq) * (query_floats / 2 + 16) * sizeof(uint32_t) : 0);
  return smem <= 160 * 1024 ? nq : 1;
}

int launch_scan_multi(const ScanParams& p, bool is_bf16, cudaStream_t stream) {
  const int width = scan_multi_width(is_bf16, p.query_floats, p.k);
  if (p.nq < 1 || p.nq > width || p.upper != nullptr || p.xv.world > 0)
    return fail(PVDB_ERR_INVALID, "scan: %d queries per pass not available here (width %d, k=%d)", p.nq, width, p.k);
  const bool sparse = p.prefilter != nullptr && getenv("PVDB_SCAN_NO_SPARSE") == nullptr;
  if (is_bf16) return sparse ? launch_scan_mma_multi_variant<true>(p, stream) : launch_scan_mma_multi_variant<false>(p, stream);
  int lpr, ch;
  scan_row_shape(p.row_chunks, &lpr, &ch);
  // Every 16-byte chunk of a row meets NQ query chunks from shared memory; with R = 16 / CH rows in flight per
  // lane group one query read serves R rows.  CH = 8 with 8 loads in flight (R = 1, what a lone query takes on
  // wide rows) spent 4 LDS.128 per loaded chunk -- ~80 % of the shared-memory pipe at HBM speed; measured
  // 3.4 TB/s on 100k x 1024 -- so the several-queries kernels cap CH at 4 (R >= 4; CH = 2 spills).
  int ch_cap = 4;
  if (const char* e = getenv("PVDB_SCAN_MULTI_CH")) ch_cap = atoi(e);
  ch = std::min(ch, std::max(1, ch_cap));
  return sparse ? launch_scan_multi_variant<true>(p, lpr, ch, stream) : launch_scan_multi_variant<false>(p, lpr, ch, stream);
}

int launch_scan(const ScanParams& p, bool is_bf16, cudaStream_t stream) {
  if (p.k < 1 || p.k > kFusedK) return fail(PVDB_ERR_INVALID, "scan: k=%d outside [1, %d]", p.k, kFusedK);
  if (p.query_floats > 16384) return fail(PVDB_ERR_UNSUPPORTED, "scan: dim > 16384 not supported");
  int lpr, ch;
  scan_row_shape(p.row_chunks, &lpr, &ch);
  // a prefilter usually leaves a small fraction of the rows: walk the bitmap, not the rows
  const bool sparse = p.prefilter != nullptr && getenv("PVDB_SCAN_NO_SPARSE") == nullptr;
  // bf16 rows: dot products on mma.sync (fewer instructions per byte: the chip sustains more of its HBM
  // bandwidth under the power cap -- see scan_mma_topk_kernel)
  if (is_bf16 && getenv("PVDB_SCAN_NO_MMA") == nullptr)
    return sparse ? launch_scan_mma_variant<true>(p, stream) : launch_scan_mma_variant<false>(p, stream);
  if (is_bf16) return sparse ? launch_scan_variant<true, true>(p, lpr, ch, stream)
                             : launch_scan_variant<true, false>(p, lpr, ch, stream);
  return sparse ? launch_scan_variant<false, true>(p, lpr, ch, stream)
                : launch_scan_variant<false, false>(p, lpr, ch, stream);
}

// ---------------------------------------------------------------------------- query preparation
// One warp per query: fp32 sum of squares, fp32 norm, IEEE division, zero query -> e0
// (picovdb/pico_vdb.py:584-591).  Output rows are zero padded to ldq floats.
__global__ void __launch_bounds__(256) prepare_queries_kernel(const float* __restrict__ raw, int64_t nq, int dim,
                                                              int already_normalised, float* __restrict__ qn,
                                                              __nv_bfloat16* __restrict__ qn16, int ldq,
                                                              float* __restrict__ qeps) {
  const int lane = threadIdx.x & 31;
  const int64_t warp = (static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = (static_cast<int64_t>(gridDim.x) * blockDim.x) >> 5;
  for (int64_t i = warp; i < nq; i += nwarps) {
    const float* v = raw + i * dim;
    float nrm = 1.f;
    bool zero = false;
    if (!already_normalised) {
      float ss = 0.f;
      for (int c = lane; c < dim; c += 32) {
        const float x = v[c];
        ss = fmaf(x, x, ss);
      }
      double d = static_cast<double>(ss);
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) d += __shfl_xor_sync(0xffffffffu, d, o);
      nrm = static_cast<float>(sqrt(d));
      zero = (nrm == 0.f);
    }
    float e_tf = 0.f, e_bf = 0.f;
    for (int c = lane; c < ldq; c += 32) {
      float y = 0.f;
      if (c < dim) y = zero ? (c == 0 ? 1.f : 0.f) : (already_normalised ? v[c] : __fdiv_rn(v[c], nrm));
      qn[i * ldq + c] = y;
      if (qn16) qn16[i * ldq + c] = __float2bfloat16_rn(y);
      const float dt = tf32_trunc_err(y), db = bf16_rn_err(y);
      e_tf = fmaf(dt, dt, e_tf);
      e_bf = fmaf(db, db, e_bf);
    }
    if (qeps != nullptr) {
      // norm of what the tensor-core operand loses of this query (exactness guard, batch.cu), and the
      // query's own norm (1 unless the caller passed "normalised" queries that are not)
      float nn = 0.f;
      for (int c = lane; c < dim; c += 32) {
        const float y = qn[i * ldq + c];
        nn = fmaf(y, y, nn);
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        e_tf += __shfl_xor_sync(0xffffffffu, e_tf, o);
        e_bf += __shfl_xor_sync(0xffffffffu, e_bf, o);
        nn += __shfl_xor_sync(0xffffffffu, nn, o);
      }
      if (lane == 0) {
        qeps[i * 4 + 0] = sqrtf(e_tf);
        qeps[i * 4 + 1] = sqrtf(e_bf);
        qeps[i * 4 + 2] = sqrtf(nn);
        qeps[i * 4 + 3] = 0.f;
      }
    }
  }
}

int launch_prepare_queries(const float* d_raw, int64_t nq, int dim, bool already_normalised, float* d_qn,
                           __nv_bfloat16* d_qn16, int ldq, float* d_qeps, cudaStream_t stream) {
  const int blocks = static_cast<int>(std::max<int64_t>(1, std::min<int64_t>((nq + 7) / 8, kNumSMs * 8)));
  prepare_queries_kernel<<<blocks, 256, 0, stream>>>(d_raw, nq, dim, already_normalised ? 1 : 0, d_qn, d_qn16, ldq,
                                                     d_qeps);
  PVDB_LAUNCH_CHECK();
  return PVDB_OK;
}

// ---------------------------------------------------------------------------- k-way merge
// One block per query: the nlists*k candidates (already (score desc, row asc) inside each list)
// are turned into keys, sorted with a shared-memory bitonic network, and the best k written out.
// This is the merge that follows the NCCL all-gather of the per-GPU results.
__global__ void __launch_bounds__(256) merge_topk_kernel(const float* __restrict__ scores,
                                                         const int64_t* __restrict__ rows, int nlists, int64_t nq,
                                                         int k, int64_t scores_stride, int64_t rows_stride,
                                                         int n_pow2, float* __restrict__ out_scores,
                                                         int64_t* __restrict__ out_rows) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  uint64_t* keys = reinterpret_cast<uint64_t*>(smem_raw);
  const int64_t q = blockIdx.x;
  const int total = nlists * k;
  for (int i = threadIdx.x; i < n_pow2; i += blockDim.x) {
    uint64_t key = 0ull;
    if (i < total) {
      const int l = i / k, j = i - l * k;
      const size_t off = static_cast<size_t>(q) * k + j;
      const int64_t r = rows[static_cast<size_t>(l) * rows_stride + off];
      const float sc = scores[static_cast<size_t>(l) * scores_stride + off];
      if (r >= 0 && sc == sc) key = make_key(sc, static_cast<uint32_t>(r));
    }
    keys[i] = key;
  }
  for (int size = 2; size <= n_pow2; size <<= 1) {
    for (int stride = size >> 1; stride > 0; stride >>= 1) {
      __syncthreads();
      for (int i = threadIdx.x; i < (n_pow2 >> 1); i += blockDim.x) {
        const int lo = 2 * i - (i & (stride - 1));
        const int hi = lo + stride;
        const bool desc = (lo & size) == 0;
        const uint64_t a = keys[lo], b = keys[hi];
        if ((a < b) == desc) {
          keys[lo] = b;
          keys[hi] = a;
        }
      }
    }
  }
  __syncthreads();
  for (int j = threadIdx.x; j < k; j += blockDim.x) {
    const uint64_t key = keys[j];
    out_scores[q * k + j] = key ? key_score(key) : -INFINITY;
    out_rows[q * k + j] = key ? static_cast<int64_t>(key_row(key)) : -1ll;
  }
}

int launch_merge_topk(const float* d_scores, const int64_t* d_rows, int nlists, int64_t nq, int k,
                      int64_t scores_stride, int64_t rows_stride, float* d_out_scores, int64_t* d_out_rows,
                      cudaStream_t stream) {
  if (scores_stride <= 0) scores_stride = nq * k;
  if (rows_stride <= 0) rows_stride = nq * k;
  if (nlists < 1 || k < 1 || nq < 0) return fail(PVDB_ERR_INVALID, "merge: bad arguments");
  if (nq == 0) return PVDB_OK;
  const int64_t total = static_cast<int64_t>(nlists) * k;
  if (total > 16384) return fail(PVDB_ERR_UNSUPPORTED, "merge: nlists*k = %lld exceeds 16384", (long long)total);
  int n_pow2 = 2;
  while (n_pow2 < total) n_pow2 <<= 1;
  const size_t smem = static_cast<size_t>(n_pow2) * sizeof(uint64_t);
  if (smem > 48 * 1024)
    PVDB_CUDA(cudaFuncSetAttribute(merge_topk_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
  merge_topk_kernel<<<static_cast<unsigned>(nq), 256, smem, stream>>>(d_scores, d_rows, nlists, nq, k, scores_stride,
                                                                      rows_stride, n_pow2, d_out_scores, d_out_rows);
  PVDB_LAUNCH_CHECK();
  return PVDB_OK;
}

}  // namespace pvdb
