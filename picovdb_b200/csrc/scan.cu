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

constexpr int kScanThreads = 512;
constexpr int kScanWarps = kScanThreads / 32;
constexpr int kScanBlocksPerSM = 2;
constexpr int kSlots = kFusedK / 32;

int scan_grid_blocks() { return kNumSMs * kScanBlocksPerSM; }

template <bool GLOBAL>
__device__ __forceinline__ uint64_t load_key(const uint64_t* p) {
  if constexpr (GLOBAL) {
    return __ldcg(reinterpret_cast<const unsigned long long*>(p));  // L2: written by other SMs
  } else {
    return *p;
  }
}

template <bool GLOBAL>
__device__ __forceinline__ void merge_list(WarpList<kSlots>& L, uint64_t& thr, const uint64_t* src, int n, int k,
                                           int lane) {
  // src: descending list of n keys; k: rank whose key is the admission threshold
  for (int base = 0; base < n; base += 32) {
    const int e = base + lane;
    const uint64_t v = (e < n) ? load_key<GLOBAL>(src + e) : 0ull;
    unsigned m = __ballot_sync(0xffffffffu, v > thr);
    if (m == 0) break;  // lists are descending: nothing further can qualify
    while (m) {
      const int srcl = __ffs(m) - 1;
      m &= m - 1;
      const uint64_t x = shfl_u64(v, srcl);
      if (x > thr) {
        L.insert(x, lane);
        thr = L.get(k - 1);
      }
    }
  }
}

__device__ __forceinline__ void store_list(const WarpList<kSlots>& L, uint64_t* dst, int k, int lane) {
#pragma unroll
  for (int s = 0; s < kSlots; ++s) {
    const int e = s * 32 + lane;
    if (e < k) dst[e] = L.slot[s];
  }
}

__device__ __forceinline__ float dot_chunk_f32(const uint4& v, const float4& q, float acc) {
  acc = fmaf(__uint_as_float(v.x), q.x, acc);
  acc = fmaf(__uint_as_float(v.y), q.y, acc);
  acc = fmaf(__uint_as_float(v.z), q.z, acc);
  acc = fmaf(__uint_as_float(v.w), q.w, acc);
  return acc;
}
// 8 bf16 values (one 16-byte chunk) against 8 fp32 query values; bf16 -> fp32 is a 16-bit shift
__device__ __forceinline__ float dot_chunk_bf16(const uint4& v, const float4& q0, const float4& q1, float acc) {
  acc = fmaf(__uint_as_float(v.x << 16), q0.x, acc);
  acc = fmaf(__uint_as_float(v.x & 0xffff0000u), q0.y, acc);
  acc = fmaf(__uint_as_float(v.y << 16), q0.z, acc);
  acc = fmaf(__uint_as_float(v.y & 0xffff0000u), q0.w, acc);
  acc = fmaf(__uint_as_float(v.z << 16), q1.x, acc);
  acc = fmaf(__uint_as_float(v.z & 0xffff0000u), q1.y, acc);
  acc = fmaf(__uint_as_float(v.w << 16), q1.z, acc);
  acc = fmaf(__uint_as_float(v.w & 0xffff0000u), q1.w, acc);
  return acc;
}

// LPR lanes cooperate on one row; each lane keeps CH 16-byte loads of R rows in flight
// (CH * R == 8 -> eight independent 128-bit loads per lane per step).
template <bool BF16, int LPR, int CH, bool SPARSE>
__global__ void __launch_bounds__(kScanThreads, kScanBlocksPerSM) scan_topk_kernel(const ScanParams p) {
  constexpr int G = 32 / LPR;  // row groups per warp
  constexpr int R = 8 / CH;    // rows in flight per group
  constexpr int RPW = G * R;   // rows per warp step: a power of two <= 32, so one bitmap word covers it
  extern __shared__ __align__(16) unsigned char smem_raw[];
  float* sq = reinterpret_cast<float*>(smem_raw);
  uint64_t* slist = reinterpret_cast<uint64_t*>(smem_raw + static_cast<size_t>(p.query_floats) * sizeof(float));
  __shared__ unsigned s_is_last;

  const int tid = threadIdx.x;
  const int lane = tid & 31;
  const int warp = tid >> 5;
  const int sub = lane % LPR;
  const int gi = lane / LPR;
  const int k = p.k;

  if (p.raw_query != nullptr) {
    // Fused query preparation (picovdb/pico_vdb.py:584-591): every block normalises the raw query
    // itself -- fp32 sum of squares, fp32 norm, IEEE division, zero query -> e0 -- which saves a
    // kernel launch and a round trip through HBM on the single-query path.
    __shared__ float s_part[kScanWarps];
    float ss = 0.f;
    for (int i = tid; i < p.query_floats; i += kScanThreads) {
      const float x = (i < p.dim) ? p.raw_query[i] : 0.f;
      sq[i] = x;
      ss = fmaf(x, x, ss);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
    if (lane == 0) s_part[warp] = ss;
    __syncthreads();
    double tot = 0.0;
#pragma unroll
    for (int w2 = 0; w2 < kScanWarps; ++w2) tot += static_cast<double>(s_part[w2]);
    const float nrm = static_cast<float>(sqrt(tot));
    for (int i = tid; i < p.query_floats; i += kScanThreads) {
      const float x = sq[i];
      sq[i] = (nrm == 0.f) ? (i == 0 ? 1.f : 0.f) : __fdiv_rn(x, nrm);
    }
  } else {
    for (int i = tid; i < p.query_floats; i += kScanThreads) sq[i] = p.query[i];
  }
  __syncthreads();
  const float4* sq4 = reinterpret_cast<const float4*>(sq);

  const uint64_t upper = p.upper ? *p.upper : ~0ull;
  WarpList<kSlots> L;
  L.clear();
  uint64_t thr = 0ull;

  const int64_t total_warps = static_cast<int64_t>(gridDim.x) * kScanWarps;
  const uint4* mat = reinterpret_cast<const uint4*>(p.matrix);
  const int row_chunks = p.row_chunks;

  // Score the (up to) R rows this lane group holds -- row[r] valid iff on[r] -- and feed the warp list.
  auto score_rows = [&](const int64_t (&row)[R], const bool (&on)[R]) {
    float acc[R];
    const uint4* rp[R];
#pragma unroll
    for (int r = 0; r < R; ++r) {
      rp[r] = mat + row[r] * row_chunks;
      acc[r] = 0.f;
    }
    for (int c0 = 0; c0 < row_chunks; c0 += LPR * CH) {
      uint4 v[R][CH];
#pragma unroll
      for (int r = 0; r < R; ++r) {
#pragma unroll
        for (int i = 0; i < CH; ++i) {
          const int idx = c0 + i * LPR + sub;
          v[r][i] = (on[r] && idx < row_chunks) ? ldg_stream(rp[r] + idx) : make_uint4(0u, 0u, 0u, 0u);
        }
      }
#pragma unroll
      for (int i = 0; i < CH; ++i) {
        const int idx = c0 + i * LPR + sub;
        if (idx < row_chunks) {
          if constexpr (BF16) {
            const float4 q0 = sq4[2 * idx];
            const float4 q1 = sq4[2 * idx + 1];
#pragma unroll
            for (int r = 0; r < R; ++r) acc[r] = dot_chunk_bf16(v[r][i], q0, q1, acc[r]);
          } else {
            const float4 q = sq4[idx];
#pragma unroll
            for (int r = 0; r < R; ++r) acc[r] = dot_chunk_f32(v[r][i], q, acc[r]);
          }
        }
      }
    }
#pragma unroll
    for (int r = 0; r < R; ++r) {
#pragma unroll
      for (int o = LPR / 2; o > 0; o >>= 1) acc[r] += __shfl_xor_sync(0xffffffffu, acc[r], o);
    }
#pragma unroll
    for (int r = 0; r < R; ++r) {
      const float sc = acc[r];
      const uint64_t key = (on[r] && sc == sc) ? make_key(sc, static_cast<uint32_t>(row[r])) : 0ull;
      unsigned m = __ballot_sync(0xffffffffu, sub == 0 && key > thr && key < upper);
      while (m) {
        const int srcl = __ffs(m) - 1;
        m &= m - 1;
        const uint64_t x = shfl_u64(key, srcl);
        if (x > thr) {
          L.insert(x, lane);
          thr = L.get(k - 1);
        }
      }
    }
  };

  if constexpr (!SPARSE) {
    // Dense walk: a warp step covers RPW consecutive rows (one bitmap word covers a step); steps are
    // interleaved over all warps of the grid.  (Fetching the bitmap word one step ahead was measured
    // on the B200 and lost: +15 % on the bf16 scan from the extra live registers.)
    const int64_t n_steps = (p.n_rows + RPW - 1) / RPW;
    for (int64_t step = static_cast<int64_t>(blockIdx.x) * kScanWarps + warp; step < n_steps; step += total_warps) {
      const int64_t base = step * RPW;
      uint32_t w = __ldg(p.active + (base >> 5));
      if (p.prefilter) w &= __ldg(p.prefilter + (base >> 5));
      w >>= (base & 31);
      if constexpr (RPW < 32) w &= (1u << RPW) - 1u;
      if (w == 0u) continue;  // every row of this step is deleted / filtered out: read nothing
      int64_t row[R];
      bool on[R];
#pragma unroll
      for (int r = 0; r < R; ++r) {
        const int local = r * G + gi;
        on[r] = (w >> local) & 1u;
        row[r] = base + local;
      }
      score_rows(row, on);
    }
  } else {
    // Sparse walk (selective prefilters): a warp takes one bitmap word = 32 consecutive rows at a
    // time and packs only the SET bits into its RPW row slots, so filtered-out rows cost neither
    // loop steps nor bitmap round trips (the dense walk pays one step per RPW rows regardless).
    const int64_t n_words = (p.n_rows + 31) >> 5;
    for (int64_t wi = static_cast<int64_t>(blockIdx.x) * kScanWarps + warp; wi < n_words; wi += total_warps) {
      uint32_t w = __ldg(p.active + wi);
      if (w != 0u && p.prefilter) w &= __ldg(p.prefilter + wi);
      while (w != 0u) {
        int64_t row[R];
        bool on[R];
#pragma unroll
        for (int r = 0; r < R; ++r) {
          const unsigned bit = __fns(w, 0, r * G + gi + 1);  // position of this slot's set bit
          on[r] = bit < 32u;
          row[r] = (wi << 5) + (on[r] ? bit : 0u);
        }
        score_rows(row, on);
        // drop the RPW lowest set bits that were just consumed
        if constexpr (RPW >= 32) {
          w = 0u;
        } else {
          const unsigned last = __fns(w, 0, RPW);
          w = (last < 31u) ? (w & (0xffffffffu << (last + 1))) : 0u;
        }
      }
    }
  }

  // ---- block merge: warp 0 folds the other warps' lists into its own
  store_list(L, slist + warp * k, k, lane);
  __syncthreads();
  if (warp == 0) {
    for (int w2 = 1; w2 < kScanWarps; ++w2) merge_list<false>(L, thr, slist + w2 * k, k, k, lane);
    store_list(L, p.partial + static_cast<size_t>(blockIdx.x) * k, k, lane);
    __threadfence();
    __syncwarp();
    if (lane == 0) {
      const unsigned t = atomicAdd(p.ticket, 1u);
      s_is_last = (t == gridDim.x - 1) ? 1u : 0u;
    }
  }
  __syncthreads();
  if (s_is_last == 0u) return;

  // ---- last block: merge all per-block lists and emit the result
  __threadfence();
  L.clear();
  thr = 0ull;
  // Each warp takes every 16th block list.  The heads (first 32 keys) of eight lists are fetched
  // together so the L2 round trips overlap; a list whose whole head qualified continues through
  // the general path.
  for (int b0 = warp; b0 < static_cast<int>(gridDim.x); b0 += kScanWarps * 8) {
    uint64_t head[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int b = b0 + j * kScanWarps;
      head[j] = (b < static_cast<int>(gridDim.x) && lane < k)
                    ? load_key<true>(p.partial + static_cast<size_t>(b) * k + lane)
                    : 0ull;
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int b = b0 + j * kScanWarps;
      if (b >= static_cast<int>(gridDim.x)) break;
      unsigned m = __ballot_sync(0xffffffffu, head[j] > thr);
      const bool head_all = (m == 0xffffffffu);
      while (m) {
        const int srcl = __ffs(m) - 1;
        m &= m - 1;
        const uint64_t x = shfl_u64(head[j], srcl);
        if (x > thr) {
          L.insert(x, lane);
          thr = L.get(k - 1);
        }
      }
      if (head_all && k > 32) merge_list<true>(L, thr, p.partial + static_cast<size_t>(b) * k + 32, k - 32, k, lane);
    }
  }
  __syncthreads();  // everyone is done reading slist from the first merge
  store_list(L, slist + warp * k, k, lane);
  __syncthreads();
  if (warp == 0) {
    for (int w2 = 1; w2 < kScanWarps; ++w2) merge_list<false>(L, thr, slist + w2 * k, k, k, lane);
#pragma unroll
    for (int s = 0; s < kSlots; ++s) {
      const int e = s * 32 + lane;
      if (e < k) {
        const uint64_t key = L.slot[s];
        p.out_scores[e] = key ? key_score(key) : -INFINITY;
        p.out_rows[e] = key ? p.row_base + static_cast<int64_t>(key_row(key)) : -1ll;
      }
    }
    const uint64_t kth = L.get(k - 1);
    if (lane == 0) {
      *p.next_upper = kth;
      *p.ticket = 0u;
    }
  }
}

template <bool BF16, int LPR, int CH>
static int launch_scan_t(const ScanParams& p, cudaStream_t stream) {
  const size_t smem = static_cast<size_t>(p.query_floats) * sizeof(float) +
                      static_cast<size_t>(kScanWarps) * p.k * sizeof(uint64_t);
  // a prefilter usually leaves a small fraction of the rows: walk the bitmap, not the rows
  const bool sparse = p.prefilter != nullptr && getenv("PVDB_SCAN_NO_SPARSE") == nullptr;
  auto kern = sparse ? scan_topk_kernel<BF16, LPR, CH, true> : scan_topk_kernel<BF16, LPR, CH, false>;
  if (smem > 48 * 1024) {
    PVDB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
  }
  kern<<<scan_grid_blocks(), kScanThreads, smem, stream>>>(p);
  PVDB_LAUNCH_CHECK();
  return PVDB_OK;
}

template <bool BF16, int LPR>
static int launch_scan_ch(const ScanParams& p, int ch, cudaStream_t stream) {
  switch (ch) {
    case 1: return launch_scan_t<BF16, LPR, 1>(p, stream);
    case 2: return launch_scan_t<BF16, LPR, 2>(p, stream);
    case 4: return launch_scan_t<BF16, LPR, 4>(p, stream);
    default: return launch_scan_t<BF16, LPR, 8>(p, stream);
  }
}

int launch_scan(const ScanParams& p, bool is_bf16, cudaStream_t stream) {
  if (p.k < 1 || p.k > kFusedK) return fail(PVDB_ERR_INVALID, "scan: k=%d outside [1, %d]", p.k, kFusedK);
  if (p.query_floats > 16384) return fail(PVDB_ERR_UNSUPPORTED, "scan: dim > 16384 not supported");
  // lanes per row: the widest of {32,16,8} that divides the row's 16-byte chunk count (else the
  // widest that does not exceed it), then the smallest unroll in {1,2,4,8} covering the row.
  const int rc = p.row_chunks;
  int lpr = 8;
  if (rc % 32 == 0) lpr = 32;
  else if (rc % 16 == 0) lpr = 16;
  else if (rc % 8 == 0) lpr = 8;
  else if (rc >= 32) lpr = 32;
  else if (rc >= 16) lpr = 16;
  const int per_lane = (rc + lpr - 1) / lpr;
  const int ch = per_lane <= 1 ? 1 : per_lane <= 2 ? 2 : per_lane <= 4 ? 4 : 8;
  if (is_bf16) {
    if (lpr == 32) return launch_scan_ch<true, 32>(p, ch, stream);
    if (lpr == 16) return launch_scan_ch<true, 16>(p, ch, stream);
    return launch_scan_ch<true, 8>(p, ch, stream);
  }
  if (lpr == 32) return launch_scan_ch<false, 32>(p, ch, stream);
  if (lpr == 16) return launch_scan_ch<false, 16>(p, ch, stream);
  return launch_scan_ch<false, 8>(p, ch, stream);
}

// ---------------------------------------------------------------------------- query preparation
// One warp per query: fp32 sum of squares, fp32 norm, IEEE division, zero query -> e0
// (picovdb/pico_vdb.py:584-591).  Output rows are zero padded to ldq floats.
__global__ void __launch_bounds__(256) prepare_queries_kernel(const float* __restrict__ raw, int64_t nq, int dim,
                                                              int already_normalised, float* __restrict__ qn,
                                                              __nv_bfloat16* __restrict__ qn16, int ldq) {
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
    for (int c = lane; c < ldq; c += 32) {
      float y = 0.f;
      if (c < dim) y = zero ? (c == 0 ? 1.f : 0.f) : (already_normalised ? v[c] : __fdiv_rn(v[c], nrm));
      qn[i * ldq + c] = y;
      if (qn16) qn16[i * ldq + c] = __float2bfloat16_rn(y);
    }
  }
}

int launch_prepare_queries(const float* d_raw, int64_t nq, int dim, bool already_normalised, float* d_qn,
                           __nv_bfloat16* d_qn16, int ldq, cudaStream_t stream) {
  const int blocks = static_cast<int>(std::max<int64_t>(1, std::min<int64_t>((nq + 7) / 8, kNumSMs * 8)));
  prepare_queries_kernel<<<blocks, 256, 0, stream>>>(d_raw, nq, dim, already_normalised ? 1 : 0, d_qn, d_qn16, ldq);
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
