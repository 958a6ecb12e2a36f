// The scan kernel templates and their per-variant launchers.  Included by scan_inst_*.cu (one translation
// unit per (kernel family, dtype, sparse) so the ~150 instantiations compile in parallel).
#pragma once

#include <cstdlib>

#include "scan.cuh"

namespace pvdb {



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

// Sparse walk (selective prefilters), shared by every scan variant.  A warp takes one bitmap word = 32
// consecutive rows at a time (words interleaved over all warps of the grid) and appends the rows whose
// (active & prefilter) bit is set to a small ring of pending rows in shared memory; whenever RPW rows are
// pending, `step(ring, head, RPW)` scores a FULL warp step.  Filtered-out rows cost neither loop steps nor
// bitmap round trips, a step never runs half empty because a word ran out of set bits (the first form packed
// the rows of ONE word per step: 2.2 live rows per word at a 10 % filter over 30 % deleted rows left a third of
// the row slots idle -- 0.68 x the HBM rate on live bytes), and the next word is in flight while the pending
// rows are scored.  ring: 64 row ids of this warp; slot j of a step is ring[(head + j) & 63].
template <int RPW, typename Step>
__device__ __forceinline__ void sparse_walk(const ScanParams& p, int64_t first_word, int64_t total_warps, int lane,
                                            uint32_t* ring, Step&& step) {
  static_assert(RPW >= 1 && RPW <= 32, "pending (< RPW) + one word (32) must fit the ring of 64");
  const int64_t n_words = (p.n_rows + 31) >> 5;
  unsigned head = 0u, cnt = 0u;
  int64_t wi = first_word;
  uint32_t w = 0u;
  if (wi < n_words) {
    w = __ldg(p.active + wi);
    if (w != 0u && p.prefilter) w &= __ldg(p.prefilter + wi);
  }
  while (wi < n_words) {
    const int64_t wn = wi + total_warps;
    uint32_t w_next = 0u;
    if (wn < n_words) {
      w_next = __ldg(p.active + wn);
      if (w_next != 0u && p.prefilter) w_next &= __ldg(p.prefilter + wn);
    }
    const unsigned c = __popc(w);
    __syncwarp();   // the previous steps' reads of the ring are done
    if (static_cast<unsigned>(lane) < c)
      ring[(head + cnt + lane) & 63u] = static_cast<uint32_t>(wi << 5) + __fns(w, 0, lane + 1);
    cnt += c;
    __syncwarp();
    while (cnt >= static_cast<unsigned>(RPW)) {
      step(ring, head, static_cast<unsigned>(RPW));
      head = (head + RPW) & 63u;
      cnt -= RPW;
    }
    w = w_next;
    wi = wn;
  }
  if (cnt > 0u) step(ring, head, cnt);   // the one partial step of this warp
}

// -DPVDB_SCAN_TRACE (variant builds only, tools/scan_trace.py): %globaltimer stamps of the phases of one launch,
// written behind the ticket word (block 0: slots 0-3, the last block: slots 4-6).
#ifdef PVDB_SCAN_TRACE
__device__ __forceinline__ void scan_trace(const ScanParams& p, int slot, bool who) {
  if (who && threadIdx.x == 0) {
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    reinterpret_cast<unsigned long long*>(p.ticket)[8 + slot] = t;
  }
}
#define PVDB_TRACE(slot, who) scan_trace(p, slot, who)
#else
#define PVDB_TRACE(slot, who) ((void)0)
#endif

// What every scan variant does first: programmatic-dependent-launch bookkeeping, then the normalised query
// goes to shared memory (sq[0 .. query_floats), zero padded).
__device__ __forceinline__ void scan_prologue(const ScanParams& p, float* sq) {
  const int tid = threadIdx.x;
  const int lane = tid & 31;
  const int warp = tid >> 5;
  // Programmatic dependent launch: let the next scan on the stream start filling SMs as soon as this
  // grid's blocks retire (its scan phase only reads the matrix), and wait for the PREVIOUS grid only
  // where this one touches what that one may still be using (below: paging bound, per-block lists).
  // Both instructions are no-ops in a launch without the attribute.
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  if (p.upper != nullptr) asm volatile("griddepcontrol.wait;" ::: "memory");

  if (p.raw_query != nullptr) {
    // Fused query preparation (picovdb/pico_vdb.py:584-591): every block normalises the raw query
    // itself -- fp32 sum of squares, fp32 norm, IEEE division, zero query -> e0 -- which saves a
    // kernel launch and a round trip through HBM on the single-query path.
    // The sum of squares is formed exactly like prepare_queries_kernel and the upsert kernel form it
    // (lane-strided fp32 partial sums of ONE warp, fp64 butterfly), so a query is normalised to the
    // same bits whether it arrives alone or inside a batch.
    __shared__ float s_nrm;
    for (int i = tid; i < p.query_floats; i += static_cast<int>(blockDim.x)) sq[i] = (i < p.dim) ? p.raw_query[i] : 0.f;
    __syncthreads();
    if (warp == 0) {
      float ss = 0.f;
      for (int c = lane; c < p.dim; c += 32) {
        const float x = sq[c];
        ss = fmaf(x, x, ss);
      }
      double d = static_cast<double>(ss);
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) d += __shfl_xor_sync(0xffffffffu, d, o);
      if (lane == 0) s_nrm = static_cast<float>(sqrt(d));
    }
    __syncthreads();
    const float nrm = s_nrm;
    for (int i = tid; i < p.query_floats; i += static_cast<int>(blockDim.x)) {
      const float x = sq[i];
      sq[i] = (nrm == 0.f) ? (i == 0 ? 1.f : 0.f) : __fdiv_rn(x, nrm);
    }
  } else {
    for (int i = tid; i < p.query_floats; i += static_cast<int>(blockDim.x)) sq[i] = p.query[i];
  }
  __syncthreads();
}

// What every scan variant does last: block merge of the warp lists, last-block-done merge of the per-block
// lists, optional cross-GPU exchange, result write-out.
// The merges used to be the scan's fixed cost (tools/scan_trace.py on a 1024-row store: block merge 3.6 us,
// folding the 296 block lists 5.8 us, final merge 4.9 us of 17 us per launch): every qualifying key cost an
// insertion (~150 cycles of dependent shuffles) and one warp did each block-level merge alone.  Now
//  * block-level merges are binary trees over the warps (block_tree_merge; bitonic networks for k <= 32);
//  * every block raises `floor_key` to the k-th key of its list before it takes its ticket.  The block with
//    the largest such key holds k keys at or above it, so nothing below the final floor can be in the top k:
//    the last block starts from that threshold and the 296 lists contribute a few dozen keys, not 296 k.
// HB: heads of per-block lists a warp of the last block fetches at once (one L2 round trip when
// n_warps * HB covers the grid).
template <int S, int HB>
__device__ __forceinline__ void scan_finish(const ScanParams& p, WarpList<S>& L, uint64_t& thr, uint64_t* slist) {
  __shared__ unsigned s_is_last;
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  const int k = p.k;
  const int n_warps = static_cast<int>(blockDim.x >> 5);   // (the mma variant runs smaller blocks)
  // ---- block merge
  store_list(L, slist + warp * k, k, lane);
  __syncthreads();
  // the previous scan's last block may still be merging the per-block lists (and owns the ticket)
  asm volatile("griddepcontrol.wait;" ::: "memory");
  block_tree_merge<S>(L, thr, slist, k, warp, lane, n_warps);
  if (warp == 0) {
    store_list(L, p.partial + static_cast<size_t>(blockIdx.x) * k, k, lane);
    const uint64_t kth = L.get(k - 1);
    fence_acq_rel_gpu();
    __syncwarp();
    if (lane == 0) {
      if (S != 1 && kth != 0ull) atomicMax(p.floor_key, static_cast<unsigned long long>(kth));   // (k <= 32 folds with bitonic trees: no threshold)
      const unsigned t = atomicAdd(p.ticket, 1u);
      s_is_last = (t == gridDim.x - 1) ? 1u : 0u;
    }
  }
  __syncthreads();
  PVDB_TRACE(3, blockIdx.x == 0);
  if (s_is_last == 0u) return;

  // ---- last block: merge all per-block lists and emit the result
  PVDB_TRACE(4, true);
  fence_acq_rel_gpu();
  PVDB_TRACE(7, true);
  L.clear();
  // keys >= floor qualify (the admission test below is "key > thr"); the load travels with the first heads
  const uint64_t floor_key = S != 1 ? __ldcg(p.floor_key) : 0ull;
  thr = 0ull;
  // Each warp takes every n_warps-th block list.  The heads (first 32 keys) of HB lists are fetched
  // together so the L2 round trips overlap; a list whose whole head qualified continues through
  // the general path.
  for (int b0 = warp; b0 < static_cast<int>(gridDim.x); b0 += n_warps * HB) {
    uint64_t head[HB];
#pragma unroll
    for (int j = 0; j < HB; ++j) {
      const int b = b0 + j * n_warps;
      head[j] = (b < static_cast<int>(gridDim.x) && lane < k)
                    ? load_key<true>(p.partial + static_cast<size_t>(b) * k + lane)
                    : 0ull;
    }
    if (b0 == warp) thr = floor_key ? floor_key - 1ull : 0ull;
    PVDB_TRACE(8, b0 == warp && head[0] != 1ull);
    if constexpr (S == 1) {
      // k <= 32: a list is one register per lane.  Pairwise bitonic merges of the HB heads (a tree: the merges
      // of a level are independent, so their shuffles overlap), then one merge into the running list --
      // ~0.5 us per HB lists whatever the data; insertions cost ~0.1 us per qualifying key (traced: 4-5 us for
      // this loop even with the floor, ~200 keys of 2960 qualify on Gaussian rows).
#pragma unroll
      for (int w = 1; w < HB; w <<= 1) {
#pragma unroll
        for (int j = 0; j + w < HB; j += 2 * w) head[j] = bitonic_merge_regs(head[j], head[j + w], lane);
      }
      L.slot[0] = bitonic_merge_regs(L.slot[0], head[0], lane);
      continue;
    }
#pragma unroll
    for (int j = 0; j < HB; ++j) {
      const int b = b0 + j * n_warps;
      if (b >= static_cast<int>(gridDim.x)) break;
      unsigned m = __ballot_sync(0xffffffffu, head[j] > thr);
      const bool head_all = (m == 0xffffffffu);
      while (m) {
        const int srcl = __ffs(m) - 1;
        m &= m - 1;
        const uint64_t x = shfl_u64(head[j], srcl);
        if (x > thr) {
          L.insert(x, lane);
          const uint64_t kth = L.get(k - 1);
          if (kth > thr) thr = kth;   // (never below the floor)
        }
      }
      if (head_all && k > 32) merge_list<true, S>(L, thr, p.partial + static_cast<size_t>(b) * k + 32, k - 32, k, lane);
    }
  }
  PVDB_TRACE(9, true);
  __syncthreads();  // everyone is done reading slist from the first merge
  PVDB_TRACE(5, true);
  store_list(L, slist + warp * k, k, lane);
  __syncthreads();
  block_tree_merge<S>(L, thr, slist, k, warp, lane, n_warps);
  if (warp == 0) {
    int64_t out_base = p.row_base;
    if (p.xv.world > 0) {
      // ---- cross-GPU exchange, fused (exchange.cuh): this GPU's list goes into every peer's
      // mailbox as keys with global rows, the peers' lists arrive in ours, and the k-way merge of
      // the `world` lists happens right here -- the kernel writes the FINAL top k on every GPU.
      const ExchangeView& v = p.xv;
      const int parity = static_cast<int>(v.seq & 1ull);
      for (int peer = 0; peer < v.world; ++peer) {
        uint64_t* dst = xv_slot(v, v.box[peer], parity, v.rank);
#pragma unroll
        for (int s = 0; s < S; ++s) {
          const int e = s * 32 + lane;
          if (e < k) dst[e] = key_to_global(L.slot[s], p.row_base);
        }
      }
      __threadfence_system();
      __syncwarp();
      if (lane < v.world) st_release_sys(xv_flag(v.box[lane], parity, v.rank, 0), v.seq);
      if (lane < v.world) xv_wait_flag(xv_flag(v.box[v.rank], parity, lane, 0), v.seq);
      __syncwarp();
      L.clear();
      thr = 0ull;
      for (int r = 0; r < v.world; ++r) merge_list<true, S>(L, thr, xv_slot(v, v.box[v.rank], parity, r), k, k, lane);
      out_base = 0;  // the merged keys carry global rows
    }
#pragma unroll
    for (int s = 0; s < S; ++s) {
      const int e = s * 32 + lane;
      if (e < k) {
        const uint64_t key = L.slot[s];
        p.out_scores[e] = key ? key_score(key) : -INFINITY;
        p.out_rows[e] = key ? out_base + static_cast<int64_t>(key_row(key)) : -1ll;
      }
    }
    const uint64_t kth = L.get(k - 1);
    PVDB_TRACE(6, true);
    if (lane == 0) {
      *p.next_upper = kth;
      *p.ticket = 0u;
      *p.floor_key = 0ull;
    }
  }
}

// ---------------------------------------------------------------------------- several queries per pass
// A batch that must be answered EXACTLY (precision "f32", or the queries the tensor-core guard could not prove)
// used to cost one pass over the matrix per query.  The scan is HBM bound with arithmetic to spare, so the
// kernels below score NQ queries against every row they load: fp32 rows x 4 queries on the CUDA cores
// (16 FMA per 16-byte chunk instead of 4), bf16 rows x 2 queries on mma.sync (the second query's three terms
// fill B columns the single-query form leaves zero).  Per query the arithmetic -- normalisation, products,
// order of the additions, key order -- is the single-query kernel's, so a query gets the same bits alone and
// in a group.  k <= 32 (one 32-key slot per list), no paging bound, no fused exchange: callers fall back to
// the single-query kernels outside that.
//
// Prologue: p.nq (<= NQ) consecutive queries go to shared memory, query q at sq[q * query_floats ...]; the
// slots of absent queries are zero and never produce candidates.
template <int NQ>
__device__ __forceinline__ void scan_prologue_multi(const ScanParams& p, float* sq) {
  const int tid = threadIdx.x;
  const int lane = tid & 31;
  const int warp = tid >> 5;
  const int qf = p.query_floats;
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  if (p.raw_query != nullptr) {
    // the single-query prologue per query: lane-strided fp32 partial sums of ONE warp, fp64 butterfly
    __shared__ float s_nrm[NQ];
    for (int i = tid; i < NQ * qf; i += static_cast<int>(blockDim.x)) {
      const int q = i / qf, c = i - q * qf;
      sq[i] = (q < p.nq && c < p.dim) ? p.raw_query[p.qsel[q] * p.dim + c] : 0.f;
    }
    __syncthreads();
    if (warp < NQ) {
      const float* s = sq + warp * qf;
      float ss = 0.f;
      for (int c = lane; c < p.dim; c += 32) {
        const float x = s[c];
        ss = fmaf(x, x, ss);
      }
      double d = static_cast<double>(ss);
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) d += __shfl_xor_sync(0xffffffffu, d, o);
      if (lane == 0) s_nrm[warp] = static_cast<float>(sqrt(d));
    }
    __syncthreads();
    for (int i = tid; i < NQ * qf; i += static_cast<int>(blockDim.x)) {
      const int q = i / qf, c = i - q * qf;
      const float nrm = s_nrm[q];
      const float x = sq[i];
      sq[i] = (q >= p.nq) ? 0.f : ((nrm == 0.f) ? (c == 0 ? 1.f : 0.f) : __fdiv_rn(x, nrm));
    }
  } else {
    for (int i = tid; i < NQ * qf; i += static_cast<int>(blockDim.x)) {
      const int q = i / qf, c = i - q * qf;
      sq[i] = (q < p.nq) ? p.query[p.qsel[q] * qf + c] : 0.f;
    }
  }
  __syncthreads();
}

// Finish, as scan_finish but for NQ lists per warp (k <= 32: bitonic merges throughout): the block-level tree
// carries all NQ lists of a warp per round, every block raises the per-query floor (floor_key[q]) before its
// ticket, and in the last block the warps split the per-block lists by query (warp w: query w % NQ, every
// (n_warps / NQ)-th block list, starting from the query's floor) before one warp per query folds those partial
// results and writes the result rows of its query.
template <int NQ>
__device__ __forceinline__ void scan_finish_multi(const ScanParams& p, WarpList<1> (&L)[NQ], uint64_t* slist) {
  __shared__ unsigned s_is_last;
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  const int k = p.k;
  const int n_warps = static_cast<int>(blockDim.x >> 5);   // a multiple of NQ
  const int grid = static_cast<int>(gridDim.x);
#pragma unroll
  for (int q = 0; q < NQ; ++q) store_list(L[q], slist + (warp * NQ + q) * k, k, lane);
  __syncthreads();
  // the previous scan's last block may still be merging the per-block lists (and owns the ticket)
  asm volatile("griddepcontrol.wait;" ::: "memory");
  for (int step = 1; step < n_warps; step <<= 1) {
    if ((warp & (2 * step - 1)) == 0 && warp + step < n_warps) {
#pragma unroll
      for (int q = 0; q < NQ; ++q) bitonic_merge_shared(L[q], slist + ((warp + step) * NQ + q) * k, k, lane);
      if (2 * step < n_warps) {
#pragma unroll
        for (int q = 0; q < NQ; ++q) store_list(L[q], slist + (warp * NQ + q) * k, k, lane);
      }
    }
    __syncthreads();
  }
  if (warp == 0) {
    uint64_t kth[NQ];
#pragma unroll
    for (int q = 0; q < NQ; ++q) {
      store_list(L[q], p.partial + (static_cast<size_t>(blockIdx.x) * NQ + q) * k, k, lane);
      kth[q] = L[q].get(k - 1);
    }
    fence_acq_rel_gpu();
    __syncwarp();
    if (lane == 0) {
#pragma unroll
      for (int q = 0; q < NQ; ++q)
        if (kth[q] != 0ull) atomicMax(p.floor_key + q, static_cast<unsigned long long>(kth[q]));
      const unsigned t = atomicAdd(p.ticket, 1u);
      s_is_last = (t == gridDim.x - 1) ? 1u : 0u;
    }
  }
  __syncthreads();
  if (s_is_last == 0u) return;

  // ---- last block
  fence_acq_rel_gpu();
  constexpr int HB = 40;   // 4 warps per query x 40 heads: one L2 round trip for 148 lists, two for 296
  const int q = warp % NQ, part = warp / NQ, parts = n_warps / NQ;
  WarpList<1> M;
  M.clear();
  const uint64_t floor_key = __ldcg(p.floor_key + q);
  uint64_t thr = 0ull;
  for (int b0 = part; b0 < grid; b0 += parts * HB) {
    uint64_t head[HB];
#pragma unroll
    for (int j = 0; j < HB; ++j) {
      const int b = b0 + j * parts;
      head[j] = (b < grid && lane < k) ? load_key<true>(p.partial + (static_cast<size_t>(b) * NQ + q) * k + lane) : 0ull;
    }
    if (b0 == part) thr = floor_key ? floor_key - 1ull : 0ull;   // keys >= floor qualify
#pragma unroll
    for (int j = 0; j < HB; ++j) {
      unsigned m = __ballot_sync(0xffffffffu, head[j] > thr);
      while (m) {
        const int srcl = __ffs(m) - 1;
        m &= m - 1;
        const uint64_t x = shfl_u64(head[j], srcl);
        if (x > thr) {
          M.insert(x, lane);
          const uint64_t kth = M.get(k - 1);
          if (kth > thr) thr = kth;   // (never below the floor)
        }
      }
    }
  }
  store_list(M, slist + (q * parts + part) * k, k, lane);   // (the tree's last barrier is behind every warp)
  __syncthreads();
  if (warp < NQ) {   // warp w finishes query w
    M.slot[0] = (lane < k) ? slist[(warp * parts) * k + lane] : 0ull;
    for (int p2 = 1; p2 < parts; ++p2) bitonic_merge_shared(M, slist + (warp * parts + p2) * k, k, lane);
    if (warp < p.nq && lane < k) {
      const uint64_t key = M.slot[0];
      const int64_t o = p.qsel[warp] * k + lane;
      p.out_scores[o] = key ? key_score(key) : -INFINITY;
      p.out_rows[o] = key ? p.row_base + static_cast<int64_t>(key_row(key)) : -1ll;
    }
  }
  if (threadIdx.x == 0) {
    *p.ticket = 0u;
#pragma unroll
    for (int q2 = 0; q2 < NQ; ++q2) p.floor_key[q2] = 0ull;
  }
}

constexpr int kScanMultiThreads = 512;  // x 1 block per SM: 128 registers per thread (NQ accumulators, lists, thresholds)
constexpr int kScanMultiWarps = kScanMultiThreads / 32;
#ifndef PVDB_SCAN_MULTI_LOADS
#define PVDB_SCAN_MULTI_LOADS 16
#endif
constexpr int kScanMultiLoads = PVDB_SCAN_MULTI_LOADS;   // 16-byte loads a lane keeps in flight per step: with half the
                                                         // warps of the single-query kernel per SM, twice its 8

// fp32 rows x NQ queries on the CUDA cores; the walk and the per-lane arithmetic are scan_topk_kernel's.
template <int LPR, int CH, bool SPARSE, int NQ>
__global__ void __launch_bounds__(kScanMultiThreads, 1) scan_multi_topk_kernel(const ScanParams p) {
  constexpr int G = 32 / LPR;  // row groups per warp
  constexpr int R = (kScanMultiLoads / CH) * G <= 32 ? kScanMultiLoads / CH : 32 / G;   // rows in flight per group
  constexpr int RPW = G * R;   // rows per warp step: one bitmap word covers it
  static_assert(kScanMultiWarps % NQ == 0, "the last block splits its warps evenly over the queries");
  extern __shared__ __align__(16) unsigned char smem_raw[];
  float* sq = reinterpret_cast<float*>(smem_raw);
  uint64_t* slist = reinterpret_cast<uint64_t*>(smem_raw + static_cast<size_t>(NQ) * p.query_floats * sizeof(float));
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  const int sub = lane % LPR;
  const int gi = lane / LPR;
  const int k = p.k;
  const int nq = p.nq;

  scan_prologue_multi<NQ>(p, sq);
  const float4* sq4 = reinterpret_cast<const float4*>(sq);
  const int q_chunks = p.query_floats / 4;   // float4 per query

  WarpList<1> L[NQ];
  uint64_t thr[NQ];
#pragma unroll
  for (int q = 0; q < NQ; ++q) {
    L[q].clear();
    thr[q] = 0ull;
  }

  const int64_t total_warps = static_cast<int64_t>(gridDim.x) * kScanMultiWarps;
  const uint4* mat = reinterpret_cast<const uint4*>(p.matrix);
  const int row_chunks = p.row_chunks;

  auto score_rows = [&](const int64_t (&row)[R], const bool (&on)[R]) {
    float acc[NQ][R];
    const uint4* rp[R];
#pragma unroll
    for (int r = 0; r < R; ++r) {
      rp[r] = mat + row[r] * row_chunks;
#pragma unroll
      for (int q = 0; q < NQ; ++q) acc[q][r] = 0.f;
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
#pragma unroll
          for (int q = 0; q < NQ; ++q) {
            const float4 qv = sq4[q * q_chunks + idx];
#pragma unroll
            for (int r = 0; r < R; ++r) acc[q][r] = dot_chunk_f32(v[r][i], qv, acc[q][r]);
          }
        }
      }
    }
    // Reduce the V = NQ * R partial sums over the LPR lanes of the group as a reduce-SCATTER: each butterfly
    // stage halves the values a lane still carries (the lane with the stage's bit set keeps the upper half), so
    // the stages cost V/2 + V/4 + ... shuffles instead of V each (16 instead of 80 at dim 384), and the
    // candidate tests below run once per query, not once per (query, row).  Every total is formed by the same
    // additions as the plain butterfly of the single-query kernel (own + partner at xor LPR/2, LPR/4, ...).
    constexpr int V = NQ * R;
    float vals[V];
#pragma unroll
    for (int q = 0; q < NQ; ++q) {
#pragma unroll
      for (int r = 0; r < R; ++r) vals[q * R + r] = acc[q][r];
    }
    constexpr int LL = LPR == 32 ? 5 : (LPR == 16 ? 4 : 3);
    constexpr int LV = V == 64 ? 6 : (V == 32 ? 5 : (V == 16 ? 4 : (V == 8 ? 3 : (V == 4 ? 2 : (V == 2 ? 1 : 0)))));
    static_assert((1 << LV) == V, "NQ * R must be a power of two");
    constexpr int NS = LL < LV ? LL : LV;          // halving stages
    constexpr int HELD = 1 << (LV - NS);           // totals a lane ends up with
#pragma unroll
    for (int st = 0; st < LL; ++st) {
      const int m = LPR >> (st + 1);
      if (st < NS) {
        const int h = V >> (st + 1);
        const bool upper = (lane & m) != 0;
#pragma unroll
        for (int i = 0; i < h; ++i) {
          const float send = upper ? vals[i] : vals[i + h];
          const float keep = upper ? vals[i + h] : vals[i];
          vals[i] = keep + __shfl_xor_sync(0xffffffffu, send, m);
        }
      } else {
        vals[0] += __shfl_xor_sync(0xffffffffu, vals[0], m);
      }
    }
    // total j of this lane is value (top << (LV - NS)) | j, top = the NS high bits of the lane's index in its
    // group; values are numbered q * R + r; when V < LPR a total is replicated on LPR / V lanes (the first counts)
    const int top = sub >> (LL - NS);
    const bool primary = (sub & ((1 << (LL - NS)) - 1)) == 0;
    uint64_t key[HELD];
    int kq[HELD];
#pragma unroll
    for (int j = 0; j < HELD; ++j) {
      const int v = (top << (LV - NS)) | j;
      const int r = v & (R - 1);
      kq[j] = v / R;
      bool onv = false;
      int64_t rowv = 0;
#pragma unroll
      for (int rr = 0; rr < R; ++rr) {
        if (r == rr) {
          onv = on[rr];
          rowv = row[rr];
        }
      }
      const float sc = vals[j];
      key[j] = (primary && onv && sc == sc) ? make_key(sc, static_cast<uint32_t>(rowv)) : 0ull;
    }
#pragma unroll
    for (int q = 0; q < NQ; ++q) {
      if (q >= nq) break;   // (warp-uniform) the last launch of a call may carry fewer than NQ queries
#pragma unroll
      for (int j = 0; j < HELD; ++j) {
        unsigned m = __ballot_sync(0xffffffffu, kq[j] == q && key[j] > thr[q]);
        while (m) {
          const int srcl = __ffs(m) - 1;
          m &= m - 1;
          const uint64_t x = shfl_u64(key[j], srcl);
          if (x > thr[q]) {
            L[q].insert(x, lane);
            thr[q] = L[q].get(k - 1);
          }
        }
      }
    }
  };

  if constexpr (!SPARSE) {
    const int64_t n_steps = (p.n_rows + RPW - 1) / RPW;
    for (int64_t step = static_cast<int64_t>(blockIdx.x) * kScanMultiWarps + warp; step < n_steps; step += total_warps) {
      const int64_t base = step * RPW;
      uint32_t w = __ldg(p.active + (base >> 5));
      if (p.prefilter) w &= __ldg(p.prefilter + (base >> 5));
      w >>= (base & 31);
      if constexpr (RPW < 32) w &= (1u << RPW) - 1u;
      if (w == 0u) continue;
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
    __shared__ uint32_t s_ring[kScanMultiWarps][64];
    sparse_walk<RPW>(p, static_cast<int64_t>(blockIdx.x) * kScanMultiWarps + warp, total_warps, lane, s_ring[warp],
                     [&](const uint32_t* ring, unsigned head, unsigned avail) {
                       int64_t row[R];
                       bool on[R];
#pragma unroll
                       for (int r = 0; r < R; ++r) {
                         const unsigned slot = r * G + gi;
                         on[r] = slot < avail;
                         row[r] = on[r] ? ring[(head + slot) & 63u] : 0u;
                       }
                       score_rows(row, on);
                     });
  }
  scan_finish_multi<NQ>(p, L, slist);
}

// LPR lanes cooperate on one row; each lane keeps CH 16-byte loads of R rows in flight
// (CH * R == 8 -> eight independent 128-bit loads per lane per step).
// S: 32-key slots of the per-warp list (1 for k <= 32 -- cheaper inserts and merges, fewer registers;
// 4 for k <= 128).
template <bool BF16, int LPR, int CH, bool SPARSE, int S>
__global__ void __launch_bounds__(kScanThreads, kScanBlocksPerSM) scan_topk_kernel(const ScanParams p) {
  constexpr int G = 32 / LPR;  // row groups per warp
  constexpr int R = 8 / CH;    // rows in flight per group
  constexpr int RPW = G * R;   // rows per warp step: a power of two <= 32, so one bitmap word covers it
  extern __shared__ __align__(16) unsigned char smem_raw[];
  float* sq = reinterpret_cast<float*>(smem_raw);
  uint64_t* slist = reinterpret_cast<uint64_t*>(smem_raw + static_cast<size_t>(p.query_floats) * sizeof(float));
  const int tid = threadIdx.x;
  const int lane = tid & 31;
  const int warp = tid >> 5;
  const int sub = lane % LPR;
  const int gi = lane / LPR;
  const int k = p.k;

  PVDB_TRACE(0, blockIdx.x == 0);
  scan_prologue(p, sq);
  PVDB_TRACE(1, blockIdx.x == 0);
  const float4* sq4 = reinterpret_cast<const float4*>(sq);

  const uint64_t upper = p.upper ? *p.upper : ~0ull;
  WarpList<S> L;
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
    // Sparse walk (selective prefilters): only the SET bits of the bitmap are packed into the RPW row slots
    // (sparse_walk above).
    __shared__ uint32_t s_ring[kScanWarps][64];
    sparse_walk<RPW>(p, static_cast<int64_t>(blockIdx.x) * kScanWarps + warp, total_warps, lane, s_ring[warp],
                     [&](const uint32_t* ring, unsigned head, unsigned avail) {
                       int64_t row[R];
                       bool on[R];
#pragma unroll
                       for (int r = 0; r < R; ++r) {
                         const unsigned slot = r * G + gi;
                         on[r] = slot < avail;
                         row[r] = on[r] ? ring[(head + slot) & 63u] : 0u;
                       }
                       score_rows(row, on);
                     });
  }

  PVDB_TRACE(2, blockIdx.x == 0);
  scan_finish<S, (S == 1 ? 10 : 8)>(p, L, thr, slist);
}



// ---------------------------------------------------------------------------- bf16 rows on the legacy tensor-core path
// scan_mma_topk_kernel: the same single-query scan for a bf16 matrix of dim <= 512, with the dot products
// on mma.sync.m16n8k16 (bf16 x bf16 -> fp32).  WHY: the scan is HBM bound, but under the 1 kW power cap the
// SM clock -- and with it what the chip sustains from HBM -- depends on how many instructions ride on each
// byte (tools/micro/hbm_read_sustained.cu, random data, 600 back-to-back launches: 8 FMA per 16-byte chunk
// 6.90 TB/s, 16: 6.81, 32: 6.25).  The CUDA-core form spends ~22 instructions per chunk (8 shifts / masks to
// widen bf16, 8 FMA, query reads from shared memory) and sustains 6.0-6.4 TB/s on the 100M x 384 store; this
// form spends one load and one MMA per chunk.
//   * A warp step scores 16 rows.  Thread (g = lane / 4, t = lane % 4) loads, for each 32-column slice, the
//     16 bytes at columns [8 t, 8 t + 8) of row g and of row g + 8: exactly the A fragments of TWO MMAs (the
//     k index inside a fragment is only a label, so "logical k = 2t, 2t+1, 2t+8, 2t+9" is mapped to this
//     thread's four consecutive columns; the query fragments use the same mapping).
//   * EXACTNESS.  The fp32 query is split into three bf16 terms q = hi + mid + lo (24 mantissa bits: exact)
//     which ride in columns 0, 1, 2 of the B operand; the other five columns are zero.  Every product
//     bf16 x bf16 is exact in fp32 and the accumulators are fp32, so a score differs from the CUDA-core
//     kernel's only by the order of the fp32 additions (~1e-7 for unit vectors; the parity tests hold it to the
//     fp32 tolerance 1e-5 / 2e-6).
//   * The query fragments stay in REGISTERS (4 per slice: no shared-memory traffic in the loop) for
//     dim <= 512; wider rows (KS == 0) read them from shared memory, one 16-byte load per slice and thread
//     (half the shared-memory bytes of the CUDA-core kernel and none of its arithmetic).
// KS: 32-column slices held in registers (dim <= 32 KS); 0 = any width, fragments in shared memory.
__device__ __forceinline__ void mma_bf16_16816(float (&d)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0,
                                               uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
      : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
      : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}

// part 0 / 1 / 2 of the three-term bf16 split of x (bit pattern in the low 16 bits)
__device__ __forceinline__ uint32_t bf16_split_part(float x, int part) {
  const __nv_bfloat16 hi = __float2bfloat16_rn(x);
  const float r1 = x - __bfloat162float(hi);
  const __nv_bfloat16 mid = __float2bfloat16_rn(r1);
  const float r2 = r1 - __bfloat162float(mid);
  const __nv_bfloat16 lo = __float2bfloat16_rn(r2);
  const __nv_bfloat16 pick = part == 0 ? hi : (part == 1 ? mid : lo);
  return static_cast<uint32_t>(__bfloat16_as_ushort(pick));
}

constexpr int kScanMmaThreads = 256;   // x 2 blocks per SM: up to 128 registers per thread (48-64 hold the query fragments)
constexpr int kScanMmaWarps = kScanMmaThreads / 32;

// NQ: queries per pass -- 1, or 2 (S == 1): the second query's three terms ride in B columns 3-5, which the
// single-query form leaves zero, so two queries cost the instructions of one ("several queries per pass" below).
template <bool SPARSE, int S, int KS, int NQ>
__global__ void __launch_bounds__(kScanMmaThreads, kScanBlocksPerSM) scan_mma_topk_kernel(const ScanParams p) {
  static_assert(NQ == 1 || (NQ == 2 && S == 1), "two queries per pass keep one 32-key slot per list");
  constexpr int RPW = 16;  // rows per warp step (the M of the MMA)
  constexpr bool QS = KS == 0;   // query fragments in shared memory
  constexpr int KR = QS ? 1 : KS;
  constexpr int NP = 3 * NQ;     // B columns in use: part (g % 3) of query (g / 3)
  extern __shared__ __align__(16) unsigned char smem_raw[];
  float* sq = reinterpret_cast<float*>(smem_raw);
  uint64_t* slist = reinterpret_cast<uint64_t*>(smem_raw + static_cast<size_t>(NQ) * p.query_floats * sizeof(float));
  // QS: [NP parts][n_slices * 16] packed bf16 pairs behind the lists (8 warps x k keys x 8 bytes: 16-byte aligned)
  static_assert((kScanMmaWarps * sizeof(uint64_t)) % 16 == 0, "the fragment words must stay 16-byte aligned");
  uint32_t* sfrag = reinterpret_cast<uint32_t*>(slist + static_cast<size_t>(NQ) * kScanMmaWarps * p.k);
  const int n_slices = (p.row_chunks + 3) / 4;
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  const int g = lane >> 2;   // fragment row (and the B column this thread feeds)
  const int t = lane & 3;    // which 8 columns of every 32-column slice
  const int k = p.k;

  if constexpr (NQ == 1) scan_prologue(p, sq);
  else scan_prologue_multi<NQ>(p, sq);

  // query fragments: slice ks, this thread's columns c0 = 32 ks + 8 t ... c0 + 7 as four bf16 pairs of
  // part g % 3 of the split of query g / 3 (threads with g >= NP feed the unused B columns: zeros)
  const float* sqg = sq + (g < NP ? g / 3 : 0) * p.query_floats;
  uint32_t bq[KR][4];
  if constexpr (!QS) {
#pragma unroll
    for (int ks = 0; ks < KS; ++ks) {
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int c = ks * 32 + t * 8 + 2 * j;
        uint32_t w = 0u;
        if (g < NP && c < p.query_floats) {
          const float x0 = sqg[c];
          const float x1 = (c + 1 < p.query_floats) ? sqg[c + 1] : 0.f;
          w = bf16_split_part(x0, g % 3) | (bf16_split_part(x1, g % 3) << 16);
        }
        bq[ks][j] = w;
      }
    }
  } else {
    // word i of part pt covers columns 2 i, 2 i + 1 of query pt / 3
    const int words = n_slices * 16;
    for (int i = threadIdx.x; i < NP * words; i += kScanMmaThreads) {
      const int part = i / words, c = 2 * (i - part * words);
      const float* sqp = sq + (part / 3) * p.query_floats;
      const float x0 = c < p.query_floats ? sqp[c] : 0.f;
      const float x1 = c + 1 < p.query_floats ? sqp[c + 1] : 0.f;
      sfrag[i] = bf16_split_part(x0, part % 3) | (bf16_split_part(x1, part % 3) << 16);
    }
    __syncthreads();
  }
  // this thread's fragment words of slice ks sit at sfrag4[ks * 4] (threads with g >= NP feed zero columns)
  const uint4* sfrag4 = reinterpret_cast<const uint4*>(sfrag + (g < NP ? g : 0) * n_slices * 16) + t;

  const uint64_t upper = p.upper ? *p.upper : ~0ull;
  WarpList<S> L[NQ];
  uint64_t thr[NQ];
#pragma unroll
  for (int q = 0; q < NQ; ++q) {
    L[q].clear();
    thr[q] = 0ull;
  }

  const int64_t total_warps = static_cast<int64_t>(gridDim.x) * kScanMmaWarps;
  const uint4* mat = reinterpret_cast<const uint4*>(p.matrix);
  const int row_chunks = p.row_chunks;
  const bool full_rows = row_chunks == 4 * KR;   // (KS > 0) the row is exactly KS slices wide

  // Score the 16 row slots of a warp step (this thread: slots g and g + 8) and feed the warp list.
  auto score_rows = [&](const int64_t (&row)[2], const bool (&on)[2]) {
    const uint4* rp0 = mat + row[0] * row_chunks + t;
    const uint4* rp1 = mat + row[1] * row_chunks + t;
    float ca[4] = {0.f, 0.f, 0.f, 0.f}, cb[4] = {0.f, 0.f, 0.f, 0.f};
    if constexpr (!QS) {
      if (full_rows && __all_sync(0xffffffffu, on[0] && on[1])) {
        // the common case -- every slot holds a live row and the row fills its KS slices exactly: plain loads
        // with immediate offsets, nothing predicated (the general loop below spends ~4 instructions per load
        // on predicates and zero fills)
#pragma unroll
        for (int ks0 = 0; ks0 < KS; ks0 += 4) {
          uint4 v0[4], v1[4];
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            if (ks0 + i < KS) {
              v0[i] = ldg_stream(rp0 + (ks0 + i) * 4);
              v1[i] = ldg_stream(rp1 + (ks0 + i) * 4);
            }
          }
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const int ks = ks0 + i;
            if (ks < KS) {
              mma_bf16_16816(ca, v0[i].x, v1[i].x, v0[i].y, v1[i].y, bq[ks][0], bq[ks][1]);
              mma_bf16_16816(cb, v0[i].z, v1[i].z, v0[i].w, v1[i].w, bq[ks][2], bq[ks][3]);
            }
          }
        }
      } else {
#pragma unroll
      for (int ks0 = 0; ks0 < KS; ks0 += 4) {
        if (ks0 * 4 >= row_chunks) break;   // (warp-uniform) slices past the row
        uint4 v0[4], v1[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int ks = ks0 + i;
          const bool in = ks < KS && ks * 4 + t < row_chunks;
          v0[i] = (on[0] && in) ? ldg_stream(rp0 + ks * 4) : make_uint4(0u, 0u, 0u, 0u);
          v1[i] = (on[1] && in) ? ldg_stream(rp1 + ks * 4) : make_uint4(0u, 0u, 0u, 0u);
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int ks = ks0 + i;
          if (ks < KS) {
            mma_bf16_16816(ca, v0[i].x, v1[i].x, v0[i].y, v1[i].y, bq[ks][0], bq[ks][1]);
            mma_bf16_16816(cb, v0[i].z, v1[i].z, v0[i].w, v1[i].w, bq[ks][2], bq[ks][3]);
          }
        }
      }
      }
    } else {
      for (int ks0 = 0; ks0 < n_slices; ks0 += 4) {
        uint4 v0[4], v1[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int ks = ks0 + i;
          const bool in = ks * 4 + t < row_chunks;
          v0[i] = (on[0] && in) ? ldg_stream(rp0 + ks * 4) : make_uint4(0u, 0u, 0u, 0u);
          v1[i] = (on[1] && in) ? ldg_stream(rp1 + ks * 4) : make_uint4(0u, 0u, 0u, 0u);
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int ks = ks0 + i;
          if (ks < n_slices) {
            const uint4 f = g < NP ? sfrag4[ks * 4] : make_uint4(0u, 0u, 0u, 0u);
            mma_bf16_16816(ca, v0[i].x, v1[i].x, v0[i].y, v1[i].y, f.x, f.y);
            mma_bf16_16816(cb, v0[i].z, v1[i].z, v0[i].w, v1[i].w, f.z, f.w);
          }
        }
      }
    }
    // thread t == 0 holds (hi, mid) of rows g / g + 8 in c[0..1] / c[2..3]; thread t == 1 holds lo in c[0] / c[2]
    const float c0 = ca[0] + cb[0], c1 = ca[1] + cb[1], c2 = ca[2] + cb[2], c3 = ca[3] + cb[3];
    const float lo0 = __shfl_down_sync(0xffffffffu, c0, 1), lo1 = __shfl_down_sync(0xffffffffu, c2, 1);
    float sc[2] = {(lo0 + c1) + c0, (lo1 + c3) + c2};
    if constexpr (NQ == 2) {
      // second query: hi in c[1] / c[3] of thread t == 1, (mid, lo) in c[0..1] / c[2..3] of thread t == 2; the
      // same (lo + mid) + hi as above, so a query scores to the same bits in either place
      const float d1 = __shfl_down_sync(0xffffffffu, c1, 1), d3 = __shfl_down_sync(0xffffffffu, c3, 1);
      if (t == 1) {
        sc[0] = (d1 + lo0) + c1;
        sc[1] = (d3 + lo1) + c3;
      }
    }
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      const uint64_t key = (on[r] && sc[r] == sc[r]) ? make_key(sc[r], static_cast<uint32_t>(row[r])) : 0ull;
#pragma unroll
      for (int q = 0; q < NQ; ++q) {
        if (NQ > 1 && q >= p.nq) break;   // (warp-uniform) the last launch of a call may carry one query
        unsigned m = __ballot_sync(0xffffffffu, t == q && key > thr[q] && key < upper);
        while (m) {
          const int srcl = __ffs(m) - 1;
          m &= m - 1;
          const uint64_t x = shfl_u64(key, srcl);
          if (x > thr[q]) {
            L[q].insert(x, lane);
            thr[q] = L[q].get(k - 1);
          }
        }
      }
    }
  };

  if constexpr (!SPARSE) {
    // dense walk: a warp step covers 16 consecutive rows (half a bitmap word)
    const int64_t n_steps = (p.n_rows + RPW - 1) / RPW;
    for (int64_t step = static_cast<int64_t>(blockIdx.x) * kScanMmaWarps + warp; step < n_steps; step += total_warps) {
      const int64_t base = step * RPW;
      uint32_t w = __ldg(p.active + (base >> 5));
      if (p.prefilter) w &= __ldg(p.prefilter + (base >> 5));
      w = (w >> (base & 31)) & 0xffffu;
      if (w == 0u) continue;  // every row of this step is deleted / filtered out: read nothing
      const int64_t row[2] = {base + g, base + g + 8};
      const bool on[2] = {((w >> g) & 1u) != 0u, ((w >> (g + 8)) & 1u) != 0u};
      score_rows(row, on);
    }
  } else {
    // sparse walk (selective prefilters): only the SET bits are packed into the 16 row slots (sparse_walk above)
    __shared__ uint32_t s_ring[kScanMmaWarps][64];
    sparse_walk<RPW>(p, static_cast<int64_t>(blockIdx.x) * kScanMmaWarps + warp, total_warps, lane, s_ring[warp],
                     [&](const uint32_t* ring, unsigned head, unsigned avail) {
                       const bool on[2] = {static_cast<unsigned>(g) < avail, static_cast<unsigned>(g + 8) < avail};
                       const int64_t row[2] = {on[0] ? ring[(head + g) & 63u] : 0u, on[1] ? ring[(head + g + 8) & 63u] : 0u};
                       score_rows(row, on);
                     });
  }
  if constexpr (NQ == 1) scan_finish<S, (S == 1 ? 20 : 8)>(p, L[0], thr[0], slist);
  else scan_finish_multi<NQ>(p, L, slist);
}

template <bool SPARSE, int S, int KS, int NQ = 1>
static int launch_scan_mma_inst(const ScanParams& p, cudaStream_t stream) {
  size_t smem = static_cast<size_t>(NQ) * p.query_floats * sizeof(float) +
                static_cast<size_t>(NQ) * kScanMmaWarps * p.k * sizeof(uint64_t);
  if (KS == 0) smem += static_cast<size_t>(3 * NQ) * ((p.row_chunks + 3) / 4) * 16 * sizeof(uint32_t);
  auto kern = scan_mma_topk_kernel<SPARSE, S, KS, NQ>;
  if (smem > 48 * 1024)
    PVDB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
  if (p.pdl) {
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(kNumSMs * kScanBlocksPerSM);
    cfg.blockDim = dim3(kScanMmaThreads);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    PVDB_CUDA(cudaLaunchKernelEx(&cfg, kern, p));
  } else {
    kern<<<kNumSMs * kScanBlocksPerSM, kScanMmaThreads, smem, stream>>>(p);
  }
  PVDB_LAUNCH_CHECK();
  return PVDB_OK;
}

// bf16 rows: up to 512 columns with the query fragments in registers, wider rows with them in shared memory
template <bool SPARSE>
int launch_scan_mma_variant(const ScanParams& p, cudaStream_t stream) {
  const int ks = (p.row_chunks + 3) / 4;
  if (p.k <= 32) {
    if (ks <= 4) return launch_scan_mma_inst<SPARSE, 1, 4>(p, stream);
    if (ks <= 8) return launch_scan_mma_inst<SPARSE, 1, 8>(p, stream);
    if (ks <= 12) return launch_scan_mma_inst<SPARSE, 1, 12>(p, stream);
    if (ks <= 16) return launch_scan_mma_inst<SPARSE, 1, 16>(p, stream);
    return launch_scan_mma_inst<SPARSE, 1, 0>(p, stream);
  }
  if (ks <= 4) return launch_scan_mma_inst<SPARSE, 4, 4>(p, stream);
  if (ks <= 8) return launch_scan_mma_inst<SPARSE, 4, 8>(p, stream);
  if (ks <= 12) return launch_scan_mma_inst<SPARSE, 4, 12>(p, stream);
  if (ks <= 16) return launch_scan_mma_inst<SPARSE, 4, 16>(p, stream);
  return launch_scan_mma_inst<SPARSE, 4, 0>(p, stream);
}

// Launch the (LPR, CH, S) instantiation for this translation unit's (BF16, SPARSE).
template <bool BF16, bool SPARSE, int LPR, int CH, int S>
static int launch_scan_inst(const ScanParams& p, cudaStream_t stream) {
  const size_t smem = static_cast<size_t>(p.query_floats) * sizeof(float) +
                      static_cast<size_t>(kScanWarps) * p.k * sizeof(uint64_t);
  auto kern = scan_topk_kernel<BF16, LPR, CH, SPARSE, S>;
  if (smem > 48 * 1024)
    PVDB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
  if (p.pdl) {
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(kNumSMs * kScanBlocksPerSM);
    cfg.blockDim = dim3(kScanThreads);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    PVDB_CUDA(cudaLaunchKernelEx(&cfg, kern, p));
  } else {
    kern<<<kNumSMs * kScanBlocksPerSM, kScanThreads, smem, stream>>>(p);
  }
  PVDB_LAUNCH_CHECK();
  return PVDB_OK;
}

template <bool BF16, bool SPARSE, int LPR, int S>
static int launch_scan_ch(const ScanParams& p, int ch, cudaStream_t stream) {
  switch (ch) {
    case 1: return launch_scan_inst<BF16, SPARSE, LPR, 1, S>(p, stream);
    case 2: return launch_scan_inst<BF16, SPARSE, LPR, 2, S>(p, stream);
    case 4: return launch_scan_inst<BF16, SPARSE, LPR, 4, S>(p, stream);
    default: return launch_scan_inst<BF16, SPARSE, LPR, 8, S>(p, stream);
  }
}

template <bool BF16, bool SPARSE>
int launch_scan_variant(const ScanParams& p, int lpr, int ch, cudaStream_t stream) {
  if (p.k <= 32) {
    if (lpr == 32) return launch_scan_ch<BF16, SPARSE, 32, 1>(p, ch, stream);
    if (lpr == 16) return launch_scan_ch<BF16, SPARSE, 16, 1>(p, ch, stream);
    return launch_scan_ch<BF16, SPARSE, 8, 1>(p, ch, stream);
  }
  if (lpr == 32) return launch_scan_ch<BF16, SPARSE, 32, 4>(p, ch, stream);
  if (lpr == 16) return launch_scan_ch<BF16, SPARSE, 16, 4>(p, ch, stream);
  return launch_scan_ch<BF16, SPARSE, 8, 4>(p, ch, stream);
}


// ---------------------------------------------------------------------------- several queries per pass: launchers
constexpr int kScanMultiF32 = 4;    // queries per pass, fp32 rows
constexpr int kScanMultiBf16 = 2;   // queries per pass, bf16 rows (mma form)

template <typename Kern>
static int launch_scan_multi_kernel(Kern kern, const ScanParams& p, int blocks, int threads, size_t smem, cudaStream_t stream) {
  if (smem > 48 * 1024)
    PVDB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(blocks);
  cfg.blockDim = dim3(threads);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = p.pdl ? 1 : 0;
  PVDB_CUDA(cudaLaunchKernelEx(&cfg, kern, p));
  PVDB_LAUNCH_CHECK();
  return PVDB_OK;
}

template <bool SPARSE, int LPR>
static int launch_scan_multi_ch(const ScanParams& p, int ch, cudaStream_t stream) {
  constexpr int NQ = kScanMultiF32;
  const size_t smem = static_cast<size_t>(NQ) * p.query_floats * sizeof(float) +
                      static_cast<size_t>(NQ) * kScanMultiWarps * p.k * sizeof(uint64_t);
  switch (ch) {
    case 1: return launch_scan_multi_kernel(scan_multi_topk_kernel<LPR, 1, SPARSE, NQ>, p, kNumSMs, kScanMultiThreads, smem, stream);
    case 2: return launch_scan_multi_kernel(scan_multi_topk_kernel<LPR, 2, SPARSE, NQ>, p, kNumSMs, kScanMultiThreads, smem, stream);
    case 4: return launch_scan_multi_kernel(scan_multi_topk_kernel<LPR, 4, SPARSE, NQ>, p, kNumSMs, kScanMultiThreads, smem, stream);
    default: return launch_scan_multi_kernel(scan_multi_topk_kernel<LPR, 8, SPARSE, NQ>, p, kNumSMs, kScanMultiThreads, smem, stream);
  }
}

// fp32 rows, kScanMultiF32 queries per pass
template <bool SPARSE>
int launch_scan_multi_variant(const ScanParams& p, int lpr, int ch, cudaStream_t stream) {
  if (lpr == 32) return launch_scan_multi_ch<SPARSE, 32>(p, ch, stream);
  if (lpr == 16) return launch_scan_multi_ch<SPARSE, 16>(p, ch, stream);
  return launch_scan_multi_ch<SPARSE, 8>(p, ch, stream);
}

// bf16 rows on mma.sync, kScanMultiBf16 queries per pass
template <bool SPARSE>
int launch_scan_mma_multi_variant(const ScanParams& p, cudaStream_t stream) {
  constexpr int NQ = kScanMultiBf16;
  const int ks = (p.row_chunks + 3) / 4;
  size_t smem = static_cast<size_t>(NQ) * p.query_floats * sizeof(float) +
                static_cast<size_t>(NQ) * kScanMmaWarps * p.k * sizeof(uint64_t);
  const int blocks = kNumSMs * kScanBlocksPerSM;
  if (ks <= 4) return launch_scan_multi_kernel(scan_mma_topk_kernel<SPARSE, 1, 4, NQ>, p, blocks, kScanMmaThreads, smem, stream);
  if (ks <= 8) return launch_scan_multi_kernel(scan_mma_topk_kernel<SPARSE, 1, 8, NQ>, p, blocks, kScanMmaThreads, smem, stream);
  if (ks <= 12) return launch_scan_multi_kernel(scan_mma_topk_kernel<SPARSE, 1, 12, NQ>, p, blocks, kScanMmaThreads, smem, stream);
  if (ks <= 16) return launch_scan_multi_kernel(scan_mma_topk_kernel<SPARSE, 1, 16, NQ>, p, blocks, kScanMmaThreads, smem, stream);
  smem += static_cast<size_t>(3 * NQ) * ks * 16 * sizeof(uint32_t);
  return launch_scan_multi_kernel(scan_mma_topk_kernel<SPARSE, 1, 0, NQ>, p, blocks, kScanMmaThreads, smem, stream);
}

}  // namespace pvdb
