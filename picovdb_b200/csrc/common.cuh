// Shared device/host helpers for the picovdb_b200 kernels (sm_100a only).
#pragma once

#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include <atomic>
#include <cstdarg>
#include <cstdio>
#include <string>

#include "../../include/picovdb_b200.h"

namespace pvdb {

// ---------------------------------------------------------------------------- error plumbing
extern thread_local std::string g_last_error;
extern std::atomic<long long> g_launches;

inline int fail(int code, const char* fmt, ...) {
  char buf[512];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  g_last_error = buf;
  return code;
}

#define PVDB_CUDA(expr)                                                                    \
  do {                                                                                     \
    cudaError_t _e = (expr);                                                               \
    if (_e != cudaSuccess) {                                                               \
      (void)cudaGetLastError();                                                            \
      return ::pvdb::fail(_e == cudaErrorMemoryAllocation ? PVDB_ERR_OOM : PVDB_ERR_CUDA,  \
                          "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, \
                          __LINE__);                                                       \
    }                                                                                      \
  } while (0)

#define PVDB_TRY(expr)            \
  do {                            \
    int _rc = (expr);             \
    if (_rc != PVDB_OK) return _rc; \
  } while (0)

#define PVDB_LAUNCH_CHECK()                         \
  do {                                              \
    ::pvdb::g_launches.fetch_add(1);                \
    PVDB_CUDA(cudaGetLastError());                  \
  } while (0)

constexpr int kNumSMs = 148;  // B200: 2 dies x 74 SMs; grids are sized in multiples of this
constexpr int kFusedK = 128;  // largest k one scan pass selects; larger k pages (see scan.cu)

// ---------------------------------------------------------------------------- ordered keys
// A candidate is one 64-bit key: high word = order-preserving image of the fp32 score, low word =
// ~row.  Larger key == better candidate, ties on score resolve to the LOWER row, keys of distinct
// rows are distinct, and key 0 is never produced by a finite or infinite score ("empty").
__host__ __device__ __forceinline__ uint32_t f32_to_ordered(float f) {
#ifdef __CUDA_ARCH__
  uint32_t u = __float_as_uint(f);
#else
  union { float f; uint32_t u; } c; c.f = f; uint32_t u = c.u;
#endif
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__host__ __device__ __forceinline__ float ordered_to_f32(uint32_t o) {
  uint32_t u = (o & 0x80000000u) ? (o & 0x7fffffffu) : ~o;
#ifdef __CUDA_ARCH__
  return __uint_as_float(u);
#else
  union { float f; uint32_t u; } c; c.u = u; return c.f;
#endif
}
__host__ __device__ __forceinline__ uint64_t make_key(float score, uint32_t row) {
  return (static_cast<uint64_t>(f32_to_ordered(score)) << 32) | static_cast<uint64_t>(0xffffffffu - row);
}
__host__ __device__ __forceinline__ float key_score(uint64_t key) {
  return ordered_to_f32(static_cast<uint32_t>(key >> 32));
}
__host__ __device__ __forceinline__ uint32_t key_row(uint64_t key) {
  return 0xffffffffu - static_cast<uint32_t>(key & 0xffffffffu);
}

#ifdef __CUDACC__
// ---------------------------------------------------------------------------- input rounding
// |x - tf32(x)| when the low 13 mantissa bits are dropped.  Truncation error is an upper bound of the
// round-to-nearest error of the same value, so a bound built from it holds whichever of the two the
// tensor core applies to its fp32 operands.
__device__ __forceinline__ float tf32_trunc_err(float x) {
  return fabsf(x - __uint_as_float(__float_as_uint(x) & 0xffffe000u));
}
// |x - bf16(x)| for the round-to-nearest-even conversion this library uses for the mirror / queries
__device__ __forceinline__ float bf16_rn_err(float x) {
  return fabsf(x - __bfloat162float(__float2bfloat16_rn(x)));
}

// ---------------------------------------------------------------------------- loads
// 128-bit streaming load: read-only path, do not allocate in L1 (every byte is used once).
__device__ __forceinline__ uint4 ldg_stream(const void* p) {
  uint4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
               : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
               : "l"(p));
  return r;
}

__device__ __forceinline__ uint64_t shfl_u64(uint64_t v, int src) {
  uint32_t lo = __shfl_sync(0xffffffffu, static_cast<uint32_t>(v), src);
  uint32_t hi = __shfl_sync(0xffffffffu, static_cast<uint32_t>(v >> 32), src);
  return (static_cast<uint64_t>(hi) << 32) | lo;
}
__device__ __forceinline__ uint64_t shfl_up_u64(uint64_t v, int delta) {
  uint32_t lo = __shfl_up_sync(0xffffffffu, static_cast<uint32_t>(v), delta);
  uint32_t hi = __shfl_up_sync(0xffffffffu, static_cast<uint32_t>(v >> 32), delta);
  return (static_cast<uint64_t>(hi) << 32) | lo;
}
__device__ __forceinline__ uint64_t shfl_xor_u64(uint64_t v, int mask) {
  uint32_t lo = __shfl_xor_sync(0xffffffffu, static_cast<uint32_t>(v), mask);
  uint32_t hi = __shfl_xor_sync(0xffffffffu, static_cast<uint32_t>(v >> 32), mask);
  return (static_cast<uint64_t>(hi) << 32) | lo;
}

// Device-scope fence for the "last block done" hand-over (lists -> fence -> ticket; ticket -> fence -> lists).
// __threadfence() is the sequentially consistent fence (MEMBAR.SC.GPU): the acquire-release one is enough for a
// fence-fence synchronisation through a relaxed atomic and is the cheaper instruction.
__device__ __forceinline__ void fence_acq_rel_gpu() { asm volatile("fence.acq_rel.gpu;" ::: "memory"); }

// ---------------------------------------------------------------------------- warp top-k list
// A warp keeps its best candidates as a descending list of 32*S keys spread over the lanes:
// entry e lives in slot e/32 of lane e%32.  Only the first k entries are ever read back; the tail
// just holds smaller keys.  Insertion is one pass of shuffles, no position search:
//   new[e] = old[e] > x ? old[e] : (old[e-1] > x ? x : old[e-1])        (old[-1] = +inf)
template <int S>
struct WarpList {
  uint64_t slot[S];

  __device__ __forceinline__ void clear() {
#pragma unroll
    for (int s = 0; s < S; ++s) slot[s] = 0ull;
  }

  __device__ __forceinline__ void insert(uint64_t x, int lane) {
#pragma unroll
    for (int s = S - 1; s >= 0; --s) {
      uint64_t cur = slot[s];
      uint64_t prev = shfl_up_u64(cur, 1);
      if (s > 0) {
        uint64_t carry = shfl_u64(slot[s - 1], 31);
        if (lane == 0) prev = carry;
      } else {
        if (lane == 0) prev = ~0ull;
      }
      slot[s] = (cur > x) ? cur : ((prev > x) ? x : prev);
    }
  }

  // key of entry e (warp-uniform e)
  __device__ __forceinline__ uint64_t get(int e) const {
    uint64_t v = slot[0];
#pragma unroll
    for (int s = 1; s < S; ++s)
      if ((e >> 5) == s) v = slot[s];
    return shfl_u64(v, e & 31);
  }
};

// ---------------------------------------------------------------------------- list merging
template <bool GLOBAL>
__device__ __forceinline__ uint64_t load_key(const uint64_t* p) {
  if constexpr (GLOBAL) {
    return __ldcg(reinterpret_cast<const unsigned long long*>(p));  // L2: written by other SMs
  } else {
    return *p;
  }
}

template <bool GLOBAL, int S>
__device__ __forceinline__ void merge_list(WarpList<S>& L, uint64_t& thr, const uint64_t* src, int n, int k,
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

// Merge a descending list of k <= 32 keys in SHARED memory into a one-slot warp list with a bitonic network:
// max(A[i], B[31 - i]) is a bitonic sequence holding the 32 best of the union; five compare-exchange stages sort
// it.  Data independent (~5 x 2 shuffles), unlike the insertion loop of merge_list, whose cost grows with the
// number of keys that qualify (~150 cycles each).
__device__ __forceinline__ void bitonic_merge_shared(WarpList<1>& L, const uint64_t* src, int k, int lane) {
  const int e = 31 - lane;
  const uint64_t b = (e < k) ? src[e] : 0ull;
  uint64_t v = L.slot[0] > b ? L.slot[0] : b;
#pragma unroll
  for (int j = 16; j > 0; j >>= 1) {
    const uint64_t o = shfl_xor_u64(v, j);
    const bool keep_max = (lane & j) == 0;
    v = (keep_max == (v > o)) ? v : o;
  }
  L.slot[0] = v;
}

// The same network for two one-slot lists held in registers (both descending over the lanes).
__device__ __forceinline__ uint64_t bitonic_merge_regs(uint64_t a, uint64_t b, int lane) {
  const uint64_t br = shfl_u64(b, 31 - lane);
  uint64_t v = a > br ? a : br;
#pragma unroll
  for (int j = 16; j > 0; j >>= 1) {
    const uint64_t o = shfl_xor_u64(v, j);
    const bool keep_max = (lane & j) == 0;
    v = (keep_max == (v > o)) ? v : o;
  }
  return v;
}

// Fold the lists the warps of a block left in shared memory (warp w at slist + w * k) into warp 0's list by a
// binary tree: log2(n_warps) rounds, each a block barrier + one merge per surviving warp, instead of warp 0
// merging the other n_warps - 1 lists one after the other.  Every warp of the block must call it.
template <int S>
__device__ __forceinline__ void block_tree_merge(WarpList<S>& L, uint64_t& thr, uint64_t* slist, int k, int warp, int lane,
                                                 int n_warps) {
  for (int step = 1; step < n_warps; step <<= 1) {
    if ((warp & (2 * step - 1)) == 0 && warp + step < n_warps) {
      if constexpr (S == 1) {
        bitonic_merge_shared(L, slist + (warp + step) * k, k, lane);
      } else {
        merge_list<false, S>(L, thr, slist + (warp + step) * k, k, k, lane);
      }
      if (2 * step < n_warps) {   // read by the next round's survivor
#pragma unroll
        for (int s = 0; s < S; ++s) {
          const int e = s * 32 + lane;
          if (e < k) slist[warp * k + e] = L.slot[s];
        }
      }
    }
    __syncthreads();
  }
  if constexpr (S == 1) thr = L.get(k - 1);
}

template <int S>
__device__ __forceinline__ void store_list(const WarpList<S>& L, uint64_t* dst, int k, int lane) {
#pragma unroll
  for (int s = 0; s < S; ++s) {
    const int e = s * 32 + lane;
    if (e < k) dst[e] = L.slot[s];
  }
}

#endif  // __CUDACC__

}  // namespace pvdb
