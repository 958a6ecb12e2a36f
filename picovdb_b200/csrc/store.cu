// Store management + the write-side kernels of the hot path:
//   fused L2-normalise + scatter + set-active  (reference: _normalize + upsert row write,
//                                               picovdb/pico_vdb.py:58-68, 413-472)
//   delete = clear active bit + zero row       (pico_vdb.py:514-531)
//   row fetch / download / raw upload / compaction (pico_vdb.py:945, 356, 233-259, 840-848)
#include <cuda.h>

#include <fcntl.h>
#include <unistd.h>

#include <algorithm>
#include <cerrno>
#include <cstdlib>
#include <cstring>
#include <thread>
#include <vector>

#include "store.cuh"

namespace pvdb {

thread_local std::string g_last_error;
std::atomic<long long> g_launches{0};

// ---------------------------------------------------------------------------- buffers
// ---- CUDA virtual memory management entry points, resolved at run time (no link-time libcuda) ----
namespace {
struct VmmApi {
  CUresult (*reserve)(CUdeviceptr*, size_t, size_t, CUdeviceptr, unsigned long long) = nullptr;
  CUresult (*addr_free)(CUdeviceptr, size_t) = nullptr;
  CUresult (*create)(CUmemGenericAllocationHandle*, size_t, const CUmemAllocationProp*, unsigned long long) = nullptr;
  CUresult (*release)(CUmemGenericAllocationHandle) = nullptr;
  CUresult (*map)(CUdeviceptr, size_t, size_t, CUmemGenericAllocationHandle, unsigned long long) = nullptr;
  CUresult (*unmap)(CUdeviceptr, size_t) = nullptr;
  CUresult (*set_access)(CUdeviceptr, size_t, const CUmemAccessDesc*, size_t) = nullptr;
  CUresult (*granularity)(size_t*, const CUmemAllocationProp*, CUmemAllocationGranularity_flags) = nullptr;
  bool ok = false;
};

template <typename F>
bool resolve(const char* name, F& fn) {
  void* ptr = nullptr;
  cudaDriverEntryPointQueryResult q;
  if (cudaGetDriverEntryPoint(name, &ptr, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess) {
    (void)cudaGetLastError();
    return false;
  }
  fn = reinterpret_cast<F>(ptr);
  return true;
}

VmmApi& vmm_api() {
  static VmmApi api;
  static bool tried = false;
  if (!tried) {
    tried = true;
    if (getenv("PVDB_NO_VMM") == nullptr) {
      api.ok = resolve("cuMemAddressReserve", api.reserve) && resolve("cuMemAddressFree", api.addr_free) &&
               resolve("cuMemCreate", api.create) && resolve("cuMemRelease", api.release) &&
               resolve("cuMemMap", api.map) && resolve("cuMemUnmap", api.unmap) &&
               resolve("cuMemSetAccess", api.set_access) &&
               resolve("cuMemGetAllocationGranularity", api.granularity);
      int dev = 0, supported = 0;
      if (api.ok && cudaGetDevice(&dev) == cudaSuccess)
        // CU_DEVICE_ATTRIBUTE_VIRTUAL_MEMORY_MANAGEMENT_SUPPORTED (102); the runtime enum has no name for it
        cudaDeviceGetAttribute(&supported, static_cast<cudaDeviceAttr>(102), dev);
      api.ok = api.ok && supported != 0;
    }
  }
  return api;
}

constexpr size_t kVaReserve = 256ull << 30;  // address range per buffer; physical memory follows demand
}  // namespace

int DeviceBuffer::grow(size_t new_bytes, cudaStream_t stream) {
  if (new_bytes <= bytes) return PVDB_OK;
  VmmApi& api = vmm_api();
  if (api.ok && (ptr == nullptr || vmm)) {
    int dev = 0;
    PVDB_CUDA(cudaGetDevice(&dev));
    CUmemAllocationProp prop{};
    prop.type = CU_MEM_ALLOCATION_TYPE_PINNED;
    prop.location.type = CU_MEM_LOCATION_TYPE_DEVICE;
    prop.location.id = dev;
    size_t gran = 0;
    if (api.granularity(&gran, &prop, CU_MEM_ALLOC_GRANULARITY_RECOMMENDED) != CUDA_SUCCESS || gran == 0)
      return fail(PVDB_ERR_CUDA, "cuMemGetAllocationGranularity failed");
    if (ptr == nullptr) {
      CUdeviceptr base = 0;
      const size_t want = std::max(kVaReserve, (new_bytes + gran - 1) / gran * gran);
      if (api.reserve(&base, want, 0, 0, 0) != CUDA_SUCCESS)
        return fail(PVDB_ERR_OOM, "cuMemAddressReserve(%zu bytes) failed", want);
      ptr = reinterpret_cast<void*>(base);
      va_bytes = want;
      vmm = true;
    }
    const size_t add_total = (new_bytes - bytes + gran - 1) / gran * gran;
    if (bytes + add_total > va_bytes)
      return fail(PVDB_ERR_OOM, "store buffer would exceed its reserved address range (%zu bytes)", va_bytes);
    // physical memory is created in pieces of at most 16 GiB: a 100M x 384 bf16 store on one GPU is a
    // single 77 GB growth step, and one allocation handle of that size is needlessly hard to place
    const size_t piece_max = std::max(gran, (size_t(16) << 30) / gran * gran);
    for (size_t done = 0; done < add_total;) {
      const size_t add = std::min(piece_max, add_total - done);
      CUmemGenericAllocationHandle h = 0;
      if (api.create(&h, add, &prop, 0) != CUDA_SUCCESS)
        return fail(PVDB_ERR_OOM, "cuMemCreate(%zu bytes) failed: out of device memory", add);
      const CUdeviceptr at = reinterpret_cast<CUdeviceptr>(ptr) + bytes;
      if (api.map(at, add, 0, h, 0) != CUDA_SUCCESS) {
        api.release(h);
        return fail(PVDB_ERR_CUDA, "cuMemMap failed");
      }
      CUmemAccessDesc acc{};
      acc.location = prop.location;
      acc.flags = CU_MEM_ACCESS_FLAGS_PROT_READWRITE;
      if (api.set_access(at, add, &acc, 1) != CUDA_SUCCESS) {
        api.unmap(at, add);
        api.release(h);
        return fail(PVDB_ERR_CUDA, "cuMemSetAccess failed");
      }
      chunks.push_back({static_cast<unsigned long long>(h), bytes, add});
      PVDB_CUDA(cudaMemsetAsync(reinterpret_cast<void*>(at), 0, add, stream));
      bytes += add;
      done += add;
    }
    return PVDB_OK;
  }
  // plain allocation (no VMM support): allocate, copy, free
  void* np = nullptr;
  PVDB_CUDA(cudaMalloc(&np, new_bytes));
  if (bytes) PVDB_CUDA(cudaMemcpyAsync(np, ptr, bytes, cudaMemcpyDeviceToDevice, stream));
  PVDB_CUDA(cudaMemsetAsync(static_cast<char*>(np) + bytes, 0, new_bytes - bytes, stream));
  PVDB_CUDA(cudaStreamSynchronize(stream));
  if (ptr) PVDB_CUDA(cudaFree(ptr));
  ptr = np;
  bytes = new_bytes;
  return PVDB_OK;
}

void DeviceBuffer::release() {
  if (vmm) {
    VmmApi& api = vmm_api();
    cudaDeviceSynchronize();
    for (const Chunk& c : chunks) {
      api.unmap(reinterpret_cast<CUdeviceptr>(ptr) + c.offset, c.size);
      api.release(static_cast<CUmemGenericAllocationHandle>(c.handle));
    }
    chunks.clear();
    if (ptr) api.addr_free(reinterpret_cast<CUdeviceptr>(ptr), va_bytes);
    vmm = false;
    va_bytes = 0;
  } else if (ptr) {
    cudaFree(ptr);
  }
  ptr = nullptr;
  bytes = 0;
}

int Scratch::ensure(size_t need) {
  if (need <= bytes) return PVDB_OK;
  size_t nb = std::max(need, bytes + bytes / 2);
  nb = (nb + 255) & ~size_t(255);
  if (ptr) {
    // outstanding work may still read the old block: the caller's stream is synchronised by
    // cudaFree / cudaFreeHost themselves (both are device-synchronising calls).
    if (pinned_host) PVDB_CUDA(cudaFreeHost(ptr)); else PVDB_CUDA(cudaFree(ptr));
    ptr = nullptr;
    bytes = 0;
  }
  if (pinned_host) PVDB_CUDA(cudaMallocHost(&ptr, nb)); else PVDB_CUDA(cudaMalloc(&ptr, nb));
  bytes = nb;
  ++gen;
  return PVDB_OK;
}

void Scratch::release() {
  if (ptr) { if (pinned_host) cudaFreeHost(ptr); else cudaFree(ptr); }
  ptr = nullptr;
  bytes = 0;
}

// ---------------------------------------------------------------------------- kernels
// One warp per input vector.  Pass 1: per-lane fp32 partial sums of squares, fp64 warp reduction,
// norm rounded to fp32 (the reference computes sqrt(dot(x,x)) in fp32 and divides by that fp32
// value, pico_vdb.py:60-68).  Pass 2: IEEE fp32 division, zero vector -> e0, write the fp32 row,
// the bf16 mirror row and the pad columns, then set the row's active bit.
__global__ void __launch_bounds__(256) upsert_normalize_scatter_kernel(
    const float* __restrict__ src, const int64_t* __restrict__ rows, int64_t row0, int64_t n, int dim,
    float* __restrict__ f32, int ld32, __nv_bfloat16* __restrict__ b16, int ld16,
    uint32_t* __restrict__ active, uint32_t* __restrict__ err_words) {
  const int lane = threadIdx.x & 31;
  const int64_t warp = (static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = (static_cast<int64_t>(gridDim.x) * blockDim.x) >> 5;
  const int ldmax = ld32 > ld16 ? ld32 : ld16;
  for (int64_t i = warp; i < n; i += nwarps) {
    const float* v = src + i * dim;
    float ss = 0.f;
    for (int c = lane; c < dim; c += 32) {
      float x = v[c];
      ss = fmaf(x, x, ss);
    }
    double d = static_cast<double>(ss);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) d += __shfl_xor_sync(0xffffffffu, d, o);
    const float nrm = static_cast<float>(sqrt(d));
    const bool zero = (nrm == 0.f);
    const int64_t row = rows ? rows[i] : row0 + i;
    float e_tf = 0.f, e_bf = 0.f;  // squared input-rounding error of this row (exactness guard)
    for (int c = lane; c < ldmax; c += 32) {
      float y = 0.f;
      if (c < dim) y = zero ? (c == 0 ? 1.f : 0.f) : __fdiv_rn(v[c], nrm);
      if (f32 != nullptr && c < ld32) f32[row * ld32 + c] = y;
      if (b16 != nullptr && c < ld16) b16[row * ld16 + c] = __float2bfloat16_rn(y);
      const float dt = tf32_trunc_err(y), db = bf16_rn_err(y);
      e_tf = fmaf(dt, dt, e_tf);
      e_bf = fmaf(db, db, e_bf);
    }
    if (err_words != nullptr) {
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        e_tf += __shfl_xor_sync(0xffffffffu, e_tf, o);
        e_bf += __shfl_xor_sync(0xffffffffu, e_bf, o);
      }
      if (lane == 0) {
        // non-negative floats order like their bit patterns
        if (__float_as_uint(e_tf) > err_words[0]) atomicMax(&err_words[0], __float_as_uint(e_tf));
        if (__float_as_uint(e_bf) > err_words[1]) atomicMax(&err_words[1], __float_as_uint(e_bf));
      }
    }
    if (lane == 0) atomicOr(&active[row >> 5], 1u << (row & 31));
  }
}

__global__ void __launch_bounds__(256) delete_rows_kernel(const int64_t* __restrict__ rows, int64_t n,
                                                          float* __restrict__ f32, int ld32,
                                                          __nv_bfloat16* __restrict__ b16, int ld16,
                                                          uint32_t* __restrict__ active) {
  const int lane = threadIdx.x & 31;
  const int64_t warp = (static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = (static_cast<int64_t>(gridDim.x) * blockDim.x) >> 5;
  for (int64_t i = warp; i < n; i += nwarps) {
    const int64_t row = rows[i];
    if (f32) for (int c = lane; c < ld32; c += 32) f32[row * ld32 + c] = 0.f;
    if (b16) for (int c = lane; c < ld16; c += 32) b16[row * ld16 + c] = __float2bfloat16_rn(0.f);
    if (lane == 0) atomicAnd(&active[row >> 5], ~(1u << (row & 31)));
  }
}

// out[i, :] = row rows[i] (or row0+i) as dense fp32; reads the fp32 matrix when present, else the
// bf16 mirror.
__global__ void __launch_bounds__(256) gather_rows_kernel(const int64_t* __restrict__ rows, int64_t row0,
                                                          int64_t n, int dim,
                                                          const float* __restrict__ f32, int ld32,
                                                          const __nv_bfloat16* __restrict__ b16, int ld16,
                                                          float* __restrict__ out) {
  const int lane = threadIdx.x & 31;
  const int64_t warp = (static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = (static_cast<int64_t>(gridDim.x) * blockDim.x) >> 5;
  for (int64_t i = warp; i < n; i += nwarps) {
    const int64_t row = rows ? rows[i] : row0 + i;
    for (int c = lane; c < dim; c += 32)
      out[i * dim + c] = f32 ? f32[row * ld32 + c] : __bfloat162float(b16[row * ld16 + c]);
  }
}

// Rows [row0, row0+n): build the bf16 mirror (and, for a bf16-only store, take the values from
// the dense staging buffer `src`).
__global__ void __launch_bounds__(256) mirror_rows_kernel(const float* __restrict__ src_dense, int dim,
                                                          const float* __restrict__ f32, int ld32,
                                                          __nv_bfloat16* __restrict__ b16, int ld16,
                                                          int64_t row0, int64_t n) {
  const int lane = threadIdx.x & 31;
  const int64_t warp = (static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = (static_cast<int64_t>(gridDim.x) * blockDim.x) >> 5;
  for (int64_t i = warp; i < n; i += nwarps) {
    const int64_t row = row0 + i;
    for (int c = lane; c < ld16; c += 32) {
      float y = 0.f;
      if (c < dim) y = src_dense ? src_dense[i * dim + c] : f32[row * ld32 + c];
      b16[row * ld16 + c] = __float2bfloat16_rn(y);
    }
  }
}

// Raw uploads bypass the upsert kernel: same rounding-error bookkeeping for rows [row0, row0+n).
__global__ void __launch_bounds__(256) row_error_kernel(const float* __restrict__ f32, int ld32, int dim,
                                                        int64_t row0, int64_t n, uint32_t* __restrict__ err_words) {
  const int lane = threadIdx.x & 31;
  const int64_t warp = (static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = (static_cast<int64_t>(gridDim.x) * blockDim.x) >> 5;
  for (int64_t i = warp; i < n; i += nwarps) {
    const float* v = f32 + (row0 + i) * ld32;
    float e_tf = 0.f, e_bf = 0.f;
    for (int c = lane; c < dim; c += 32) {
      const float y = v[c];
      const float dt = tf32_trunc_err(y), db = bf16_rn_err(y);
      e_tf = fmaf(dt, dt, e_tf);
      e_bf = fmaf(db, db, e_bf);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      e_tf += __shfl_xor_sync(0xffffffffu, e_tf, o);
      e_bf += __shfl_xor_sync(0xffffffffu, e_bf, o);
    }
    if (lane == 0) {
      if (__float_as_uint(e_tf) > err_words[0]) atomicMax(&err_words[0], __float_as_uint(e_tf));
      if (__float_as_uint(e_bf) > err_words[1]) atomicMax(&err_words[1], __float_as_uint(e_bf));
    }
  }
}

// active[w] for the words covering rows [row0, row0+n): take `bits` (word 0 == rows row0..row0+31)
// or all ones, restricted to the range; bits outside the range keep their value.
__global__ void set_active_range_kernel(uint32_t* __restrict__ active, int64_t row0, int64_t n,
                                        const uint32_t* __restrict__ bits) {
  const int64_t w0 = row0 >> 5;
  const int64_t w1 = (row0 + n + 31) >> 5;
  for (int64_t w = w0 + blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; w < w1;
       w += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int64_t lo = w << 5;
    uint32_t mask = 0xffffffffu;
    if (lo < row0) mask &= 0xffffffffu << (row0 - lo);
    if (lo + 32 > row0 + n) mask &= 0xffffffffu >> (lo + 32 - (row0 + n));
    const uint32_t val = bits ? bits[w - w0] : 0xffffffffu;
    active[w] = (active[w] & ~mask) | (val & mask);
  }
}

__global__ void popcount_kernel(const uint32_t* __restrict__ words, int64_t nwords,
                                unsigned long long* __restrict__ out) {
  unsigned long long c = 0;
  for (int64_t w = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; w < nwords;
       w += static_cast<int64_t>(gridDim.x) * blockDim.x)
    c += __popc(words[w]);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
  if ((threadIdx.x & 31) == 0 && c) atomicAdd(out, c);
}

// scratch row i := matrix row keep[i]; rows are `row_u4` 16-byte units long (both matrices keep
// 16-byte aligned row strides), so one kernel serves the fp32 matrix and the bf16 mirror
__global__ void __launch_bounds__(256) compact_gather_kernel(const int64_t* __restrict__ keep, int64_t n,
                                                             const uint4* __restrict__ src, int row_u4,
                                                             uint4* __restrict__ dst) {
  const int lane = threadIdx.x & 31;
  const int64_t warp = (static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = (static_cast<int64_t>(gridDim.x) * blockDim.x) >> 5;
  for (int64_t i = warp; i < n; i += nwarps) {
    const uint4* s = src + keep[i] * row_u4;
    for (int c = lane; c < row_u4; c += 32) dst[i * row_u4 + c] = s[c];
  }
}

static inline int warp_grid(int64_t n_items) {
  // 8 warps per block; enough blocks for one warp per item, capped at 16 blocks per SM
  int64_t blocks = (n_items + 7) / 8;
  return static_cast<int>(std::max<int64_t>(1, std::min<int64_t>(blocks, kNumSMs * 16)));
}

}  // namespace pvdb

using namespace pvdb;

// ---------------------------------------------------------------------------- store methods
int pvdb_store::acquire_io_slot() {
  std::unique_lock<std::mutex> lk(io_mu);
  for (;;) {
    for (int i = 0; i < kIoSlots; ++i)
      if (!io[i].busy) {
        io[i].busy = true;
        return i;
      }
    io_cv.wait(lk);
  }
}

void pvdb_store::release_io_slot(int slot) {
  {
    std::lock_guard<std::mutex> lk(io_mu);
    io[slot].busy = false;
  }
  io_cv.notify_one();
}

int pvdb_store::use_stream(cudaStream_t s) {
  // last_stream is always a live stream (it starts as the store's own); nullptr is CUDA's legacy
  // default stream, which is just another stream here.
  if (s != last_stream) {
    PVDB_CUDA(cudaEventRecord(order_event, last_stream));
    PVDB_CUDA(cudaStreamWaitEvent(s, order_event, 0));
    last_stream = s;
  }
  return PVDB_OK;
}

int pvdb_store::ensure_capacity(int64_t need_rows, cudaStream_t s) {
  if (need_rows <= capacity) return PVDB_OK;
  if ((flags & PVDB_STORE_FIXED_CAPACITY) && capacity > 0)
    return fail(PVDB_ERR_CAPACITY, "Database capacity exceeded (%lld > %lld rows)",
                static_cast<long long>(need_rows), static_cast<long long>(capacity));
  // geometric growth keeps appends amortised O(1); with VMM-backed buffers a growth step only maps
  // more memory (no copy), without it the step is a reallocate-and-copy
  int64_t cap = std::max<int64_t>(need_rows, capacity + capacity / 2);
  cap = (cap + 1023) & ~int64_t(1023);
  // + kTailSlack: the batch path's last K box of a row may extend past the row (see encode_map)
  if (flags & PVDB_STORE_F32) PVDB_TRY(f32.grow(static_cast<size_t>(cap) * ld_f32 * sizeof(float) + kTailSlack, s));
  if (flags & PVDB_STORE_BF16)
    PVDB_TRY(bf16.grow(static_cast<size_t>(cap) * ld_bf16 * sizeof(__nv_bfloat16) + kTailSlack, s));
  PVDB_TRY(active.grow(static_cast<size_t>(cap / 32) * sizeof(uint32_t), s));
  for (DeviceBuffer& col : column)
    if (col.ptr) PVDB_TRY(col.grow(static_cast<size_t>(cap) * sizeof(uint32_t), s));
  capacity = cap;
  return PVDB_OK;
}

#define PVDB_ENTER(s)                                                         \
  if ((s) == nullptr) return fail(PVDB_ERR_INVALID, "null store handle");     \
  std::lock_guard<std::mutex> _guard((s)->mu);                                \
  PVDB_CUDA(cudaSetDevice((s)->device))

// ---------------------------------------------------------------------------- C ABI: library
extern "C" int pvdb_abi_version(void) { return PVDB_ABI_VERSION; }
extern "C" const char* pvdb_last_error(void) { return g_last_error.c_str(); }
extern "C" int64_t pvdb_kernel_launches(void) { return g_launches.load(); }

extern "C" int pvdb_device_count(int* out_count) {
  if (!out_count) return fail(PVDB_ERR_INVALID, "out_count is null");
  int n = 0;
  PVDB_CUDA(cudaGetDeviceCount(&n));
  *out_count = n;
  return PVDB_OK;
}

// ---------------------------------------------------------------------------- C ABI: store
extern "C" int pvdb_store_create(pvdb_store_t** out, int device, int dim, int64_t reserve_rows, int flags) {
  if (!out) return fail(PVDB_ERR_INVALID, "out is null");
  *out = nullptr;
  if (dim <= 0 || dim > 65536) return fail(PVDB_ERR_INVALID, "dim %d out of range [1, 65536]", dim);
  if (reserve_rows < 0) return fail(PVDB_ERR_INVALID, "reserve_rows < 0");
  if ((flags & (PVDB_STORE_F32 | PVDB_STORE_BF16)) == 0) flags |= PVDB_STORE_F32;
  int ndev = 0;
  PVDB_CUDA(cudaGetDeviceCount(&ndev));
  if (device < 0 || device >= ndev) return fail(PVDB_ERR_INVALID, "device %d not in [0, %d)", device, ndev);
  PVDB_CUDA(cudaSetDevice(device));
  cudaDeviceProp prop;
  PVDB_CUDA(cudaGetDeviceProperties(&prop, device));
  if (prop.major != 10)
    return fail(PVDB_ERR_UNSUPPORTED, "device %d is sm_%d%d; this library is built for sm_100a only", device,
                prop.major, prop.minor);
  pvdb_store* s = new pvdb_store();
  s->device = device;
  s->dim = dim;
  s->ld_f32 = (dim + 3) & ~3;
  s->ld_bf16 = (dim + 7) & ~7;
  s->ldq = (dim + 63) & ~63;  // whole K boxes of the batch path's TMA loads (32 fp32 / 64 bf16), zero padded
  s->flags = flags;
  s->h_pinned.pinned_host = true;
  s->h_flag.pinned_host = true;
  s->h_pipe[0].pinned_host = s->h_pipe[1].pinned_host = true;
  cudaError_t e = cudaStreamCreateWithFlags(&s->stream, cudaStreamNonBlocking);
  for (int b = 0; b < 2 && e == cudaSuccess; ++b) e = cudaEventCreateWithFlags(&s->pipe_ev[b], cudaEventDisableTiming);
  for (pvdb_store::IoSlot& io : s->io) {
    io.h_res.pinned_host = true;
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&io.done, cudaEventDisableTiming);
  }
  if (e == cudaSuccess) e = cudaMalloc(reinterpret_cast<void**>(&s->d_err_words), 16);
  if (e == cudaSuccess) e = cudaMemset(s->d_err_words, 0, 16);
  if (e == cudaSuccess) e = cudaEventCreateWithFlags(&s->order_event, cudaEventDisableTiming);
  if (e != cudaSuccess) {
    delete s;
    return fail(PVDB_ERR_CUDA, "stream/event creation failed: %s", cudaGetErrorString(e));
  }
  s->last_stream = s->stream;
  int rc = PVDB_OK;
  if (reserve_rows > 0) {
    int keep = s->flags;
    s->flags &= ~PVDB_STORE_FIXED_CAPACITY;  // the first allocation is always allowed
    rc = s->ensure_capacity(reserve_rows, s->stream);
    s->flags = keep;
  }
  if (rc != PVDB_OK) {
    pvdb_store_destroy(s);
    return rc;
  }
  *out = s;
  return PVDB_OK;
}

extern "C" int pvdb_store_destroy(pvdb_store_t* s) {
  if (!s) return PVDB_OK;
  cudaSetDevice(s->device);
  cudaDeviceSynchronize();
  s->f32.release();
  s->bf16.release();
  s->active.release();
  s->drop_columns();
  for (Scratch* sc : {&s->d_in, &s->d_rows, &s->d_prefilter, &s->d_qn, &s->d_qn16, &s->d_partial, &s->d_out,
                      &s->d_misc, &s->h_pinned, &s->d_qeps, &s->d_flag, &s->h_flag, &s->d_xloc, &s->h_pipe[0],
                      &s->h_pipe[1]})
    sc->release();
  for (cudaEvent_t ev : s->pipe_ev)
    if (ev) cudaEventDestroy(ev);
  for (pvdb_store::IoSlot& io : s->io) {
    io.h_res.release();
    if (io.done) cudaEventDestroy(io.done);
  }
  if (s->d_err_words) cudaFree(s->d_err_words);
  if (s->order_event) cudaEventDestroy(s->order_event);
  if (s->stream) cudaStreamDestroy(s->stream);
  delete s;
  return PVDB_OK;
}

extern "C" int pvdb_store_reserve(pvdb_store_t* s, int64_t rows) {
  PVDB_ENTER(s);
  PVDB_TRY(s->use_stream(s->stream));
  int keep = s->flags;
  if (s->capacity == 0) s->flags &= ~PVDB_STORE_FIXED_CAPACITY;
  int rc = s->ensure_capacity(rows, s->stream);
  s->flags = keep;
  return rc;
}

extern "C" int pvdb_store_set_row_base(pvdb_store_t* s, int64_t row_base) {
  PVDB_ENTER(s);
  s->row_base = row_base;
  return PVDB_OK;
}

extern "C" int pvdb_store_info(pvdb_store_t* s, pvdb_store_info_t* out) {
  PVDB_ENTER(s);
  if (!out) return fail(PVDB_ERR_INVALID, "out is null");
  PVDB_TRY(s->use_stream(s->stream));
  unsigned long long act = 0;
  if (s->rows > 0) {
    PVDB_TRY(s->d_misc.ensure(sizeof(unsigned long long)));
    PVDB_CUDA(cudaMemsetAsync(s->d_misc.ptr, 0, sizeof(unsigned long long), s->stream));
    const int64_t nwords = (s->rows + 31) >> 5;
    const int blocks = static_cast<int>(std::min<int64_t>((nwords + 255) / 256, kNumSMs * 8));
    popcount_kernel<<<blocks, 256, 0, s->stream>>>(static_cast<const uint32_t*>(s->active.ptr), nwords,
                                                  static_cast<unsigned long long*>(s->d_misc.ptr));
    PVDB_LAUNCH_CHECK();
    PVDB_CUDA(cudaMemcpyAsync(&act, s->d_misc.ptr, sizeof(act), cudaMemcpyDeviceToHost, s->stream));
    PVDB_CUDA(cudaStreamSynchronize(s->stream));
  }
  out->dim = s->dim;
  out->ld_f32 = s->ld_f32;
  out->ld_bf16 = s->ld_bf16;
  out->flags = s->flags;
  out->device = s->device;
  out->reserved = 0;
  out->rows = s->rows;
  out->capacity = s->capacity;
  out->active = static_cast<int64_t>(act);
  out->row_base = s->row_base;
  out->device_bytes = s->f32.bytes + s->bf16.bytes + s->active.bytes;
  return PVDB_OK;
}

// ---------------------------------------------------------------------------- host <-> device streaming
static constexpr int64_t kPipeBytes = 32ll << 20;  // per pinned buffer

// memcpy split over a few threads: one core copies ~10 GB/s, PCIe 5 moves ~50 GB/s
static void parallel_memcpy(void* dst, const void* src, size_t bytes) {
  unsigned hw = std::thread::hardware_concurrency();
  const unsigned want = std::min<unsigned>(8, std::max<unsigned>(1, hw / 2));
  const size_t min_piece = size_t(2) << 20;
  const unsigned n = static_cast<unsigned>(std::min<size_t>(want, std::max<size_t>(1, bytes / min_piece)));
  if (n <= 1) {
    std::memcpy(dst, src, bytes);
    return;
  }
  std::vector<std::thread> th;
  const size_t piece = ((bytes + n - 1) / n + 63) & ~size_t(63);
  for (unsigned i = 1; i < n; ++i) {
    const size_t off = std::min(bytes, piece * i), len = std::min(bytes - off, piece);
    if (len) th.emplace_back([=]() { std::memcpy(static_cast<char*>(dst) + off, static_cast<const char*>(src) + off, len); });
  }
  std::memcpy(dst, src, std::min(bytes, piece));
  for (auto& t : th) t.join();
}

// Host -> device, block by block through the pinned pair.  consume(d_block, first_item, items) enqueues
// the work that reads the staged block on `st` (it runs after the block's DMA, stream order).
template <typename F>
static int stream_h2d(pvdb_store* s, const void* host, int64_t items, size_t item_bytes, int64_t align_items,
                      cudaStream_t st, F&& consume) {
  int64_t per = std::max<int64_t>(1, kPipeBytes / static_cast<int64_t>(item_bytes));
  if (align_items > 1) per = std::max<int64_t>(align_items, per / align_items * align_items);
  const size_t block_bytes = static_cast<size_t>(std::min(per, items)) * item_bytes;
  PVDB_TRY(s->h_pipe[0].ensure(block_bytes));
  if (items > per) PVDB_TRY(s->h_pipe[1].ensure(block_bytes));
  PVDB_TRY(s->d_in.ensure(block_bytes));
  int b = 0;
  for (int64_t i0 = 0; i0 < items; i0 += per, b ^= 1) {
    const int64_t m = std::min(per, items - i0);
    const size_t bytes = static_cast<size_t>(m) * item_bytes;
    PVDB_CUDA(cudaEventSynchronize(s->pipe_ev[b]));  // the DMA that last read this pinned buffer is done
    parallel_memcpy(s->h_pipe[b].ptr, static_cast<const char*>(host) + static_cast<size_t>(i0) * item_bytes, bytes);
    PVDB_CUDA(cudaMemcpyAsync(s->d_in.ptr, s->h_pipe[b].ptr, bytes, cudaMemcpyHostToDevice, st));
    PVDB_CUDA(cudaEventRecord(s->pipe_ev[b], st));
    PVDB_TRY(consume(s->d_in.ptr, i0, m));
  }
  return PVDB_OK;
}

// Device -> host: produce(d_block, first_item, items) enqueues the work that fills the device staging
// block; its DMA into one pinned buffer overlaps the host copy of the previous block out of the other.
template <typename F>
static int stream_d2h(pvdb_store* s, void* host, int64_t items, size_t item_bytes, cudaStream_t st, F&& produce) {
  const int64_t per = std::max<int64_t>(1, kPipeBytes / static_cast<int64_t>(item_bytes));
  const size_t block_bytes = static_cast<size_t>(std::min(per, items)) * item_bytes;
  PVDB_TRY(s->h_pipe[0].ensure(block_bytes));
  if (items > per) PVDB_TRY(s->h_pipe[1].ensure(block_bytes));
  int b = 0;
  int64_t prev0 = -1, prev_m = 0;
  for (int64_t i0 = 0; i0 < items; i0 += per, b ^= 1) {
    const int64_t m = std::min(per, items - i0);
    PVDB_TRY(produce(s->h_pipe[b].ptr, i0, m));      // enqueues the D2H into pinned buffer b
    PVDB_CUDA(cudaEventRecord(s->pipe_ev[b], st));
    if (prev0 >= 0) {
      PVDB_CUDA(cudaEventSynchronize(s->pipe_ev[b ^ 1]));
      parallel_memcpy(static_cast<char*>(host) + static_cast<size_t>(prev0) * item_bytes, s->h_pipe[b ^ 1].ptr,
                      static_cast<size_t>(prev_m) * item_bytes);
    }
    prev0 = i0;
    prev_m = m;
  }
  if (prev0 >= 0) {
    PVDB_CUDA(cudaEventSynchronize(s->pipe_ev[b ^ 1]));
    parallel_memcpy(static_cast<char*>(host) + static_cast<size_t>(prev0) * item_bytes, s->h_pipe[b ^ 1].ptr,
                    static_cast<size_t>(prev_m) * item_bytes);
  }
  return PVDB_OK;
}

// shared tail of the four upsert entry points: data already on the device
static int upsert_device(pvdb_store* s, const float* d_vecs, const int64_t* d_rows, int64_t row0, int64_t n,
                         int64_t max_row, cudaStream_t st) {
  PVDB_TRY(s->ensure_capacity(max_row + 1, st));
  upsert_normalize_scatter_kernel<<<warp_grid(n), 256, 0, st>>>(
      d_vecs, d_rows, row0, n, s->dim, static_cast<float*>(s->f32.ptr), s->ld_f32,
      static_cast<__nv_bfloat16*>(s->bf16.ptr), s->ld_bf16, static_cast<uint32_t*>(s->active.ptr),
      s->f32.ptr ? s->d_err_words : nullptr);
  PVDB_LAUNCH_CHECK();
  s->rows = std::max(s->rows, max_row + 1);
  return PVDB_OK;
}

static constexpr int64_t kStageBytes = 64ll << 20;  // host data is staged through HBM in 64 MiB pieces

extern "C" int pvdb_store_upsert(pvdb_store_t* s, const float* vecs, const int64_t* rows, int64_t n) {
  PVDB_ENTER(s);
  if (n == 0) return PVDB_OK;
  if (!vecs || !rows || n < 0) return fail(PVDB_ERR_INVALID, "upsert: null buffer or negative count");
  int64_t max_row = -1;
  for (int64_t i = 0; i < n; ++i) {
    if (rows[i] < 0 || rows[i] > 0xfffffffell) return fail(PVDB_ERR_INVALID, "upsert: row %lld out of range", (long long)rows[i]);
    max_row = std::max(max_row, rows[i]);
  }
  cudaStream_t st = s->stream;
  PVDB_TRY(s->use_stream(st));
  PVDB_TRY(s->ensure_capacity(max_row + 1, st));
  const int64_t chunk = std::max<int64_t>(1, kStageBytes / (static_cast<int64_t>(s->dim) * 4));
  PVDB_TRY(s->d_in.ensure(static_cast<size_t>(std::min(chunk, n)) * s->dim * sizeof(float)));
  PVDB_TRY(s->d_rows.ensure(static_cast<size_t>(std::min(chunk, n)) * sizeof(int64_t)));
  for (int64_t i0 = 0; i0 < n; i0 += chunk) {
    const int64_t m = std::min(chunk, n - i0);
    PVDB_CUDA(cudaMemcpyAsync(s->d_in.ptr, vecs + i0 * s->dim, static_cast<size_t>(m) * s->dim * sizeof(float),
                              cudaMemcpyHostToDevice, st));
    PVDB_CUDA(cudaMemcpyAsync(s->d_rows.ptr, rows + i0, static_cast<size_t>(m) * sizeof(int64_t),
                              cudaMemcpyHostToDevice, st));
    PVDB_TRY(upsert_device(s, static_cast<const float*>(s->d_in.ptr), static_cast<const int64_t*>(s->d_rows.ptr),
                           0, m, max_row, st));
  }
  PVDB_CUDA(cudaStreamSynchronize(st));
  return PVDB_OK;
}

extern "C" int pvdb_store_upsert_range(pvdb_store_t* s, const float* vecs, int64_t row0, int64_t n) {
  PVDB_ENTER(s);
  if (n == 0) return PVDB_OK;
  if (!vecs || n < 0 || row0 < 0 || row0 + n - 1 > 0xfffffffell)
    return fail(PVDB_ERR_INVALID, "upsert_range: bad arguments");
  cudaStream_t st = s->stream;
  PVDB_TRY(s->use_stream(st));
  PVDB_TRY(s->ensure_capacity(row0 + n, st));
  // pinned double buffering: the host copy of block i+1 overlaps the DMA + kernel of block i
  PVDB_TRY(stream_h2d(s, vecs, n, static_cast<size_t>(s->dim) * sizeof(float), 1, st,
                      [&](void* d_block, int64_t i0, int64_t m) {
                        return upsert_device(s, static_cast<const float*>(d_block), nullptr, row0 + i0, m,
                                             row0 + i0 + m - 1, st);
                      }));
  PVDB_CUDA(cudaStreamSynchronize(st));
  return PVDB_OK;
}

extern "C" int pvdb_store_upsert_dev(pvdb_store_t* s, const float* d_vecs, const int64_t* d_rows, int64_t n,
                                     int64_t max_row, void* stream) {
  PVDB_ENTER(s);
  if (n == 0) return PVDB_OK;
  if (!d_vecs || !d_rows || n < 0 || max_row < 0 || max_row > 0xfffffffell)
    return fail(PVDB_ERR_INVALID, "upsert_dev: bad arguments");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  PVDB_TRY(s->use_stream(st));
  return upsert_device(s, d_vecs, d_rows, 0, n, max_row, st);
}

extern "C" int pvdb_store_upsert_range_dev(pvdb_store_t* s, const float* d_vecs, int64_t row0, int64_t n,
                                           void* stream) {
  PVDB_ENTER(s);
  if (n == 0) return PVDB_OK;
  if (!d_vecs || n < 0 || row0 < 0 || row0 + n - 1 > 0xfffffffell)
    return fail(PVDB_ERR_INVALID, "upsert_range_dev: bad arguments");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  PVDB_TRY(s->use_stream(st));
  return upsert_device(s, d_vecs, nullptr, row0, n, row0 + n - 1, st);
}

extern "C" int pvdb_store_delete(pvdb_store_t* s, const int64_t* rows, int64_t n) {
  PVDB_ENTER(s);
  if (n == 0) return PVDB_OK;
  if (!rows || n < 0) return fail(PVDB_ERR_INVALID, "delete: null rows or negative count");
  for (int64_t i = 0; i < n; ++i)
    if (rows[i] < 0 || rows[i] >= s->rows)
      return fail(PVDB_ERR_INVALID, "delete: row %lld outside [0, %lld)", (long long)rows[i], (long long)s->rows);
  cudaStream_t st = s->stream;
  PVDB_TRY(s->use_stream(st));
  PVDB_TRY(s->d_rows.ensure(static_cast<size_t>(n) * sizeof(int64_t)));
  PVDB_CUDA(cudaMemcpyAsync(s->d_rows.ptr, rows, static_cast<size_t>(n) * sizeof(int64_t), cudaMemcpyHostToDevice, st));
  delete_rows_kernel<<<warp_grid(n), 256, 0, st>>>(static_cast<const int64_t*>(s->d_rows.ptr), n,
                                                   static_cast<float*>(s->f32.ptr), s->ld_f32,
                                                   static_cast<__nv_bfloat16*>(s->bf16.ptr), s->ld_bf16,
                                                   static_cast<uint32_t*>(s->active.ptr));
  PVDB_LAUNCH_CHECK();
  PVDB_CUDA(cudaStreamSynchronize(st));
  return PVDB_OK;
}

extern "C" int pvdb_store_fetch(pvdb_store_t* s, const int64_t* rows, int64_t n, float* out) {
  PVDB_ENTER(s);
  if (n == 0) return PVDB_OK;
  if (!rows || !out || n < 0) return fail(PVDB_ERR_INVALID, "fetch: null buffer or negative count");
  for (int64_t i = 0; i < n; ++i)
    if (rows[i] < 0 || rows[i] >= s->rows)
      return fail(PVDB_ERR_INVALID, "fetch: row %lld outside [0, %lld)", (long long)rows[i], (long long)s->rows);
  cudaStream_t st = s->stream;
  PVDB_TRY(s->use_stream(st));
  const int64_t chunk = std::max<int64_t>(1, kStageBytes / (static_cast<int64_t>(s->dim) * 4));
  PVDB_TRY(s->d_in.ensure(static_cast<size_t>(std::min(chunk, n)) * s->dim * sizeof(float)));
  PVDB_TRY(s->d_rows.ensure(static_cast<size_t>(std::min(chunk, n)) * sizeof(int64_t)));
  for (int64_t i0 = 0; i0 < n; i0 += chunk) {
    const int64_t m = std::min(chunk, n - i0);
    PVDB_CUDA(cudaMemcpyAsync(s->d_rows.ptr, rows + i0, static_cast<size_t>(m) * sizeof(int64_t), cudaMemcpyHostToDevice, st));
    gather_rows_kernel<<<warp_grid(m), 256, 0, st>>>(static_cast<const int64_t*>(s->d_rows.ptr), 0, m, s->dim,
                                                     static_cast<const float*>(s->f32.ptr), s->ld_f32,
                                                     static_cast<const __nv_bfloat16*>(s->bf16.ptr), s->ld_bf16,
                                                     static_cast<float*>(s->d_in.ptr));
    PVDB_LAUNCH_CHECK();
    PVDB_CUDA(cudaMemcpyAsync(out + i0 * s->dim, s->d_in.ptr, static_cast<size_t>(m) * s->dim * sizeof(float),
                              cudaMemcpyDeviceToHost, st));
  }
  PVDB_CUDA(cudaStreamSynchronize(st));
  return PVDB_OK;
}

extern "C" int pvdb_store_download(pvdb_store_t* s, int64_t row0, int64_t n, float* out) {
  PVDB_ENTER(s);
  if (n == 0) return PVDB_OK;
  if (!out || n < 0 || row0 < 0 || row0 + n > s->rows)
    return fail(PVDB_ERR_INVALID, "download: range [%lld, %lld) outside [0, %lld)", (long long)row0,
                (long long)(row0 + n), (long long)s->rows);
  cudaStream_t st = s->stream;
  PVDB_TRY(s->use_stream(st));
  // streamed through the pinned pair: the DMA of block i+1 overlaps the host copy of block i into the
  // caller's buffer (a memory-mapped .npy file in save(), pico_vdb.py:356)
  const size_t row_bytes = static_cast<size_t>(s->dim) * sizeof(float);
  if (!s->f32.ptr) PVDB_TRY(s->d_in.ensure(static_cast<size_t>(std::min<int64_t>(n, kPipeBytes / row_bytes + 1)) * row_bytes));
  PVDB_TRY(stream_d2h(s, out, n, row_bytes, st, [&](void* pinned, int64_t i0, int64_t m) -> int {
    if (s->f32.ptr) {
      const float* src = static_cast<const float*>(s->f32.ptr) + (row0 + i0) * s->ld_f32;
      PVDB_CUDA(cudaMemcpy2DAsync(pinned, row_bytes, src, static_cast<size_t>(s->ld_f32) * 4, row_bytes,
                                  static_cast<size_t>(m), cudaMemcpyDeviceToHost, st));
    } else {
      gather_rows_kernel<<<warp_grid(m), 256, 0, st>>>(nullptr, row0 + i0, m, s->dim, nullptr, s->ld_f32,
                                                       static_cast<const __nv_bfloat16*>(s->bf16.ptr), s->ld_bf16,
                                                       static_cast<float*>(s->d_in.ptr));
      PVDB_LAUNCH_CHECK();
      PVDB_CUDA(cudaMemcpyAsync(pinned, s->d_in.ptr, static_cast<size_t>(m) * row_bytes, cudaMemcpyDeviceToHost, st));
    }
    return PVDB_OK;
  }));
  return PVDB_OK;
}

// The bf16 mirror as it is: n x dim bf16 (uint16 bit patterns), for persisting a bf16-only store at half
// the size of its fp32 expansion.
extern "C" int pvdb_store_download_bf16(pvdb_store_t* s, int64_t row0, int64_t n, uint16_t* out) {
  PVDB_ENTER(s);
  if (n == 0) return PVDB_OK;
  if (!s->bf16.ptr) return fail(PVDB_ERR_UNSUPPORTED, "download_bf16: this store keeps no bf16 mirror");
  if (!out || n < 0 || row0 < 0 || row0 + n > s->rows)
    return fail(PVDB_ERR_INVALID, "download_bf16: range [%lld, %lld) outside [0, %lld)", (long long)row0,
                (long long)(row0 + n), (long long)s->rows);
  cudaStream_t st = s->stream;
  PVDB_TRY(s->use_stream(st));
  const size_t row_bytes = static_cast<size_t>(s->dim) * sizeof(uint16_t);
  PVDB_TRY(stream_d2h(s, out, n, row_bytes, st, [&](void* pinned, int64_t i0, int64_t m) -> int {
    const __nv_bfloat16* src = static_cast<const __nv_bfloat16*>(s->bf16.ptr) + (row0 + i0) * s->ld_bf16;
    PVDB_CUDA(cudaMemcpy2DAsync(pinned, row_bytes, src, static_cast<size_t>(s->ld_bf16) * 2, row_bytes,
                                static_cast<size_t>(m), cudaMemcpyDeviceToHost, st));
    return PVDB_OK;
  }));
  return PVDB_OK;
}

extern "C" int pvdb_store_upload(pvdb_store_t* s, int64_t row0, int64_t n, const float* vecs,
                                 const uint32_t* active_bits) {
  PVDB_ENTER(s);
  if (n == 0) return PVDB_OK;
  if (!vecs || n < 0 || row0 < 0 || row0 + n - 1 > 0xfffffffell)
    return fail(PVDB_ERR_INVALID, "upload: bad arguments");
  if (active_bits && (row0 & 31)) return fail(PVDB_ERR_INVALID, "upload: row0 must be a multiple of 32 with active_bits");
  cudaStream_t st = s->stream;
  PVDB_TRY(s->use_stream(st));
  PVDB_TRY(s->ensure_capacity(row0 + n, st));
  const size_t row_bytes = static_cast<size_t>(s->dim) * sizeof(float);
  PVDB_TRY(stream_h2d(s, vecs, n, row_bytes, 32, st, [&](void* d_block, int64_t i0, int64_t m) -> int {
    const float* dense = static_cast<const float*>(d_block);
    if (s->f32.ptr) {
      float* dst = static_cast<float*>(s->f32.ptr) + (row0 + i0) * s->ld_f32;
      PVDB_CUDA(cudaMemcpy2DAsync(dst, static_cast<size_t>(s->ld_f32) * 4, dense, row_bytes, row_bytes,
                                  static_cast<size_t>(m), cudaMemcpyDeviceToDevice, st));
    }
    if (s->bf16.ptr) {
      mirror_rows_kernel<<<warp_grid(m), 256, 0, st>>>(dense, s->dim, static_cast<const float*>(s->f32.ptr), s->ld_f32,
                                                       static_cast<__nv_bfloat16*>(s->bf16.ptr), s->ld_bf16, row0 + i0, m);
      PVDB_LAUNCH_CHECK();
    }
    if (s->f32.ptr) {
      row_error_kernel<<<warp_grid(m), 256, 0, st>>>(static_cast<const float*>(s->f32.ptr), s->ld_f32, s->dim,
                                                     row0 + i0, m, s->d_err_words);
      PVDB_LAUNCH_CHECK();
    }
    return PVDB_OK;
  }));
  const uint32_t* d_bits = nullptr;
  if (active_bits) {
    const size_t nwords = static_cast<size_t>((n + 31) >> 5);
    PVDB_TRY(s->d_prefilter.ensure(nwords * sizeof(uint32_t)));
    PVDB_CUDA(cudaMemcpyAsync(s->d_prefilter.ptr, active_bits, nwords * sizeof(uint32_t), cudaMemcpyHostToDevice, st));
    d_bits = static_cast<const uint32_t*>(s->d_prefilter.ptr);
  }
  {
    const int64_t nwords = ((row0 + n + 31) >> 5) - (row0 >> 5);
    const int blocks = static_cast<int>(std::min<int64_t>((nwords + 255) / 256, kNumSMs * 8));
    set_active_range_kernel<<<blocks, 256, 0, st>>>(static_cast<uint32_t*>(s->active.ptr), row0, n, d_bits);
    PVDB_LAUNCH_CHECK();
  }
  s->rows = std::max(s->rows, row0 + n);
  PVDB_CUDA(cudaStreamSynchronize(st));
  return PVDB_OK;
}

// pwrite of one block, split over a few threads (the copy into the page cache is the slow side)
static int parallel_pwrite(int fd, const void* src, size_t bytes, int64_t offset) {
  unsigned hw = std::thread::hardware_concurrency();
  const unsigned want = std::min<unsigned>(4, std::max<unsigned>(1, hw / 2));
  const size_t min_piece = size_t(4) << 20;
  const unsigned n = static_cast<unsigned>(std::min<size_t>(want, std::max<size_t>(1, bytes / min_piece)));
  const size_t piece = ((bytes + n - 1) / n + 4095) & ~size_t(4095);
  std::vector<int> err(n, 0);
  auto work = [&](unsigned i) {
    size_t off = std::min(bytes, piece * i);
    size_t len = std::min(bytes - off, piece);
    while (len > 0) {
      const ssize_t w = pwrite(fd, static_cast<const char*>(src) + off, len, offset + static_cast<int64_t>(off));
      if (w < 0) {
        if (errno == EINTR) continue;
        err[i] = errno;
        return;
      }
      off += static_cast<size_t>(w);
      len -= static_cast<size_t>(w);
    }
  };
  std::vector<std::thread> th;
  for (unsigned i = 1; i < n; ++i) th.emplace_back(work, i);
  work(0);
  for (auto& t : th) t.join();
  for (int e : err)
    if (e) return fail(PVDB_ERR_INVALID, "write to the vector file failed: %s", strerror(e));
  return PVDB_OK;
}

// Rows [row0, row0 + n) straight into an existing file at file_offset (dense fp32 rows, or the bf16
// mirror's bit patterns): the save() of a large store.  The device->host DMA of block i+1 overlaps the
// pwrite of block i out of the other pinned buffer; nothing is staged in pageable memory and no page of
// a mapped file is faulted in (mapping a fresh 6 GB .npy and storing into it ran at 1.9 GB/s).
extern "C" int pvdb_store_write_file(pvdb_store_t* s, const char* path, int64_t file_offset, int64_t row0, int64_t n,
                                     int as_bf16) {
  PVDB_ENTER(s);
  if (n == 0) return PVDB_OK;
  if (!path || n < 0 || row0 < 0 || row0 + n > s->rows || file_offset < 0)
    return fail(PVDB_ERR_INVALID, "write_file: bad arguments (rows [%lld, %lld) of %lld)", (long long)row0,
                (long long)(row0 + n), (long long)s->rows);
  if (as_bf16 && !s->bf16.ptr) return fail(PVDB_ERR_UNSUPPORTED, "write_file: this store keeps no bf16 mirror");
  const int fd = open(path, O_WRONLY);
  if (fd < 0) return fail(PVDB_ERR_INVALID, "write_file: cannot open %s: %s", path, strerror(errno));
  cudaStream_t st = s->stream;
  int rc = s->use_stream(st);
  const size_t row_bytes = static_cast<size_t>(s->dim) * (as_bf16 ? 2 : 4);
  const int64_t per = std::max<int64_t>(1, kPipeBytes / static_cast<int64_t>(row_bytes));
  if (rc == PVDB_OK) rc = s->h_pipe[0].ensure(static_cast<size_t>(std::min(per, n)) * row_bytes);
  if (rc == PVDB_OK && n > per) rc = s->h_pipe[1].ensure(static_cast<size_t>(std::min(per, n)) * row_bytes);
  if (rc == PVDB_OK && !as_bf16 && !s->f32.ptr) rc = s->d_in.ensure(static_cast<size_t>(std::min(per, n)) * row_bytes);
  auto enqueue = [&](void* pinned, int64_t i0, int64_t m) -> int {
    if (as_bf16) {
      const __nv_bfloat16* src = static_cast<const __nv_bfloat16*>(s->bf16.ptr) + (row0 + i0) * s->ld_bf16;
      PVDB_CUDA(cudaMemcpy2DAsync(pinned, row_bytes, src, static_cast<size_t>(s->ld_bf16) * 2, row_bytes,
                                  static_cast<size_t>(m), cudaMemcpyDeviceToHost, st));
    } else if (s->f32.ptr) {
      const float* src = static_cast<const float*>(s->f32.ptr) + (row0 + i0) * s->ld_f32;
      PVDB_CUDA(cudaMemcpy2DAsync(pinned, row_bytes, src, static_cast<size_t>(s->ld_f32) * 4, row_bytes,
                                  static_cast<size_t>(m), cudaMemcpyDeviceToHost, st));
    } else {
      gather_rows_kernel<<<warp_grid(m), 256, 0, st>>>(nullptr, row0 + i0, m, s->dim, nullptr, s->ld_f32,
                                                       static_cast<const __nv_bfloat16*>(s->bf16.ptr), s->ld_bf16,
                                                       static_cast<float*>(s->d_in.ptr));
      PVDB_LAUNCH_CHECK();
      PVDB_CUDA(cudaMemcpyAsync(pinned, s->d_in.ptr, static_cast<size_t>(m) * row_bytes, cudaMemcpyDeviceToHost, st));
    }
    return PVDB_OK;
  };
  int b = 0;
  int64_t prev0 = -1, prev_m = 0;
  for (int64_t i0 = 0; i0 < n && rc == PVDB_OK; i0 += per, b ^= 1) {
    const int64_t m = std::min(per, n - i0);
    rc = enqueue(s->h_pipe[b].ptr, i0, m);
    if (rc == PVDB_OK && cudaEventRecord(s->pipe_ev[b], st) != cudaSuccess) rc = fail(PVDB_ERR_CUDA, "cudaEventRecord failed");
    if (rc == PVDB_OK && prev0 >= 0) {
      if (cudaEventSynchronize(s->pipe_ev[b ^ 1]) != cudaSuccess) rc = fail(PVDB_ERR_CUDA, "cudaEventSynchronize failed");
      if (rc == PVDB_OK)
        rc = parallel_pwrite(fd, s->h_pipe[b ^ 1].ptr, static_cast<size_t>(prev_m) * row_bytes,
                             file_offset + prev0 * static_cast<int64_t>(row_bytes));
    }
    prev0 = i0;
    prev_m = m;
  }
  if (rc == PVDB_OK && prev0 >= 0) {
    if (cudaEventSynchronize(s->pipe_ev[b ^ 1]) != cudaSuccess) rc = fail(PVDB_ERR_CUDA, "cudaEventSynchronize failed");
    if (rc == PVDB_OK)
      rc = parallel_pwrite(fd, s->h_pipe[b ^ 1].ptr, static_cast<size_t>(prev_m) * row_bytes,
                           file_offset + prev0 * static_cast<int64_t>(row_bytes));
  }
  cudaStreamSynchronize(st);
  close(fd);
  return rc;
}

// Raw load of bf16 rows (what pvdb_store_download_bf16 wrote) into a bf16-only store.
extern "C" int pvdb_store_upload_bf16(pvdb_store_t* s, int64_t row0, int64_t n, const uint16_t* vecs,
                                      const uint32_t* active_bits) {
  PVDB_ENTER(s);
  if (n == 0) return PVDB_OK;
  if (!s->bf16.ptr && s->capacity > 0) return fail(PVDB_ERR_UNSUPPORTED, "upload_bf16: this store keeps no bf16 mirror");
  if (s->flags & PVDB_STORE_F32)
    return fail(PVDB_ERR_UNSUPPORTED, "upload_bf16: the store keeps an fp32 matrix; load fp32 rows instead");
  if (!vecs || n < 0 || row0 < 0 || row0 + n - 1 > 0xfffffffell) return fail(PVDB_ERR_INVALID, "upload_bf16: bad arguments");
  if (active_bits && (row0 & 31)) return fail(PVDB_ERR_INVALID, "upload_bf16: row0 must be a multiple of 32 with active_bits");
  cudaStream_t st = s->stream;
  PVDB_TRY(s->use_stream(st));
  PVDB_TRY(s->ensure_capacity(row0 + n, st));
  const size_t row_bytes = static_cast<size_t>(s->dim) * sizeof(uint16_t);
  PVDB_TRY(stream_h2d(s, vecs, n, row_bytes, 32, st, [&](void* d_block, int64_t i0, int64_t m) -> int {
    __nv_bfloat16* dst = static_cast<__nv_bfloat16*>(s->bf16.ptr) + (row0 + i0) * s->ld_bf16;
    PVDB_CUDA(cudaMemcpy2DAsync(dst, static_cast<size_t>(s->ld_bf16) * 2, d_block, row_bytes, row_bytes,
                                static_cast<size_t>(m), cudaMemcpyDeviceToDevice, st));
    return PVDB_OK;
  }));
  const uint32_t* d_bits = nullptr;
  if (active_bits) {
    const size_t nwords = static_cast<size_t>((n + 31) >> 5);
    PVDB_TRY(s->d_prefilter.ensure(nwords * sizeof(uint32_t)));
    PVDB_CUDA(cudaMemcpyAsync(s->d_prefilter.ptr, active_bits, nwords * sizeof(uint32_t), cudaMemcpyHostToDevice, st));
    d_bits = static_cast<const uint32_t*>(s->d_prefilter.ptr);
  }
  const int64_t nwords = ((row0 + n + 31) >> 5) - (row0 >> 5);
  const int blocks = static_cast<int>(std::min<int64_t>((nwords + 255) / 256, kNumSMs * 8));
  set_active_range_kernel<<<blocks, 256, 0, st>>>(static_cast<uint32_t*>(s->active.ptr), row0, n, d_bits);
  PVDB_LAUNCH_CHECK();
  s->rows = std::max(s->rows, row0 + n);
  PVDB_CUDA(cudaStreamSynchronize(st));
  return PVDB_OK;
}

extern "C" int pvdb_store_active_bits(pvdb_store_t* s, uint32_t* out_words) {
  PVDB_ENTER(s);
  if (s->rows == 0) return PVDB_OK;
  if (!out_words) return fail(PVDB_ERR_INVALID, "out_words is null");
  PVDB_TRY(s->use_stream(s->stream));
  PVDB_CUDA(cudaMemcpyAsync(out_words, s->active.ptr, static_cast<size_t>((s->rows + 31) >> 5) * sizeof(uint32_t),
                            cudaMemcpyDeviceToHost, s->stream));
  PVDB_CUDA(cudaStreamSynchronize(s->stream));
  return PVDB_OK;
}

extern "C" int pvdb_store_compact(pvdb_store_t* s, const int64_t* keep_rows, int64_t n) {
  PVDB_ENTER(s);
  if (n < 0 || (n > 0 && !keep_rows)) return fail(PVDB_ERR_INVALID, "compact: bad arguments");
  for (int64_t i = 0; i < n; ++i) {
    if (keep_rows[i] < 0 || keep_rows[i] >= s->rows || (i > 0 && keep_rows[i] <= keep_rows[i - 1]))
      return fail(PVDB_ERR_INVALID, "compact: keep_rows must be strictly ascending rows of the store");
  }
  cudaStream_t st = s->stream;
  PVDB_TRY(s->use_stream(st));
  // In place, block by block: keep_rows is strictly ascending, so keep[i] >= i -- a block of new rows
  // [a, b) only reads old rows >= a, and once it has been staged in scratch and written back, later
  // blocks read old rows >= b, which no earlier write touched.  No second copy of the matrices (the
  // first version allocated one: 2x the store's HBM, impossible for a C5-sized shard).
  if (n > 0) {
    PVDB_TRY(s->d_rows.ensure(static_cast<size_t>(n) * sizeof(int64_t)));
    PVDB_CUDA(cudaMemcpyAsync(s->d_rows.ptr, keep_rows, static_cast<size_t>(n) * sizeof(int64_t), cudaMemcpyHostToDevice, st));
    const int64_t* d_keep = static_cast<const int64_t*>(s->d_rows.ptr);
    struct Mat { void* ptr; size_t row_bytes; };
    const Mat mats[2] = {{s->f32.ptr, static_cast<size_t>(s->ld_f32) * sizeof(float)},
                         {s->bf16.ptr, static_cast<size_t>(s->ld_bf16) * sizeof(__nv_bfloat16)}};
    for (const Mat& m : mats) {
      if (m.ptr == nullptr) continue;
      const int64_t chunk = std::max<int64_t>(1, kStageBytes / static_cast<int64_t>(m.row_bytes));
      PVDB_TRY(s->d_in.ensure(static_cast<size_t>(std::min(chunk, n)) * m.row_bytes));
      // leading rows that stay where they are need no copy at all
      int64_t first = 0;
      while (first < n && keep_rows[first] == first) ++first;
      for (int64_t i0 = first; i0 < n; i0 += chunk) {
        const int64_t cnt = std::min(chunk, n - i0);
        compact_gather_kernel<<<warp_grid(cnt), 256, 0, st>>>(d_keep + i0, cnt, static_cast<const uint4*>(m.ptr),
                                                              static_cast<int>(m.row_bytes / 16),
                                                              static_cast<uint4*>(s->d_in.ptr));
        PVDB_LAUNCH_CHECK();
        PVDB_CUDA(cudaMemcpyAsync(static_cast<unsigned char*>(m.ptr) + static_cast<size_t>(i0) * m.row_bytes,
                                  s->d_in.ptr, static_cast<size_t>(cnt) * m.row_bytes, cudaMemcpyDeviceToDevice, st));
      }
      // rows past the new end read as zeros again (deleted-row convention, pico_vdb.py:523)
      if (s->rows > n)
        PVDB_CUDA(cudaMemsetAsync(static_cast<unsigned char*>(m.ptr) + static_cast<size_t>(n) * m.row_bytes, 0,
                                  static_cast<size_t>(s->rows - n) * m.row_bytes, st));
    }
  } else {
    if (s->f32.ptr && s->rows > 0)
      PVDB_CUDA(cudaMemsetAsync(s->f32.ptr, 0, static_cast<size_t>(s->rows) * s->ld_f32 * sizeof(float), st));
    if (s->bf16.ptr && s->rows > 0)
      PVDB_CUDA(cudaMemsetAsync(s->bf16.ptr, 0, static_cast<size_t>(s->rows) * s->ld_bf16 * sizeof(__nv_bfloat16), st));
  }
  if (s->active.ptr) {
    PVDB_CUDA(cudaMemsetAsync(s->active.ptr, 0, s->active.bytes, st));
    if (n > 0) {
      const int64_t nwords = (n + 31) >> 5;
      const int blocks = static_cast<int>(std::min<int64_t>((nwords + 255) / 256, kNumSMs * 8));
      set_active_range_kernel<<<blocks, 256, 0, st>>>(static_cast<uint32_t*>(s->active.ptr), 0, n, nullptr);
      PVDB_LAUNCH_CHECK();
    }
  }
  PVDB_CUDA(cudaStreamSynchronize(st));
  s->drop_columns();  // row numbers changed: the host re-uploads the columns it still needs
  s->rows = n;
  return PVDB_OK;
}

// ---------------------------------------------------------------------------- metadata columns
namespace pvdb {
// column[row] = code + 1 (0 = absent)
__global__ void column_write_kernel(uint32_t* __restrict__ col, const int64_t* __restrict__ rows, int64_t row0,
                                    const int32_t* __restrict__ codes, int64_t n) {
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < n;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int64_t row = rows ? rows[i] : row0 + i;
    col[row] = static_cast<uint32_t>(codes[i] + 1);
  }
}

// bits[w] = { row r in word w : active(r) && column[r] - 1 in wanted[] } (& extra[w]); also counts them.
// One thread per row, one ballot per 32 rows.  `wanted` (sorted ascending) is searched linearly when
// short, by bisection otherwise.
__global__ void __launch_bounds__(256) column_filter_kernel(const uint32_t* __restrict__ col,
                                                            const uint32_t* __restrict__ active,
                                                            const uint32_t* __restrict__ extra, int64_t n_rows,
                                                            const int32_t* __restrict__ wanted, int n_wanted,
                                                            uint32_t* __restrict__ bits,
                                                            unsigned long long* __restrict__ count) {
  const int lane = threadIdx.x & 31;
  const int64_t n_words = (n_rows + 31) >> 5;
  unsigned long long local = 0;
  for (int64_t w = (blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x) >> 5; w < n_words;
       w += (static_cast<int64_t>(gridDim.x) * blockDim.x) >> 5) {
    uint32_t aw = __ldg(active + w);
    if (extra) aw &= __ldg(extra + w);
    bool hit = false;
    const int64_t row = (w << 5) + lane;
    if (((aw >> lane) & 1u) && row < n_rows) {
      const int32_t code = static_cast<int32_t>(__ldg(col + row)) - 1;
      if (code >= 0) {
        if (n_wanted <= 8) {
          for (int j = 0; j < n_wanted; ++j) hit |= (wanted[j] == code);
        } else {
          int lo = 0, hi = n_wanted - 1;
          while (lo <= hi) {
            const int mid = (lo + hi) >> 1;
            const int32_t v = wanted[mid];
            if (v == code) { hit = true; break; }
            if (v < code) lo = mid + 1; else hi = mid - 1;
          }
        }
      }
    }
    const uint32_t word = __ballot_sync(0xffffffffu, hit);
    if (lane == 0) {
      bits[w] = word;
      local += __popc(word);
    }
  }
  if (lane == 0 && local) atomicAdd(count, local);
}
}  // namespace pvdb

void pvdb_store::drop_columns() {
  for (DeviceBuffer& col : column) col.release();
}

extern "C" int pvdb_store_column_write(pvdb_store_t* s, int column, const int64_t* rows, int64_t row0,
                                       const int32_t* codes, int64_t n) {
  PVDB_ENTER(s);
  if (column < 0 || column >= pvdb_store::kMaxColumns) return fail(PVDB_ERR_INVALID, "column %d out of range", column);
  if (n == 0) return PVDB_OK;
  if (!codes || n < 0 || row0 < 0) return fail(PVDB_ERR_INVALID, "column_write: bad arguments");
  if (rows) {
    for (int64_t i = 0; i < n; ++i)
      if (rows[i] < 0 || rows[i] >= s->rows)
        return fail(PVDB_ERR_INVALID, "column_write: row %lld outside [0, %lld)", (long long)rows[i], (long long)s->rows);
  } else if (row0 + n > s->rows) {
    return fail(PVDB_ERR_INVALID, "column_write: range [%lld, %lld) outside the store", (long long)row0, (long long)(row0 + n));
  }
  cudaStream_t st = s->stream;
  PVDB_TRY(s->use_stream(st));
  PVDB_TRY(s->column[column].grow(static_cast<size_t>(s->capacity) * sizeof(uint32_t), st));
  PVDB_TRY(s->d_in.ensure(static_cast<size_t>(n) * sizeof(int32_t)));
  PVDB_CUDA(cudaMemcpyAsync(s->d_in.ptr, codes, static_cast<size_t>(n) * sizeof(int32_t), cudaMemcpyHostToDevice, st));
  const int64_t* d_rows = nullptr;
  if (rows) {
    PVDB_TRY(s->d_rows.ensure(static_cast<size_t>(n) * sizeof(int64_t)));
    PVDB_CUDA(cudaMemcpyAsync(s->d_rows.ptr, rows, static_cast<size_t>(n) * sizeof(int64_t), cudaMemcpyHostToDevice, st));
    d_rows = static_cast<const int64_t*>(s->d_rows.ptr);
  }
  const int blocks = static_cast<int>(std::max<int64_t>(1, std::min<int64_t>((n + 255) / 256, kNumSMs * 8)));
  column_write_kernel<<<blocks, 256, 0, st>>>(static_cast<uint32_t*>(s->column[column].ptr), d_rows, row0,
                                              static_cast<const int32_t*>(s->d_in.ptr), n);
  PVDB_LAUNCH_CHECK();
  PVDB_CUDA(cudaStreamSynchronize(st));
  return PVDB_OK;
}

extern "C" int pvdb_store_column_drop(pvdb_store_t* s, int column) {
  PVDB_ENTER(s);
  if (column < 0 || column >= pvdb_store::kMaxColumns) return fail(PVDB_ERR_INVALID, "column %d out of range", column);
  PVDB_CUDA(cudaStreamSynchronize(s->stream));
  s->column[column].release();
  return PVDB_OK;
}

namespace pvdb {
// Build the prefilter bitmap for (column in wanted) on the device; returns the device bitmap and
// leaves the eligible-row count in *d_count (device).  Used by pvdb_search_where (api.cu).
int build_column_filter(pvdb_store* s, int column, const int32_t* wanted, int n_wanted, const uint32_t* extra_bits,
                        const uint32_t** d_bits_out, unsigned long long** d_count_out, cudaStream_t st) {
  if (column < 0 || column >= pvdb_store::kMaxColumns || s->column[column].ptr == nullptr)
    return fail(PVDB_ERR_INVALID, "search_where: column %d has not been written", column);
  if (n_wanted < 0 || (n_wanted > 0 && !wanted)) return fail(PVDB_ERR_INVALID, "search_where: bad wanted list");
  const size_t n_words = static_cast<size_t>((s->rows + 31) >> 5);
  // scratch layout: [bitmap][extra bitmap][wanted codes][count]
  const size_t bm = (n_words * sizeof(uint32_t) + 255) & ~size_t(255);
  const size_t wb = (static_cast<size_t>(n_wanted) * sizeof(int32_t) + 255) & ~size_t(255);
  PVDB_TRY(s->d_prefilter.ensure(2 * bm + wb + 256));
  unsigned char* base = static_cast<unsigned char*>(s->d_prefilter.ptr);
  uint32_t* d_bits = reinterpret_cast<uint32_t*>(base);
  uint32_t* d_extra = nullptr;
  int32_t* d_wanted = reinterpret_cast<int32_t*>(base + 2 * bm);
  unsigned long long* d_count = reinterpret_cast<unsigned long long*>(base + 2 * bm + wb);
  if (extra_bits) {
    d_extra = reinterpret_cast<uint32_t*>(base + bm);
    PVDB_CUDA(cudaMemcpyAsync(d_extra, extra_bits, n_words * sizeof(uint32_t), cudaMemcpyHostToDevice, st));
  }
  if (n_wanted) PVDB_CUDA(cudaMemcpyAsync(d_wanted, wanted, static_cast<size_t>(n_wanted) * sizeof(int32_t), cudaMemcpyHostToDevice, st));
  PVDB_CUDA(cudaMemsetAsync(d_count, 0, sizeof(unsigned long long), st));
  const int64_t threads = static_cast<int64_t>(n_words) * 32;
  const int blocks = static_cast<int>(std::max<int64_t>(1, std::min<int64_t>((threads + 255) / 256, kNumSMs * 16)));
  column_filter_kernel<<<blocks, 256, 0, st>>>(static_cast<const uint32_t*>(s->column[column].ptr),
                                               static_cast<const uint32_t*>(s->active.ptr), d_extra, s->rows, d_wanted,
                                               n_wanted, d_bits, d_count);
  PVDB_LAUNCH_CHECK();
  *d_bits_out = d_bits;
  *d_count_out = d_count;
  return PVDB_OK;
}
}  // namespace pvdb
