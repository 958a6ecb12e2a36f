// Peer-memory mailboxes for the cross-GPU top-k exchange, and the stand-alone exchange + merge
// kernel used after a batched search (the single-query scan kernel carries the same steps in its
// last block, scan_kernel.cuh).  Protocol and safety argument: exchange.cuh.
#include <algorithm>
#include <cstring>

#include "exchange.cuh"

namespace pvdb {

// One block per slice = contiguous range of queries; blocks never wait on blocks of their own grid,
// only on the SAME slice of the peers' grids, whose publish step never waits: no deadlock whatever
// the order in which blocks get scheduled.
template <int S>
__global__ void __launch_bounds__(256) exchange_merge_kernel(const ExchangeView v, const float* __restrict__ loc_scores,
                                                             const int64_t* __restrict__ loc_rows, int64_t nq, int k,
                                                             int q_per_slice, float* __restrict__ out_scores,
                                                             int64_t* __restrict__ out_rows) {
  const int slice = blockIdx.x;
  const int64_t q0 = static_cast<int64_t>(slice) * q_per_slice;
  const int64_t q1 = min(nq, q0 + q_per_slice);
  const int64_t off = q0 * k;
  const int n = static_cast<int>((q1 - q0) * k);
  const int parity = static_cast<int>(v.seq & 1ull);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

  // 1. publish this rank's lists of the slice, as keys, into every mailbox (its own included)
  for (int i = tid; i < n; i += blockDim.x) {
    const int64_t r = loc_rows[off + i];
    const float sc = loc_scores[off + i];
    const uint64_t key = (r >= 0 && sc == sc) ? make_key(sc, static_cast<uint32_t>(r)) : 0ull;
    for (int p = 0; p < v.world; ++p) xv_slot(v, v.box[p], parity, v.rank)[off + i] = key;
  }
  __threadfence_system();
  __syncthreads();
  // 2. tell every peer; 3. wait for every peer
  if (tid < v.world) st_release_sys(xv_flag(v.box[tid], parity, v.rank, slice), v.seq);
  if (tid < v.world) xv_wait_flag(xv_flag(v.box[v.rank], parity, tid, slice), v.seq);
  __syncthreads();
  // 4. merge: one warp per query
  for (int64_t q = q0 + warp; q < q1; q += blockDim.x / 32) {
    WarpList<S> L;
    L.clear();
    uint64_t thr = 0ull;
    for (int r = 0; r < v.world; ++r)
      merge_list<true, S>(L, thr, xv_slot(v, v.box[v.rank], parity, r) + q * k, k, k, lane);
#pragma unroll
    for (int s = 0; s < S; ++s) {
      const int e = s * 32 + lane;
      if (e < k) {
        const uint64_t key = L.slot[s];
        out_scores[q * k + e] = key ? key_score(key) : -INFINITY;
        out_rows[q * k + e] = key ? static_cast<int64_t>(key_row(key)) : -1ll;
      }
    }
  }
}

int launch_exchange_merge(pvdb_exchange* ex, const float* d_loc_scores, const int64_t* d_loc_rows, int64_t nq, int k,
                          float* d_out_scores, int64_t* d_out_rows, cudaStream_t st) {
  if (!ex || !ex->connected) return fail(PVDB_ERR_INVALID, "exchange: not connected");
  if (k < 1 || k > kFusedK) return fail(PVDB_ERR_UNSUPPORTED, "exchange: k=%d outside [1, %d]", k, kFusedK);
  if (nq * k > ex->slot_keys)
    return fail(PVDB_ERR_INVALID, "exchange: %lld x %d results exceed the mailbox slot (%lld keys)", (long long)nq, k,
                (long long)ex->slot_keys);
  if (nq == 0) return PVDB_OK;
  const int q_per_slice = static_cast<int>((nq + kMaxSlices - 1) / kMaxSlices);
  const int slices = static_cast<int>((nq + q_per_slice - 1) / q_per_slice);
  const ExchangeView v = ex->next_view();
  if (k <= 32)
    exchange_merge_kernel<1><<<slices, 256, 0, st>>>(v, d_loc_scores, d_loc_rows, nq, k, q_per_slice, d_out_scores,
                                                     d_out_rows);
  else
    exchange_merge_kernel<4><<<slices, 256, 0, st>>>(v, d_loc_scores, d_loc_rows, nq, k, q_per_slice, d_out_scores,
                                                     d_out_rows);
  PVDB_LAUNCH_CHECK();
  return PVDB_OK;
}

}  // namespace pvdb

using namespace pvdb;

static_assert(sizeof(cudaIpcMemHandle_t) == PVDB_IPC_HANDLE_BYTES, "header constant must match the runtime's handle");

extern "C" int pvdb_exchange_create(pvdb_exchange_t** out, int device, int world, int rank, int64_t slot_keys) {
  if (!out) return fail(PVDB_ERR_INVALID, "out is null");
  *out = nullptr;
  if (world < 1 || world > kMaxWorld || rank < 0 || rank >= world || slot_keys < 1)
    return fail(PVDB_ERR_INVALID, "exchange_create: world=%d (max %d), rank=%d, slot_keys=%lld", world, kMaxWorld, rank,
                (long long)slot_keys);
  PVDB_CUDA(cudaSetDevice(device));
  pvdb_exchange* ex = new pvdb_exchange();
  ex->device = device;
  ex->world = world;
  ex->rank = rank;
  ex->slot_keys = slot_keys;
  ex->box_bytes = kBoxHeaderBytes + static_cast<size_t>(2) * world * slot_keys * sizeof(uint64_t);
  cudaError_t e = cudaMalloc(reinterpret_cast<void**>(&ex->box), ex->box_bytes);
  if (e == cudaSuccess) e = cudaMemset(ex->box, 0, ex->box_bytes);
  if (e == cudaSuccess) e = cudaDeviceSynchronize();
  if (e != cudaSuccess) {
    if (ex->box) cudaFree(ex->box);
    delete ex;
    (void)cudaGetLastError();
    return fail(e == cudaErrorMemoryAllocation ? PVDB_ERR_OOM : PVDB_ERR_CUDA, "exchange_create: %s", cudaGetErrorString(e));
  }
  ex->peers[rank] = ex->box;
  ex->connected = (world == 1);
  *out = ex;
  return PVDB_OK;
}

extern "C" int pvdb_exchange_ipc_handle(pvdb_exchange_t* ex, void* out_handle) {
  if (!ex || !out_handle) return fail(PVDB_ERR_INVALID, "exchange_ipc_handle: null argument");
  PVDB_CUDA(cudaSetDevice(ex->device));
  cudaIpcMemHandle_t h;
  PVDB_CUDA(cudaIpcGetMemHandle(&h, ex->box));
  std::memcpy(out_handle, &h, sizeof(h));
  return PVDB_OK;
}

extern "C" int pvdb_exchange_connect_ipc(pvdb_exchange_t* ex, const void* handles) {
  if (!ex || !handles) return fail(PVDB_ERR_INVALID, "exchange_connect_ipc: null argument");
  PVDB_CUDA(cudaSetDevice(ex->device));
  const unsigned char* hb = static_cast<const unsigned char*>(handles);
  for (int p = 0; p < ex->world; ++p) {
    if (p == ex->rank || ex->peers[p] != nullptr) continue;
    cudaIpcMemHandle_t h;
    std::memcpy(&h, hb + static_cast<size_t>(p) * sizeof(h), sizeof(h));
    void* ptr = nullptr;
    PVDB_CUDA(cudaIpcOpenMemHandle(&ptr, h, cudaIpcMemLazyEnablePeerAccess));
    ex->peers[p] = static_cast<unsigned char*>(ptr);
    ex->ipc_mapped[p] = true;
  }
  ex->connected = true;
  return PVDB_OK;
}

extern "C" int pvdb_exchange_connect_local(pvdb_exchange_t** exs, int world) {
  if (!exs || world < 1 || world > kMaxWorld) return fail(PVDB_ERR_INVALID, "exchange_connect_local: bad arguments");
  for (int i = 0; i < world; ++i) {
    if (!exs[i] || exs[i]->world != world || exs[i]->rank != i)
      return fail(PVDB_ERR_INVALID, "exchange_connect_local: handle %d does not belong to this group", i);
    for (int j = 0; j < i; ++j)
      if (exs[j]->device == exs[i]->device)
        return fail(PVDB_ERR_INVALID, "exchange_connect_local: device %d appears twice (kernels that wait on "
                                      "each other must not share a GPU)", exs[i]->device);
  }
  for (int i = 0; i < world; ++i) {
    PVDB_CUDA(cudaSetDevice(exs[i]->device));
    for (int j = 0; j < world; ++j) {
      if (i == j) continue;
      int can = 0;
      PVDB_CUDA(cudaDeviceCanAccessPeer(&can, exs[i]->device, exs[j]->device));
      if (!can) return fail(PVDB_ERR_UNSUPPORTED, "device %d cannot access device %d", exs[i]->device, exs[j]->device);
      cudaError_t e = cudaDeviceEnablePeerAccess(exs[j]->device, 0);
      if (e == cudaErrorPeerAccessAlreadyEnabled) (void)cudaGetLastError();
      else PVDB_CUDA(e);
      exs[i]->peers[j] = exs[j]->box;
    }
    exs[i]->connected = true;
  }
  return PVDB_OK;
}

extern "C" int pvdb_exchange_disconnect(pvdb_exchange_t* ex) {
  if (!ex) return PVDB_OK;
  PVDB_CUDA(cudaSetDevice(ex->device));
  PVDB_CUDA(cudaDeviceSynchronize());
  for (int p = 0; p < ex->world; ++p) {
    if (p == ex->rank) continue;
    if (ex->ipc_mapped[p] && ex->peers[p]) PVDB_CUDA(cudaIpcCloseMemHandle(ex->peers[p]));
    ex->ipc_mapped[p] = false;
    ex->peers[p] = nullptr;
  }
  ex->connected = (ex->world == 1);
  return PVDB_OK;
}

extern "C" int pvdb_exchange_destroy(pvdb_exchange_t* ex) {
  if (!ex) return PVDB_OK;
  cudaSetDevice(ex->device);
  cudaDeviceSynchronize();
  for (int p = 0; p < ex->world; ++p)
    if (ex->ipc_mapped[p] && ex->peers[p]) cudaIpcCloseMemHandle(ex->peers[p]);
  if (ex->box) cudaFree(ex->box);
  (void)cudaGetLastError();
  delete ex;
  return PVDB_OK;
}

extern "C" int pvdb_exchange_info(pvdb_exchange_t* ex, int* out_world, int* out_rank, int64_t* out_slot_keys,
                                  int64_t* out_launches) {
  if (!ex) return fail(PVDB_ERR_INVALID, "null exchange handle");
  if (out_world) *out_world = ex->world;
  if (out_rank) *out_rank = ex->rank;
  if (out_slot_keys) *out_slot_keys = ex->slot_keys;
  if (out_launches) *out_launches = static_cast<int64_t>(ex->seq);
  return PVDB_OK;
}
