// Cross-GPU top-k exchange over peer memory (NVLink / NVSwitch), fused into the kernels that
// produce the per-shard lists.  SURVEY.md 8(e): rows are sharded over the GPUs of one box, every GPU
// selects a local top-k, the lists are exchanged and k-way merged.  Round 1 did the exchange with one
// NCCL all-gather + a merge kernel per query (20-30 us of launch latency for 120 bytes); here every
// GPU owns a MAILBOX that all peers can write (CUDA IPC between processes, peer access inside one
// process), and the kernel that finishes a local list
//   1. stores it, as 64-bit keys with GLOBAL row numbers, into its slot of EVERY peer's mailbox,
//   2. publishes a sequence number in each peer's flag word (st.release.sys after a system fence),
//   3. spins until all peers' flags in its OWN mailbox carry that sequence number (ld.acquire.sys),
//   4. merges the `world` lists it now holds and writes the final result
// -- no collective call, no extra launch, every GPU ends with the same answer.
//
// Slots are double buffered by the parity of the sequence number.  That is enough: a GPU can only
// start launch s+2 after its launch s+1 completed, which needed every peer's flag of s+1, which a
// peer writes only after its own launch s (stream order) has finished reading the parity-s slots.
//
// One process per GPU or one GPU per mailbox ONLY: kernels that wait on each other must never share
// a GPU (B200_PROFILING.md).  The wait is bounded: after ~4 s without the peers' flags the kernel
// traps, so a missing rank surfaces as a CUDA error instead of a hang.
#pragma once

#include "common.cuh"

namespace pvdb {

constexpr int kMaxWorld = 8;       // GPUs of one box
constexpr int kMaxSlices = 64;     // independent query ranges of one exchange launch (one flag each)
constexpr size_t kBoxHeaderBytes = 2 * kMaxWorld * kMaxSlices * sizeof(uint64_t);  // flags[parity][src][slice]

// What a kernel needs to take part in one exchange (passed by value).
struct ExchangeView {
  unsigned char* box[kMaxWorld];  // every rank's mailbox as mapped into THIS process (box[rank] is local)
  int world;                      // 0: no exchange (1 = a self-mailbox, used by the single-GPU tests)
  int rank;
  uint64_t seq;                   // sequence number of this launch (1, 2, ...), the same on every rank
  int64_t slot_keys;              // capacity of one (parity, source rank) slot, in keys
};

#ifdef __CUDACC__
__device__ __forceinline__ uint64_t* xv_flag(unsigned char* box, int parity, int src, int slice) {
  return reinterpret_cast<uint64_t*>(box) + (static_cast<size_t>(parity) * kMaxWorld + src) * kMaxSlices + slice;
}
__device__ __forceinline__ uint64_t* xv_slot(const ExchangeView& v, unsigned char* box, int parity, int src) {
  return reinterpret_cast<uint64_t*>(box + kBoxHeaderBytes) +
         (static_cast<size_t>(parity) * v.world + src) * static_cast<size_t>(v.slot_keys);
}
__device__ __forceinline__ void st_release_sys(uint64_t* p, uint64_t v) {
  asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ uint64_t ld_acquire_sys(const uint64_t* p) {
  uint64_t v;
  asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ unsigned long long global_timer_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
// Spin until *flag >= seq.  Bounded: traps after ~4 s (a rank that never launched, a dead peer).
__device__ __forceinline__ void xv_wait_flag(const uint64_t* flag, uint64_t seq) {
  unsigned long long t0 = 0;
  for (unsigned spin = 0;; ++spin) {
    if (ld_acquire_sys(flag) >= seq) return;
    if ((spin & 1023u) == 1023u) {
      const unsigned long long now = global_timer_ns();
      if (t0 == 0) t0 = now;
      else if (now - t0 > 4000000000ull) __trap();
    }
  }
}
// key with a shard-local row -> key with the global row (the row lives negated in the low word)
__device__ __forceinline__ uint64_t key_to_global(uint64_t key, int64_t row_base) {
  return key ? key - static_cast<uint64_t>(row_base) : 0ull;
}
#endif  // __CUDACC__

}  // namespace pvdb

// Host-side state of one rank's end of the exchange (C ABI handle pvdb_exchange_t).
struct pvdb_exchange {
  int device = 0;
  int world = 1;
  int rank = 0;
  int64_t slot_keys = 0;
  size_t box_bytes = 0;
  unsigned char* box = nullptr;                      // this rank's mailbox (cudaMalloc: IPC exportable)
  unsigned char* peers[pvdb::kMaxWorld] = {};        // all mailboxes, peers[rank] == box
  bool ipc_mapped[pvdb::kMaxWorld] = {};             // opened with cudaIpcOpenMemHandle (to be closed)
  bool connected = false;
  uint64_t seq = 0;                                  // launches so far; identical on every rank by construction

  pvdb::ExchangeView next_view() {
    pvdb::ExchangeView v{};
    for (int i = 0; i < world; ++i) v.box[i] = peers[i];
    v.world = world;
    v.rank = rank;
    v.seq = ++seq;
    v.slot_keys = slot_keys;
    return v;
  }
};

namespace pvdb {
// Publish local (scores, rows) lists [nq][k] (rows already global, -1 = empty) to every peer, wait for
// theirs, merge, write [nq][k].  One launch, `kMaxSlices` independent query ranges at most.
int launch_exchange_merge(pvdb_exchange* ex, const float* d_loc_scores, const int64_t* d_loc_rows, int64_t nq, int k,
                          float* d_out_scores, int64_t* d_out_rows, cudaStream_t st);
}  // namespace pvdb
