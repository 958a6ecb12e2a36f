// Device-resident vector store: the B200 replacement of the reference's `_vectors` matrix and
// `_active_indices` array (picovdb/pico_vdb.py:136,143).
#pragma once

#include <condition_variable>
#include <mutex>
#include <vector>

#include "common.cuh"

namespace pvdb {

// One growable device allocation.  Growth keeps the contents and zero-fills the new tail.
// Backed by CUDA virtual memory management when the driver offers it: a large address range is
// reserved once and physical memory is mapped behind it as the store grows, so growth never copies,
// never needs 2x transient memory, and the base pointer (hence every TMA tensor map) stays valid.
// (The reference re-allocates and copies the whole matrix on every append, pico_vdb.py:451-462.)
struct DeviceBuffer {
  void* ptr = nullptr;
  size_t bytes = 0;      // usable (mapped) bytes
  bool vmm = false;
  size_t va_bytes = 0;   // reserved address range (vmm)
  struct Chunk { unsigned long long handle; size_t offset, size; };
  std::vector<Chunk> chunks;
  int grow(size_t new_bytes, cudaStream_t stream);  // no-op when new_bytes <= bytes
  void release();
};

// Scratch that grows on demand and is reused between calls (never shrinks).
struct Scratch {
  void* ptr = nullptr;
  size_t bytes = 0;
  bool pinned_host = false;
  uint64_t gen = 0;  // bumped on every (re)allocation
  int ensure(size_t need);
  void release();
};

}  // namespace pvdb

struct pvdb_store;
namespace pvdb {
// (column in wanted) & active (& extra_bits) -> device bitmap + device count (store.cu)
int build_column_filter(pvdb_store* s, int column, const int32_t* wanted, int n_wanted, const uint32_t* extra_bits,
                        const uint32_t** d_bits_out, unsigned long long** d_count_out, cudaStream_t st);
}  // namespace pvdb

// bytes kept allocated past the last row of the fp32 / bf16 matrices
constexpr size_t kTailSlack = 256;

struct pvdb_store {
  int device = 0;
  int dim = 0;
  int ld_f32 = 0;   // fp32 row stride in elements (multiple of 4 -> 16-byte aligned rows)
  int ld_bf16 = 0;  // bf16 row stride in elements (multiple of 8 -> 16-byte aligned rows)
  int ldq = 0;      // padded query length in floats (multiple of 64, zero padded)
  int flags = 0;
  int64_t rows = 0;      // high-water mark of slots in use
  int64_t capacity = 0;  // slots allocated (multiple of 1024)
  int64_t row_base = 0;

  pvdb::DeviceBuffer f32;     // capacity x ld_f32 floats, pad columns are zero
  pvdb::DeviceBuffer bf16;    // capacity x ld_bf16 bf16, pad columns are zero
  pvdb::DeviceBuffer active;  // capacity/32 words
  static constexpr int kMaxColumns = 16;
  pvdb::DeviceBuffer column[kMaxColumns];  // metadata codes + 1 per row (0 = absent), see pvdb_store_column_write

  cudaStream_t stream = nullptr;       // the store's own stream (host entry points)
  cudaStream_t last_stream = nullptr;  // stream of the most recent call (cross-stream ordering)
  cudaEvent_t order_event = nullptr;
  std::mutex mu;

  // per-call scratch (guarded by mu)
  pvdb::Scratch d_in;       // staged host inputs (vectors / queries)
  pvdb::Scratch d_rows;     // staged row indices
  pvdb::Scratch d_prefilter;
  pvdb::Scratch d_qn;       // normalised queries, nq x ldq fp32
  pvdb::Scratch d_qn16;     // normalised queries in bf16 (tensor path)
  pvdb::Scratch d_partial;  // per-block candidate lists + scan control words
  pvdb::Scratch d_out;      // results before the D2H copy
  pvdb::Scratch d_misc;
  pvdb::Scratch h_pinned;   // pinned bounce buffer for results
  pvdb::Scratch d_qeps;     // per query: ||q - tf32(q)||, ||q - bf16(q)|| (exactness guard of the tensor paths)
  pvdb::Scratch d_flag;     // guard: [count][flagged query indices]
  // Host <-> device streaming (bulk ingest, load, save): two pinned buffers; the host-side copy of
  // block i+1 (several threads) overlaps the DMA of block i.  pipe_ev[b] = last DMA that used pin[b].
  pvdb::Scratch h_pipe[2];
  cudaEvent_t pipe_ev[2] = {nullptr, nullptr};
  pvdb::Scratch d_xloc;     // this shard's lists before the cross-GPU exchange: [nq*k] rows, then scores
  pvdb::Scratch h_flag;     // pinned copy of d_flag
  // Largest input-rounding error over the rows ever written, as the uint image of two non-negative
  // floats: [0] = max_r ||v_r - tf32_trunc(v_r)||^2, [1] = max_r ||v_r - bf16_rn(v_r)||^2.  Only
  // maintained when the store keeps the fp32 matrix (otherwise the bf16 rows ARE the exact data).
  uint32_t* d_err_words = nullptr;
  int64_t guard_flagged_last = 0;   // queries of the last search that fell back to the exact scan
  int64_t guard_flagged_total = 0;
  uint64_t partial_gen_inited = 0;  // d_partial.gen whose control words have been zeroed

  // Result slots of the host search entry points.  A search holds the store mutex only while it
  // ENQUEUES its copies and kernels on the store's stream (stream order keeps the shared device scratch
  // safe); it then waits for its own event and copies its results out of its own pinned buffer with
  // the mutex released, so concurrent readers of one store overlap their host-side latency (submit,
  // wake-up, result copy) with each other's GPU work -- the reference computes outside its lock too
  // (pico_vdb.py:670-714).
  static constexpr int kIoSlots = 8;
  struct IoSlot {
    pvdb::Scratch h_res;          // pinned: kernels write small results straight into it
    cudaEvent_t done = nullptr;
    bool busy = false;
  };
  IoSlot io[kIoSlots];
  std::mutex io_mu;
  std::condition_variable io_cv;
  int acquire_io_slot();
  void release_io_slot(int slot);

  // Make `s` the stream the store's data is ordered on (inserts an event edge when it changes).
  int use_stream(cudaStream_t s);
  int ensure_capacity(int64_t need_rows, cudaStream_t s);
  void drop_columns();
};
