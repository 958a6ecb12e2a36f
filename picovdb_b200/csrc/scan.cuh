// Single-query masked scan + fused top-k ("GEMV path"): the HBM-bound replacement of
//   scores = q @ V.T ; argpartition ; argsort        (picovdb/pico_vdb.py:683-714)
#pragma once

#include "common.cuh"
#include "exchange.cuh"

namespace pvdb {

// launch shape of the scan kernel (tunable at build time for experiments)
#ifndef PVDB_SCAN_THREADS
#define PVDB_SCAN_THREADS 512
#endif
#ifndef PVDB_SCAN_BLOCKS_PER_SM
#define PVDB_SCAN_BLOCKS_PER_SM 2
#endif
constexpr int kScanThreads = PVDB_SCAN_THREADS;
constexpr int kScanWarps = kScanThreads / 32;
constexpr int kScanBlocksPerSM = PVDB_SCAN_BLOCKS_PER_SM;

struct ScanParams {
  const void* matrix;        // fp32 or bf16 row-major matrix
  int64_t n_rows;            // rows to scan (high-water mark)
  int row_chunks;            // 16-byte chunks per row (ld * sizeof(T) / 16)
  const uint32_t* active;    // active bitmap (always present)
  const uint32_t* prefilter; // optional second bitmap, ANDed with `active`
  const float* query;        // normalised query, padded with zeros to a multiple of 8 floats
  const float* raw_query;    // if non-null: raw query of `dim` floats, normalised inside the kernel
  int dim;
  int query_floats;          // padded query length
  int k;                     // 1 .. kFusedK for this pass
  int nq;                    // queries this launch serves: 1, or up to NQ in the several-queries-per-pass kernels
                             // (query j of the launch is query qsel[j] of query / raw_query and writes result
                             // row qsel[j] of out_scores / out_rows)
  int64_t qsel[4];
  const uint64_t* upper;     // only keys strictly below *upper qualify (paging); NULL = no bound
  uint64_t* partial;         // [gridDim.x][k] per-block lists
  unsigned int* ticket;      // last-block-done counter (self-resetting)
  uint64_t* next_upper;      // receives the k-th key of this pass (0 when fewer than k found)
  unsigned long long* floor_key;  // max over the blocks of "k-th key of the block's list": no key below it can be
                             // in the top k, so the last block admits nothing smaller (self-resetting)
  float* out_scores;         // [k]
  int64_t* out_rows;         // [k]
  int64_t row_base;
  int pdl;                   // host side: launch with programmatic dependent launch (the previous launch on
                             // the stream is a scan of the same call: its inputs are complete, only the
                             // per-block lists / ticket / paging bound are shared -- the kernel waits for
                             // the previous grid right before it touches those)
  ExchangeView xv;           // xv.world > 0: the last block exchanges its list with the peer GPUs and
                             // merges theirs before writing the result (exchange.cuh); not with paging
};

// Launch one scan pass on `stream`.  `is_bf16` selects the mirror layout.
int launch_scan(const ScanParams& p, bool is_bf16, cudaStream_t stream);
// Several queries per pass over the matrix (scan_kernel.cuh, "several queries per pass"): how many queries one
// launch can take for this store layout and k (1 = only the single-query kernels apply), and the launch itself
// (p.nq in [2, scan_multi_width]; no paging bound, no exchange).
int scan_multi_width(bool is_bf16, int query_floats, int k);
int launch_scan_multi(const ScanParams& p, bool is_bf16, cudaStream_t stream);
int scan_grid_blocks();

// Query preparation (picovdb/pico_vdb.py:584-591): L2-normalise each query in fp32, zero -> e0,
// write nq x ldq fp32 (zero padded) and optionally a bf16 copy with row stride ldq.  d_qeps (optional,
// nq x 4 floats): per query ||q - tf32(q)||, ||q - bf16(q)||, ||q|| (the exactness guard's inputs).
int launch_prepare_queries(const float* d_raw, int64_t nq, int dim, bool already_normalised, float* d_qn,
                           __nv_bfloat16* d_qn16, int ldq, float* d_qeps, cudaStream_t stream);

int launch_merge_topk(const float* d_scores, const int64_t* d_rows, int nlists, int64_t nq, int k,
                      int64_t scores_stride, int64_t rows_stride, float* d_out_scores, int64_t* d_out_rows,
                      cudaStream_t stream);

}  // namespace pvdb
