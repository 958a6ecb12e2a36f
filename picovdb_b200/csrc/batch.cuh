// Batched queries: tcgen05 / TMEM tensor-core GEMM with a fused mask + top-k epilogue, followed by
// an fp32 re-scoring pass (replaces the batched form of picovdb/pico_vdb.py:683-714).
#pragma once

#include "common.cuh"

struct pvdb_store;

namespace pvdb {

constexpr int64_t kBatchMinQueries = 2;  // measured: one tensor-core pass (84 % of HBM peak) beats two scans

bool batch_path_available();
// largest k the fused tensor-core path selects in one pass (larger k uses the paged scan path)
// Rows the prepared query matrix must hold for a batch of nq queries: whole 128-row TMA boxes
// (zero rows past nq).
int64_t batch_query_rows(int64_t nq);
int batch_max_k(bool use_bf16, bool rescore);

// d_qn: nq x ldq normalised fp32 queries; d_qn16: the same in bf16 (only for use_bf16).
// d_qeps / d_flag_count / d_flag_list: exactness guard (see GuardParams in batch.cu); d_qeps == NULL
// switches it off.  Flagged queries are appended to d_flag_list (the caller zeroes *d_flag_count).
int search_batch(pvdb_store* s, bool use_bf16, const float* d_qn, const __nv_bfloat16* d_qn16, int64_t nq, int k,
                 const uint32_t* d_pref, bool no_rescore, const float* d_qeps, unsigned* d_flag_count,
                 int* d_flag_list, float* d_out_scores, int64_t* d_out_rows, cudaStream_t st);

}  // namespace pvdb
