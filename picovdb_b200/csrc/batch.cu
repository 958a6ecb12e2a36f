#include "batch.cuh"
#include "store.cuh"

namespace pvdb {

bool batch_path_available() { return false; }

int search_batch(pvdb_store*, bool, const float*, const __nv_bfloat16*, int64_t, int, const uint32_t*, bool, float*,
                 int64_t*, cudaStream_t) {
  return fail(PVDB_ERR_UNSUPPORTED, "batched tensor-core path not built");
}

}  // namespace pvdb
