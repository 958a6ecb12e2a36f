// Scan kernel instantiations: bf16 rows on mma.sync, two queries per pass, sparse walk (see scan_kernel.cuh).
#include "scan_kernel.cuh"

namespace pvdb {
template int launch_scan_mma_multi_variant<true>(const ScanParams&, cudaStream_t);
}  // namespace pvdb
