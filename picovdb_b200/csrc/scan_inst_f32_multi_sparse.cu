// Scan kernel instantiations: f32, several queries per pass, sparse walk (see scan_kernel.cuh).
#include "scan_kernel.cuh"

namespace pvdb {
template int launch_scan_multi_variant<true>(const ScanParams&, int, int, cudaStream_t);
}  // namespace pvdb
