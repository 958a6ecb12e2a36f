// Scan kernel instantiations: f32, sparse walk (see scan_kernel.cuh).
#include "scan_kernel.cuh"

namespace pvdb {
template int launch_scan_variant<false, true>(const ScanParams&, int, int, cudaStream_t);
}  // namespace pvdb
