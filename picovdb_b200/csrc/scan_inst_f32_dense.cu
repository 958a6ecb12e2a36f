// Scan kernel instantiations: f32, dense walk (see scan_kernel.cuh).
#include "scan_kernel.cuh"

namespace pvdb {
template int launch_scan_variant<false, false>(const ScanParams&, int, int, cudaStream_t);
}  // namespace pvdb
