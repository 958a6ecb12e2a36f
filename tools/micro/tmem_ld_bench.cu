// Micro-benchmark: how fast can epilogue warps read a TMEM accumulator?  (B200, sm_100a)
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/micro/tmem_ld_bench tools/micro/tmem_ld_bench.cu
// One CTA per SM allocates all 512 TMEM columns; W warps per SM sub-partition loop over
// tcgen05.ld.32x32b.xN of their lane quarter (the data is whatever TMEM holds: only timing matters).
// Prints cycles per load instruction and the implied bytes / clock / SM for each (W, N, depth) --
// depth = loads in flight per warp before the wait.
#include <cstdio>
#include <cuda_runtime.h>
#include <stdint.h>

template <int N>
__device__ __forceinline__ void tmem_ld(uint32_t taddr, uint32_t (&v)[N]);

template <>
__device__ __forceinline__ void tmem_ld<32>(uint32_t taddr, uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
        "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
        "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
        "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr)
      : "memory");
}
template <>
__device__ __forceinline__ void tmem_ld<16>(uint32_t taddr, uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
        "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
      : "r"(taddr)
      : "memory");
}
template <>
__device__ __forceinline__ void tmem_ld<64>(uint32_t taddr, uint32_t (&v)[64]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x64.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, "
      "%32, %33, %34, %35, %36, %37, %38, %39, %40, %41, %42, %43, %44, %45, %46, %47, "
      "%48, %49, %50, %51, %52, %53, %54, %55, %56, %57, %58, %59, %60, %61, %62, %63}, [%64];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
        "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
        "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
        "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31]),
        "=r"(v[32]), "=r"(v[33]), "=r"(v[34]), "=r"(v[35]), "=r"(v[36]), "=r"(v[37]), "=r"(v[38]), "=r"(v[39]),
        "=r"(v[40]), "=r"(v[41]), "=r"(v[42]), "=r"(v[43]), "=r"(v[44]), "=r"(v[45]), "=r"(v[46]), "=r"(v[47]),
        "=r"(v[48]), "=r"(v[49]), "=r"(v[50]), "=r"(v[51]), "=r"(v[52]), "=r"(v[53]), "=r"(v[54]), "=r"(v[55]),
        "=r"(v[56]), "=r"(v[57]), "=r"(v[58]), "=r"(v[59]), "=r"(v[60]), "=r"(v[61]), "=r"(v[62]), "=r"(v[63])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// N columns per load, DEPTH loads issued back to back before one wait
template <int N, int DEPTH>
__global__ void __launch_bounds__(640, 1) bench(int warps_per_quarter, int iters, unsigned long long* out_cycles,
                                                unsigned* out_sink) {
  __shared__ uint32_t tmem_slot;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                     static_cast<uint32_t>(__cvta_generic_to_shared(&tmem_slot))),
                 "r"(512)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t base = tmem_slot;
  const int ew = warp & 3;
  const int idx = warp >> 2;  // which warp of its quarter
  unsigned acc = 0;
  long long t0 = 0, t1 = 0;
  if (idx < warps_per_quarter) {
    uint32_t v[DEPTH][N];
    const uint32_t lane_base = base + (static_cast<uint32_t>(ew * 32) << 16);
    t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
      for (int d = 0; d < DEPTH; ++d) {
        const uint32_t col = static_cast<uint32_t>(((it * DEPTH + d) * N + idx * 64) & 511) & ~static_cast<uint32_t>(N - 1);
        tmem_ld<N>(lane_base + col, v[d]);
      }
      tmem_wait();
#pragma unroll
      for (int d = 0; d < DEPTH; ++d) acc += v[d][0] ^ v[d][N - 1];
    }
    t1 = clock64();
  }
  __syncthreads();
  if (lane == 0 && idx < warps_per_quarter) {
    atomicAdd(out_cycles, static_cast<unsigned long long>(t1 - t0));
    if (acc == 0x12345678u) *out_sink = acc;
  }
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(base), "r"(512) : "memory");
}

template <int N, int DEPTH>
static void run(int wpq) {
  unsigned long long* d_cyc;
  unsigned* d_sink;
  cudaMalloc(&d_cyc, 8);
  cudaMalloc(&d_sink, 4);
  const int iters = 4096 / DEPTH;
  for (int rep = 0; rep < 2; ++rep) {
    cudaMemset(d_cyc, 0, 8);
    bench<N, DEPTH><<<148, 32 * 4 * wpq, 0>>>(wpq, iters, d_cyc, d_sink);
    cudaDeviceSynchronize();
  }
  unsigned long long cyc = 0;
  cudaMemcpy(&cyc, d_cyc, 8, cudaMemcpyDeviceToHost);
  cudaError_t e = cudaGetLastError();
  const double warps = 148.0 * 4 * wpq;
  const double per_warp = cyc / warps;                       // cycles one warp spent in its loop
  const double loads = static_cast<double>(iters) * DEPTH;   // loads per warp
  const double bytes_per_sm = loads * N * 128.0 * 4 * wpq;   // 32 lanes x N cols x 4 B per load, 4*wpq warps per SM
  printf("{\"cols_per_ld\": %d, \"depth\": %d, \"warps_per_quarter\": %d, \"cycles_per_ld_per_warp\": %.1f, "
         "\"tmem_read_bytes_per_clk_per_sm\": %.1f, \"err\": \"%s\"}\n",
         N, DEPTH, wpq, per_warp / loads, bytes_per_sm / per_warp, cudaGetErrorString(e));
  cudaFree(d_cyc);
  cudaFree(d_sink);
}

int main() {
  for (int wpq = 1; wpq <= 4; ++wpq) {
    run<16, 1>(wpq);
    run<32, 1>(wpq);
    run<32, 2>(wpq);
    run<64, 1>(wpq);
  }
  run<32, 1>(5);
  return 0;
}
