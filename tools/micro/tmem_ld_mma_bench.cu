// Micro-benchmark: tcgen05.ld latency / throughput WHILE the tensor pipe accumulates into TMEM, and what the
// loads cost the MMAs.  (B200, sm_100a)
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/micro/tmem_ld_mma_bench tools/micro/tmem_ld_mma_bench.cu
// One CTA per SM (512 TMEM columns, one 16 KB A slice + one 32 KB B slice of zeros in shared memory, 128B
// swizzle layout).  Warp 1 issues tcgen05.mma kind::f16 M=128 N=256 K=16 in groups of 4 (one "stage") followed
// by a tcgen05.commit, at most QD stages ahead of the oldest unfinished one (the ring depth of the real
// kernel); 24 MMAs go to one accumulator, then the other.  W epilogue warps per TMEM lane quarter loop over
// tcgen05.ld.32x32b.xN with DEPTH loads in flight, on the accumulator half the MMAs are NOT writing
// (disjoint = 1) or anywhere (disjoint = 0), until the MMA warp is done.
// Prints, per configuration: cycles per MMA instruction (128 = full rate), cycles per load per warp, and the
// bytes / clock / SM the loads achieved.
#include <cstdio>
#include <cuda_runtime.h>
#include <stdint.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t done = 0;
  for (uint32_t spin = 0; !done; ++spin) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(bar), "r"(parity)
        : "memory");
    if (!done && spin > (1u << 26)) __trap();
  }
}
__device__ __forceinline__ void tcgen05_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
      "l"(da), "l"(db), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred P;\n\telect.sync _|P, 0xffffffff;\n\tselp.u32 %0, 1, 0, P;\n\t}" : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t smem_addr) {
  const uint64_t lo = static_cast<uint64_t>((smem_addr & 0x3ffffu) >> 4) | (1ull << 16);
  const uint64_t hi = static_cast<uint64_t>(1024u >> 4) | (1ull << 14) | (2ull << 29);
  return lo | (hi << 32);
}
__host__ __device__ constexpr uint32_t make_idesc(int n) {
  return (1u << 4) | (1u << 7) | (1u << 10) | (static_cast<uint32_t>(n >> 3) << 17) | (static_cast<uint32_t>(128 >> 4) << 24);
}

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

struct Out {
  unsigned long long mma_cycles;   // summed over CTAs
  unsigned long long mma_instrs;
  unsigned long long ld_cycles;    // summed over epilogue warps
  unsigned long long ld_count;
  unsigned long long ld_lat_sum;   // issue -> data for DEPTH == 1 (warp 4 of each CTA only)
  unsigned long long ld_lat_max;
  unsigned sink;
};

// warps 0-3 control, up to 4 epilogue warps per lane quarter (2 when a warp holds 128 data registers)
template <int N, int DEPTH>
constexpr int kThreads = 128 + 128 * (N * DEPTH > 64 ? 2 : 4);

template <int N, int DEPTH>
__global__ void __launch_bounds__((kThreads<N, DEPTH>), 1)
bench(int mma_on, int mma_n, int qd, int n_stages, int wpq, int disjoint, int ld_iters, int chain, int nacc, int randomize,
      int rotate, int lat_on, int uniform, Out* out) {
  extern __shared__ unsigned char smem_raw[];
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t bars[16];
  __shared__ uint32_t tmem_slot;
  __shared__ volatile int mma_done;
  __shared__ volatile int mma_acc;  // accumulator the MMAs are writing
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // operands: zeros, or bf16 values in (-1, 1) from a hash (four 48 KB stage buffers)
  for (int i = threadIdx.x; i < (4 * 48 * 1024) / 4; i += blockDim.x) {
    uint32_t h = static_cast<uint32_t>(i) * 2654435761u + blockIdx.x * 40503u;
    h ^= h >> 15;
    // two bf16: sign + exponent 0x3e/0x3f range -> |x| in [0.125, 1)
    const uint32_t lo = (h & 0x807fu) | 0x3e00u | ((h >> 3) & 0x0100u);
    const uint32_t hi = ((h >> 16) & 0x807fu) | 0x3e00u | ((h >> 19) & 0x0100u);
    reinterpret_cast<uint32_t*>(smem)[i] = randomize ? (lo | (hi << 16)) : 0u;
  }
  if (threadIdx.x == 0) {
    for (int s = 0; s < 16; ++s) mbar_init(smem_u32(&bars[s]), 1);
    mma_done = mma_on ? 0 : 1;
    mma_acc = 0;
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"(512)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t base = tmem_slot;
  const int warp_u = __shfl_sync(0xffffffffu, warp, 0);

  if (warp_u == 1 && uniform) {
    if (mma_on) {
      const uint32_t idesc = make_idesc(mma_n);
      const uint32_t base_u = __shfl_sync(0xffffffffu, base, 0);
      const long long t0 = clock64();
      // (no divisions in this loop: one thread issues everything, its scalar latency is on the critical path)
      int slot = 0;
      uint32_t par = 1;          // parity to wait for on a slot's PREVIOUS use (first lap: passes at once)
      int acc = 0, in_chain = 0, warm = nacc * chain;
      uint32_t stage_off = 0;
      const uint32_t s0 = smem_u32(smem);
      for (int s = 0; s < n_stages; ++s) {
        mbar_wait(smem_u32(&bars[slot]), par);
        const uint64_t da = make_smem_desc(s0 + stage_off);
        const uint64_t db = make_smem_desc(s0 + stage_off + 16384);
        // every lane runs the loop (warp-uniform operands stay in uniform registers); one elected lane issues
        const bool leader = elect_one();
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          if (leader) umma_bf16(base_u + acc * mma_n, da + 2 * j, db + 2 * j, idesc, warm <= 0 ? 1u : 0u);
          --warm;
          if (++in_chain == chain) {
            in_chain = 0;
            if (++acc == nacc) acc = 0;
          }
        }
        if (leader) tcgen05_commit(smem_u32(&bars[slot]));
        __syncwarp();
        if (rotate) stage_off = (stage_off == 3 * 49152) ? 0 : stage_off + 49152;
        if (++slot == qd) {
          slot = 0;
          par ^= 1u;
        }
        if ((s & 7) == 7) mma_acc = acc;
      }
      for (int s = 0; s < qd; ++s) {
        mbar_wait(smem_u32(&bars[slot]), par);
        if (++slot == qd) {
          slot = 0;
          par ^= 1u;
        }
      }
      const long long t1 = clock64();
      if (lane == 0) {
        mma_done = 1;
        atomicAdd(&out->mma_cycles, static_cast<unsigned long long>(t1 - t0));
        atomicAdd(&out->mma_instrs, static_cast<unsigned long long>(n_stages) * 4ull);
      }
    }
  } else if (warp == 1) {
    if (lane == 0 && mma_on) {
      const uint32_t idesc = make_idesc(mma_n);
      const long long t0 = clock64();
      // (no divisions in this loop: one thread issues everything, its scalar latency is on the critical path)
      int slot = 0;
      uint32_t par = 1;          // parity to wait for on a slot's PREVIOUS use (first lap: passes at once)
      int acc = 0, in_chain = 0, warm = nacc * chain;
      uint32_t stage_off = 0;
      const uint32_t s0 = smem_u32(smem);
      for (int s = 0; s < n_stages; ++s) {
        mbar_wait(smem_u32(&bars[slot]), par);
        const uint64_t da = make_smem_desc(s0 + stage_off);
        const uint64_t db = make_smem_desc(s0 + stage_off + 16384);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          // the first `nacc * chain` MMAs overwrite, everything later accumulates (values do not matter)
          umma_bf16(base + acc * mma_n, da + 2 * j, db + 2 * j, idesc, warm <= 0 ? 1u : 0u);
          --warm;
          if (++in_chain == chain) {
            in_chain = 0;
            if (++acc == nacc) acc = 0;
          }
        }
        tcgen05_commit(smem_u32(&bars[slot]));
        if (rotate) stage_off = (stage_off == 3 * 49152) ? 0 : stage_off + 49152;
        if (++slot == qd) {
          slot = 0;
          par ^= 1u;
        }
        if ((s & 7) == 7) mma_acc = acc;
      }
      for (int s = 0; s < qd; ++s) {
        mbar_wait(smem_u32(&bars[slot]), par);
        if (++slot == qd) {
          slot = 0;
          par ^= 1u;
        }
      }
      const long long t1 = clock64();
      mma_done = 1;
      atomicAdd(&out->mma_cycles, static_cast<unsigned long long>(t1 - t0));
      atomicAdd(&out->mma_instrs, static_cast<unsigned long long>(n_stages) * 4ull);
    }
  } else if (warp >= 4) {
    const int ew = warp & 3;
    const int idx = (warp - 4) >> 2;
    if (idx < wpq) {
      uint32_t v[DEPTH][N];
      unsigned sink = 0;
      const uint32_t lane_base = base + (static_cast<uint32_t>(ew * 32) << 16);
      unsigned long long count = 0, lat_sum = 0, lat_max = 0;
      const long long t0 = clock64();
      int other = 0, stop = 0;
      const int span = disjoint ? 256 : 512;
      for (int it = 0; !stop; ++it) {
        if ((it & 7) == 0) {  // look at the MMA warp's state every 8 groups only (keeps the loop about loads)
          stop = mma_on ? mma_done : (it >= ld_iters);
          other = disjoint ? ((mma_acc ^ 1) * 256) : 0;
        }
        long long a0 = 0;
        if (lat_on) a0 = clock64();
#pragma unroll
        for (int d = 0; d < DEPTH; ++d) {
          const uint32_t col = static_cast<uint32_t>(other + ((((it * DEPTH + d) * N) + idx * 64) & (span - 1))) &
                               ~static_cast<uint32_t>(N - 1);
          tmem_ld<N>(lane_base + col, v[d]);
        }
        tmem_wait();
#pragma unroll
        for (int d = 0; d < DEPTH; ++d) sink += v[d][0] ^ v[d][N - 1];
        if (lat_on) {
          const unsigned long long lat = static_cast<unsigned long long>(clock64() - a0);
          lat_sum += lat;
          if (lat > lat_max) lat_max = lat;
        }
        count += DEPTH;
      }
      const long long t1 = clock64();
      if (lane == 0) {
        atomicAdd(&out->ld_cycles, static_cast<unsigned long long>(t1 - t0));
        atomicAdd(&out->ld_count, count);
        atomicAdd(&out->ld_lat_sum, lat_sum);
        atomicMax(&out->ld_lat_max, lat_max);
        if (sink == 0x12345678u) out->sink = sink;
      }
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(base), "r"(512) : "memory");
  }
}

template <int N, int DEPTH>
static void run(int mma_on, int mma_n, int qd, int wpq, int disjoint, int chain = 24, int nacc = 2, int randomize = 0,
                int rotate = 0, int lat_on = 0, int uniform = 1) {
  if (128 + 128 * wpq > kThreads<N, DEPTH>) return;
  Out* d_out;
  cudaMalloc(&d_out, sizeof(Out));
  const int n_stages = 6 * 400;  // 400 tiles of 24 MMAs
  const int ld_iters = 4096;
  const size_t smem = 4 * 48 * 1024 + 1024;
  cudaFuncSetAttribute(bench<N, DEPTH>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
  Out h{};
  for (int rep = 0; rep < 2; ++rep) {
    cudaMemset(d_out, 0, sizeof(Out));
    bench<N, DEPTH><<<148, kThreads<N, DEPTH>, smem>>>(mma_on, mma_n, qd, n_stages, wpq, disjoint, ld_iters, chain, nacc,
                                                        randomize, rotate, lat_on, uniform, d_out);
    cudaDeviceSynchronize();
  }
  cudaMemcpy(&h, d_out, sizeof(Out), cudaMemcpyDeviceToHost);
  cudaError_t e = cudaGetLastError();
  const double cyc_per_mma = h.mma_instrs ? static_cast<double>(h.mma_cycles) / h.mma_instrs : 0.0;
  const double warps = 148.0 * 4 * wpq;
  const double ld_per_warp = wpq ? h.ld_count / warps : 0.0;
  const double cyc_per_warp = wpq ? h.ld_cycles / warps : 0.0;
  const double cyc_per_ld = ld_per_warp > 0 ? cyc_per_warp / ld_per_warp : 0.0;
  const double bytes_clk_sm = cyc_per_warp > 0 ? ld_per_warp * N * 128.0 * 4 * wpq / cyc_per_warp : 0.0;
  printf("{\"chain\": %d, \"nacc\": %d, \"random\": %d, \"rotate\": %d, \"lat_on\": %d, \"uniform_issue\": %d, ", chain, nacc, randomize, rotate, lat_on, uniform);
  printf("\"mma\": %d, \"mma_n\": %d, \"queue_stages\": %d, \"cols_per_ld\": %d, \"depth\": %d, \"warps_per_quarter\": %d, "
         "\"disjoint\": %d, \"cycles_per_mma\": %.1f, \"cycles_per_ld_per_warp\": %.1f, \"group_latency_avg\": %.1f, "
         "\"group_latency_max\": %llu, \"tmem_read_bytes_per_clk_per_sm\": %.1f, \"err\": \"%s\"}\n",
         mma_on, mma_n, qd, N, DEPTH, wpq, disjoint, cyc_per_mma, cyc_per_ld,
         h.ld_count ? static_cast<double>(h.ld_lat_sum) * DEPTH / h.ld_count : 0.0, h.ld_lat_max, bytes_clk_sm,
         cudaGetErrorString(e));
  fflush(stdout);
  cudaFree(d_out);
}

int main(int argc, char** argv) {
  const bool part2 = argc > 1;
  if (!part2) {
    // MMAs alone: N, dependent-chain length, operand values, stage rotation
    for (int n = 64; n <= 256; n *= 2) {
      run<32, 1>(1, n, 4, 0, 1, 24, 2, 0, 0);
      run<32, 1>(1, n, 4, 0, 1, 24, 2, 1, 0);
      run<32, 1>(1, n, 4, 0, 1, 24, 2, 1, 1);
      run<32, 1>(1, n, 4, 0, 1, 4, 2, 1, 1);
      run<32, 1>(1, n, 4, 0, 1, 1, 2, 1, 1);
      run<32, 1>(1, n, 4, 0, 1, 1, 1, 1, 1);
      if (n <= 128) run<32, 1>(1, n, 4, 0, 1, 1, 4, 1, 1);
    }
    run<32, 1>(1, 256, 4, 0, 1, 24, 2, 1, 1, 0, 0);  // issued from a divergent `if (lane == 0)` branch
    run<32, 1>(1, 128, 4, 0, 1, 24, 2, 1, 1, 0, 0);
    run<32, 1>(1, 256, 1, 0, 1, 24, 2, 1, 1);
    run<32, 1>(1, 256, 2, 0, 1, 24, 2, 1, 1);
    run<32, 1>(1, 256, 8, 0, 1, 24, 2, 1, 1);
    // loads alone
    run<32, 1>(0, 256, 4, 1, 1);
    run<32, 1>(0, 256, 4, 2, 1);
    run<32, 2>(0, 256, 4, 2, 1);
    run<64, 1>(0, 256, 4, 2, 1);
    run<32, 4>(0, 256, 4, 2, 1);
    // both (random operands, rotating stages)
    for (int wpq = 1; wpq <= 4; ++wpq) {
      run<32, 1>(1, 256, 4, wpq, 1, 24, 2, 1, 1);
      run<32, 2>(1, 256, 4, wpq, 1, 24, 2, 1, 1);
      run<32, 4>(1, 256, 4, wpq, 1, 24, 2, 1, 1);
      run<64, 1>(1, 256, 4, wpq, 1, 24, 2, 1, 1);
      run<64, 2>(1, 256, 4, wpq, 1, 24, 2, 1, 1);
      run<32, 1>(1, 256, 4, wpq, 0, 24, 2, 1, 1);
      run<32, 2>(1, 256, 4, wpq, 1, 1, 2, 1, 1);
      run<32, 1>(1, 256, 4, wpq, 1, 24, 2, 1, 1, 1);
    }
  }
  return 0;
}
