// Batched queries: S = Qn . V^T on the 5th-generation tensor cores with a fused mask + top-k
// epilogue, then an fp32 re-scoring pass.  Replaces the batched form of
//     scores = vecs @ V.T ; argpartition ; argsort        (picovdb/pico_vdb.py:683-714)
// without ever writing the Q x N score matrix to HBM.
//
// Structure (one persistent CTA per SM, 384 threads, warp-specialised):
//   warp 0      TMA producer: cp.async.bulk.tensor loads a 128-row query tile slice (A) and a
//               256-row database tile slice (B), 128 bytes of K each, into a 4-stage shared-memory
//               ring (128B swizzle), completion on mbarriers.
//   warp 1      MMA issuer: one elected lane issues tcgen05.mma (kind::tf32 on the fp32 matrix or
//               kind::f16 on the bf16 mirror), M=128 x N=256, accumulating in TMEM; tcgen05.commit
//               releases ring slots and publishes finished accumulators.  TMEM holds two
//               accumulators (2 x 256 columns) so the epilogue of tile i overlaps the MMAs of i+1.
//   warps 4-11  epilogue: two warps per TMEM lane quarter, one per 128-column HALF of the
//               accumulator (the tcgen05.ld round trip, not instruction issue, bounds one warp; two
//               warps per SM sub-partition keep two loads in flight).  Thread t of a warp owns query
//               t of the tile (TMEM lane t) for its half: it streams its 128 scores out of TMEM
//               (tcgen05.ld 32x32b.x32, next chunk in flight while the current one is examined),
//               turns masked columns into -inf, and compares the MAXIMUM of each 8-column group
//               with the query's running threshold; only groups that beat it are examined column by
//               column and appended to the (query, half) candidate pool.  When a pool fills, the
//               warp sorts it co-operatively (bitonic network in registers), keeps the best k_sel
//               and raises the threshold, which both halves share through shared memory.
//
// Schedule: a "visit" is (database tile t, query tile qt); a WORK ITEM is a block of R <= 8 consecutive
// database tiles for one query tile (pair, with 2-CTA clusters), items are numbered tile-block-major and
// unit b (CTA or cluster) takes items b, b + units, b + 2 units, ... (VisitSeq below).  At any moment the
// 148 CTAs therefore work inside a window of a few tile blocks, each database tile is fetched from HBM once
// and then served to the other query tiles from L2 (the first version walked one query tile down a long
// chunk of database tiles per CTA; CTAs drifted apart and the 10M x 768 case re-read the database 24x from
// HBM), and inside an item the epilogue keeps the query tile's thresholds / counters / pool pointers in
// registers.  A CTA keeps one running (threshold, count, pool) per query tile it meets; at the end it leaves
// each pool as it is (unsorted, with its count).
// Both control warps run their loops with ALL lanes and let one elected lane issue (elect_one()): issued
// from a divergent single-lane branch every tcgen05.mma cost 168 cycles of operand shuffling in SASS and
// the tensor pipe (128 cycles per 128 x 256 x 16 MMA) was paced by instruction issue.
// finalize_batch_kernel streams those pools per query through a threshold filter (the best published
// k_sel-th score is a proven lower bound), sorts what survives, re-scores the best exactly in fp32
// against the fp32 matrix (tensor-core inputs are rounded to tf32 / bf16; the north star asks for
// 1e-5 fp32 scores) and writes the top k.
//
// Passes of one search (search_batch): (1) seed -- the first <= 32 tiles are multiplied as a plain
// GEMM with the scores written out, a radix select gives each query's k_sel-th best as a starting
// threshold; (2) sample (large stores) -- 1/16 of the tiles is searched and merged, its exact k_sel-th
// best per query seeds (3) the main pass over the remaining tiles; both merges carry their lists
// into the final top k.  Every bound used to drop a candidate is a proven lower bound of the
// query's final k_sel-th best, so the result is exact regardless of timing.
//
// Tensor-core roofline: algorithmic flops = 2 * Q * N * dim per batch.
#include <cuda.h>

#include <algorithm>
#include <cstdlib>
#include <mutex>
#include <utility>
#include <vector>

#include "batch.cuh"
#include "store.cuh"

namespace pvdb {

// ---------------------------------------------------------------------------- tile constants
constexpr int kBM = 128;            // queries per tile (UMMA M, TMEM lanes)
constexpr int kBN = 256;            // database rows per tile (UMMA N, TMEM columns per accumulator)
constexpr int kKBytes = 128;        // K bytes per pipeline stage == one 128B swizzle atom
constexpr int kStages = 4;
constexpr int kStageABytes = kBM * kKBytes;  // 16 KB
constexpr int kStageBBytes = kBN * kKBytes;  // 32 KB
constexpr int kStageBytes = kStageABytes + kStageBBytes;
constexpr int kEpiHalves = 2;       // epilogue warps per TMEM lane quarter (each owns kBN / 2 columns)
constexpr int kEpiThreads = 128 * kEpiHalves;
constexpr int kBatchThreads = 128 + kEpiThreads;
constexpr int kTmemCols = 512;      // two fp32 accumulators of 256 columns
constexpr int kMaxSel = 160;        // largest k_sel (a pool of 256 keeps two 32-column chunks of headroom + 32 slots)
// Extra candidates kept for the exact re-scoring.  The exactness guard (finalize) accepts a query only
// when the re-scored k-th best beats the weakest kept candidate by more than the input-rounding bound,
// so the slack must span that bound in score space.  The spacing of scores at rank r shrinks like
// 1 / r: small k gets by with 32 (tf32) / 54 (bf16); large k takes what the pool allows (k_sel <= 160).
constexpr int kSlackTF32 = 32;
constexpr int kSlackBF16 = 54;
constexpr int kSlackLargeK = 60;
constexpr int kSampleFraction = 16; // the sample pass covers 1/16 of the database tiles
// The pair (cta_group::2) variant is validated (all batch tests pass with PVDB_BATCH_PAIR=1).  It measured 15 %
// slower than 2-CTA clusters with cta_group::1 MMAs + TMA multicast until its accumulator hand-back stopped
// issuing a cluster-wide memory barrier per visit (mbar_arrive_cluster); since then it is within +-2 % of
// the default (12.5M x 384 bf16: 30.6-31.3 vs 30.0-30.4 ms), so it stays opt-in.
constexpr bool kPairDefault = false;
constexpr int kClusterDefault = 2;  // CTAs per cluster sharing a database tile; 4 and 8 work but measured 3 % / 10 % slower
constexpr int kMaxQTiles = 32;      // query tiles per launch (4096 queries); larger batches are split
constexpr int kMaxTileBlock = 8;    // consecutive database tiles a unit gives to ONE query tile (group) -- see VisitSeq
// shared memory: [ring][barriers + tmem slot (256 B)][thr: 32 x 128 ordered u32][cnt: 2 x 32 x 128 u16][touched: 32 B]
constexpr size_t kRingBytes = static_cast<size_t>(kStages) * kStageBytes;
constexpr size_t kStateBytes =
    static_cast<size_t>(kMaxQTiles) * kBM * (sizeof(uint32_t) + kEpiHalves * sizeof(uint16_t)) + kMaxQTiles;
constexpr size_t kBatchSmem = 1024 /*align*/ + kRingBytes + 256 + kStateBytes;
static_assert(kBatchSmem <= 227 * 1024, "the batch kernel's shared memory must fit one SM");

struct BatchParams {
  int64_t nq;            // queries in this launch (<= 4096)
  int64_t n_rows;        // database rows (high-water mark)
  int k_blocks;          // ceil(dim * elem / 128)
  int q_tiles;           // ceil(nq / 128)
  int n_tiles;           // ceil(n_rows / 256)
  int k_sel;             // candidates kept per (CTA, query)
  int pool_cap;          // 64 / 128 / 256 keys per (CTA, query tile, query)
  const uint32_t* active;
  const uint32_t* prefilter;
  uint64_t* pools;       // [grid][q_tiles][half][128][pool_cap]
  uint16_t* counts;      // [grid][q_tiles][half][128]: keys left in each pool (written for touched tiles)
  uint8_t* touched;      // [grid][q_tiles]: CTA b met query tile qt (zeroed before the launch)
  uint32_t* shared_thr;  // [nq] ordered-int image of the best published k_sel-th score (zeroed)
  int tile_begin;        // first database tile of this launch (n_tiles counts from here)
  int tile_block;        // R: database tiles per work item (VisitSeq)
  int visit_stride;      // 0: units stride over the work items by their count (every unit meets many
                         // query tiles); else a multiple of the query-tile(-pair) count: unit u keeps ONE
                         // query tile and units >= visit_stride stay idle (cheap cold start, sample pass)
  const float* init_thr; // optional [nq]: a proven lower bound of each query's k_sel-th best score
  float* dump;           // seed pass only: [q_tiles * 128][dump_ld] masked tensor-core scores are written
  int dump_ld;           //   here (column = row - tile_begin * 256) instead of being selected
};

// Work items and visits.  A "visit" is (database tile t, query tile qt).  A WORK ITEM is a block of R
// consecutive database tiles for one query-tile group (one query tile; with clusters of CL CTAs: CL
// consecutive query tiles, CTA c of the cluster taking the c-th): item = tile_block * n_groups + group,
// i.e. items are numbered tile-block-major, and unit u takes items u, u + step, u + 2 step, ...  At any
// moment the units therefore work inside a window of a few tile blocks (each database tile comes from HBM
// once and is served to the other query tiles from L2), while inside an item the query tile does not change:
// the epilogue keeps its thresholds, counters and pool pointers in registers for R visits instead of going
// through shared memory and a 64-bit division per visit (R = 1 is the round-1 visit order).  The TMA, MMA
// and epilogue warps all walk the same sequence with this generator; nothing in it divides after start().
template <int CL>
struct VisitSeq {
  int n_groups, n_tiles, R;
  int step_tb, step_g;   // item step split into (tile blocks, groups)
  int64_t item, n_items, step;
  int tb, g;             // current item
  int r, nr;             // visit inside the item, visits in the item
  __device__ __forceinline__ void start(const BatchParams& p, int64_t first_item, int64_t item_step) {
    n_groups = (p.q_tiles + CL - 1) / CL;
    n_tiles = p.n_tiles;
    R = p.tile_block;
    const int n_tb = (n_tiles + R - 1) / R;
    n_items = static_cast<int64_t>(n_tb) * n_groups;
    step = item_step;
    step_tb = static_cast<int>(item_step / n_groups);
    step_g = static_cast<int>(item_step - static_cast<int64_t>(step_tb) * n_groups);
    item = first_item < n_items ? first_item : n_items;
    tb = static_cast<int>(item / n_groups);
    g = static_cast<int>(item - static_cast<int64_t>(tb) * n_groups);
    r = 0;
    nr = min(R, n_tiles - tb * R);
  }
  __device__ __forceinline__ bool done() const { return item >= n_items; }
  __device__ __forceinline__ int t() const { return tb * R + r; }
  __device__ __forceinline__ int qt(uint32_t cta_rank) const { return g * CL + static_cast<int>(cta_rank); }
  __device__ __forceinline__ bool first_in_item() const { return r == 0; }
  __device__ __forceinline__ bool last_in_item() const { return r + 1 == nr; }
  __device__ __forceinline__ void next() {
    if (++r < nr) return;
    item += step;
    tb += step_tb;
    g += step_g;
    if (g >= n_groups) {
      g -= n_groups;
      ++tb;
    }
    r = 0;
    nr = min(R, n_tiles - tb * R);
  }
};

// ---------------------------------------------------------------------------- PTX wrappers
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
// Spin on the phase parity.  A broken pipeline would otherwise hang the GPU until the watchdog;
// after ~2^28 polls the kernel traps so the failure surfaces as a CUDA error instead.
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t done = 0;
  for (uint32_t spin = 0; !done; ++spin) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t"
        "}"
        : "=r"(done)
        : "r"(bar), "r"(parity)
        : "memory");
    if (!done && spin > (1u << 28)) __trap();
  }
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(bar), "r"(c0), "r"(c1)
      : "memory");
}
// multicast variant: the box lands at the same CTA-relative offset in every CTA of `mask` and
// completes bytes on the mbarrier at the same offset in each of them
__device__ __forceinline__ void tma_load_2d_mc(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1,
                                               uint16_t mask) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster"
      " [%0], [%1, {%3, %4}], [%2], %5;"
      ::"r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(bar), "r"(c0), "r"(c1), "h"(mask)
      : "memory");
}
__device__ __forceinline__ void tcgen05_commit_mc(uint32_t bar, uint16_t mask) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
               ::"r"(bar), "h"(mask)
               : "memory");
}
// ---- cta_group::2 (CTA pair) forms.  Shared-memory addresses of the two CTAs of a pair differ in
// one bit of the shared::cluster window; clearing it addresses the even (leader) CTA.
constexpr uint32_t kPeerBitMask = 0xFEFFFFFFu;
__device__ __forceinline__ void tma_load_2d_pair(uint32_t dst, const CUtensorMap* map, uint32_t leader_bar, int c0,
                                                 int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4}], [%2];"
      ::"r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(leader_bar), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tcgen05_commit_pair(uint32_t bar, uint16_t mask) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
               ::"r"(bar), "h"(mask)
               : "memory");
}
// Arrival on the PEER CTA's barrier (pair variant: the epilogue hands an accumulator back to the leader's MMA
// warp).  Relaxed: what the arrival orders are this warp's tcgen05.ld reads, which have completed
// (tcgen05.wait::ld) and are fenced by tcgen05.fence::before_thread_sync; a .release.cluster arrival compiled to
// a cluster-wide memory barrier that waited for the warp's outstanding pool stores on every visit (13 % of the
// pair variant's stall samples).
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.relaxed.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
template <bool BF16>
__device__ __forceinline__ void umma_pair(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                          uint32_t accumulate) {
  if constexpr (BF16) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t"
        "}" ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
        : "memory");
  } else {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::tf32 [%0], %1, %2, %3, p;\n\t"
        "}" ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
        : "memory");
  }
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// One lane of a CONVERGED warp.  The TMA and MMA warps run their loops with all 32 lanes and let the elected
// lane issue: operands that are warp-uniform then stay in uniform registers and UTMALDG / UTCHMMA are issued
// back to back.  Issued from a divergent `if (lane == 0)` branch instead, every tcgen05.mma became an
// ELECT + 5 x R2UR.BROADCAST + branch "waterfall" loop in SASS, which a micro-benchmark
// (tools/micro/tmem_ld_mma_bench.cu) timed at 168 cycles per instruction whatever its N -- slower than the
// 128 cycles a 128 x 256 x 16 MMA needs, i.e. the tensor pipe was paced by instruction issue.
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile(
      "{\n\t"
      ".reg .pred P;\n\t"
      "elect.sync _|P, 0xffffffff;\n\t"
      "selp.u32 %0, 1, 0, P;\n\t"
      "}"
      : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ void tcgen05_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tcgen05_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tcgen05_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
template <bool BF16>
__device__ __forceinline__ void umma(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                     uint32_t accumulate) {
  if constexpr (BF16) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
        "}" ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
        : "memory");
  } else {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t"
        "}" ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
        : "memory");
  }
}
// 32 lanes x 32 consecutive fp32 columns of this warp's TMEM lane quarter (asynchronous: the
// registers are valid after tmem_ld_wait)
__device__ __forceinline__ void tmem_ld_32x32(uint32_t taddr, uint32_t (&v)[32]) {
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
// Wait for outstanding tcgen05.ld.  The registers are passed as in/out operands so the compiler
// cannot move any use of them above the wait.
__device__ __forceinline__ void tmem_ld_wait(uint32_t (&v)[32]) {
  asm volatile("tcgen05.wait::ld.sync.aligned;"
               : "+r"(v[0]), "+r"(v[1]), "+r"(v[2]), "+r"(v[3]), "+r"(v[4]), "+r"(v[5]), "+r"(v[6]), "+r"(v[7]),
                 "+r"(v[8]), "+r"(v[9]), "+r"(v[10]), "+r"(v[11]), "+r"(v[12]), "+r"(v[13]), "+r"(v[14]), "+r"(v[15]),
                 "+r"(v[16]), "+r"(v[17]), "+r"(v[18]), "+r"(v[19]), "+r"(v[20]), "+r"(v[21]), "+r"(v[22]), "+r"(v[23]),
                 "+r"(v[24]), "+r"(v[25]), "+r"(v[26]), "+r"(v[27]), "+r"(v[28]), "+r"(v[29]), "+r"(v[30]), "+r"(v[31])
               :
               : "memory");
}

// shared-memory matrix descriptor: K-major tile, 128-byte swizzle, rows 128 B apart, 8-row groups
// 1024 B apart (SBO), descriptor version 1 (sm_100)
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t smem_addr) {
  const uint64_t lo = static_cast<uint64_t>((smem_addr & 0x3ffffu) >> 4) | (1ull << 16);
  const uint64_t hi = static_cast<uint64_t>(1024u >> 4) | (1ull << 14) | (2ull << 29);
  return lo | (hi << 32);
}
// instruction descriptor: D=f32, A/B = tf32 (2) or bf16 (1), both K-major, N=256, M=128
__host__ __device__ constexpr uint32_t make_idesc(bool bf16, int m = kBM) {
  const uint32_t fmt = bf16 ? 1u : 2u;
  return (1u << 4) | (fmt << 7) | (fmt << 10) | (static_cast<uint32_t>(kBN >> 3) << 17) |
         (static_cast<uint32_t>(m >> 4) << 24);
}

// ---------------------------------------------------------------------------- warp bitonic sort
// 32*NI keys, element e = i*32 + lane, sorted descending.  The cross-lane stages run as ROLLED loops
// over (size, stride) with the NI keys of a lane statically indexed: the fully unrolled network was
// ~3000 instructions (48 KB) of straight-line code that missed the instruction cache on every prune
// (measured 11-16k cycles per prune of 256 keys).
template <int NI>
__device__ __forceinline__ void warp_cross_stage(uint64_t (&key)[NI], int lane, int size, int stride) {
  const bool lower = (lane & stride) == 0;
#pragma unroll
  for (int i = 0; i < NI; ++i) {
    const uint64_t other = shfl_xor_u64(key[i], stride);
    const bool desc = ((i * 32 + lane) & size) == 0;
    // the lower lane of a descending pair keeps the larger key
    const bool take_max = (desc == lower);
    if ((key[i] < other) == take_max) key[i] = other;
  }
}

template <int NI>
__device__ __forceinline__ void warp_sort_desc(uint64_t (&key)[NI], int lane) {
  // sizes 2..32: only cross-lane stages
#pragma unroll 1
  for (int size = 2; size <= 32; size <<= 1) {
#pragma unroll 1
    for (int stride = size >> 1; stride > 0; stride >>= 1) warp_cross_stage<NI>(key, lane, size, stride);
  }
  // sizes 64..32*NI: a few in-register stages (static indices), then the five cross-lane stages
#pragma unroll
  for (int size = 64; size <= 32 * NI; size <<= 1) {
#pragma unroll
    for (int stride = size >> 1; stride >= 32; stride >>= 1) {
      const int sj = stride >> 5;
#pragma unroll
      for (int i = 0; i < NI; ++i) {
        if ((i & sj) == 0) {
          const bool desc = ((i * 32) & size) == 0;
          const uint64_t a = key[i], b = key[i | sj];
          if ((a < b) == desc) {
            key[i] = b;
            key[i | sj] = a;
          }
        }
      }
    }
#pragma unroll 1
    for (int stride = 16; stride > 0; stride >>= 1) warp_cross_stage<NI>(key, lane, size, stride);
  }
}

// Pools of the 32 queries of a warp are interleaved sector by sector: slot s of lane l lives at
// key index ((s / 4) * 32 + l) * 4 + s % 4 of the warp's block of 32 * pool_cap keys.  While a tile
// is examined with no threshold yet (cold start) all lanes append in lock step and a warp-wide
// append touches 8 consecutive 128-byte lines instead of 32 scattered ones; the readers (prune,
// finalize) still get whole 32-byte sectors of one query.
__device__ __forceinline__ size_t pool_slot(int lane, int s) {
  return (static_cast<size_t>(s >> 2) * 32 + lane) * 4 + (s & 3);
}

// Co-operative prune of the pool of lane `src`: keep the best k_sel keys (sorted, at the front).
// Returns (through the references) the pool's new count and threshold.
template <int NI>
__device__ __forceinline__ void prune_pool(uint64_t* wpool, int src, int count, int k_sel, int lane, int& new_count,
                                           float& new_thr) {
  uint64_t key[NI];
#pragma unroll
  for (int i = 0; i < NI; ++i) {
    const int e = i * 32 + lane;
    key[i] = (e < count) ? wpool[pool_slot(src, e)] : 0ull;
  }
  warp_sort_desc<NI>(key, lane);
#pragma unroll
  for (int i = 0; i < NI; ++i) {
    const int e = i * 32 + lane;
    if (e < k_sel) wpool[pool_slot(src, e)] = key[i];
  }
  // k-th key (entry k_sel-1) decides the new threshold
  const int ke = k_sel - 1;
  uint64_t kv = key[0];
#pragma unroll
  for (int i = 1; i < NI; ++i)
    if ((ke >> 5) == i) kv = key[i];
  kv = shfl_u64(kv, ke & 31);
  new_count = count < k_sel ? count : k_sel;
  new_thr = (count >= k_sel && kv != 0ull) ? key_score(kv) : -INFINITY;
}

// One 32-column chunk of scores for this thread's query: masked columns become -inf, then only
// 8-column groups whose maximum beats the threshold are examined column by column.
__device__ __forceinline__ void scan_chunk(uint32_t (&v)[32], uint32_t mw, float thr, uint64_t* wpool, int lane,
                                           int& cnt, uint32_t row_base) {
  if (mw != 0xffffffffu) {
#pragma unroll
    for (int j = 0; j < 32; ++j)
      if (!((mw >> j) & 1u)) v[j] = 0xff800000u;  // -inf
  }
  float g[4];
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    float m = __uint_as_float(v[q * 8]);
#pragma unroll
    for (int j = 1; j < 8; ++j) m = fmaxf(m, __uint_as_float(v[q * 8 + j]));
    g[q] = m;
  }
  const float m_all = fmaxf(fmaxf(g[0], g[1]), fmaxf(g[2], g[3]));
  if (m_all > thr) {
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      if (g[q] > thr) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float sc = __uint_as_float(v[q * 8 + j]);
          if (sc > thr) {
            wpool[pool_slot(lane, cnt)] = make_key(sc, row_base + q * 8 + j);
            ++cnt;
          }
        }
      }
    }
  }
}

// ---------------------------------------------------------------------------- optional cycle counters
// -DPVDB_BATCH_STATS (PVDB_NVCC_EXTRA): where the epilogue and MMA warps spend their time; read with
// pvdb_debug_batch_stats().  Not compiled into the shipped library.
#ifdef PVDB_BATCH_STATS
__device__ unsigned long long g_batch_stats[16];
#define STAT_T(var) const long long var = clock64()
#define STAT_ADD(slot, val) stat_local[slot] += static_cast<unsigned long long>(val)
#else
#define STAT_T(var)
#define STAT_ADD(slot, val)
#endif

// ---------------------------------------------------------------------------- the GEMM + top-k kernel
// PAIR (implies CL == 2): the two CTAs of a cluster form one cta_group::2 MMA unit -- M = 256 (each
// CTA's own 128-query tile), N = 256 with each CTA holding HALF of the database tile's rows.  A stage
// is then 16 KB + 16 KB per CTA, so the same 192 KB ring holds 6 stages instead of 4 (deeper
// pipeline against L2 latency) and each CTA pulls 32 KB instead of 48 KB per K block.  The leader
// (even) CTA issues the MMAs and owns the full / tmem_empty barriers; commits are multicast.
template <bool BF16, int NI, int CL, bool PAIR>
__global__ void __launch_bounds__(kBatchThreads, 1)
batch_topk_kernel(const __grid_constant__ CUtensorMap map_q, const __grid_constant__ CUtensorMap map_db,
                  const BatchParams p) {
  extern __shared__ unsigned char smem_raw[];
  // 1024-byte alignment is required by the 128B swizzle atoms
  unsigned char* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  unsigned char* stage_base = smem;
  static_assert(!PAIR || CL == 2, "a CTA pair is a cluster of two");
  constexpr int NS = PAIR ? 6 : kStages;                                        // ring stages
  constexpr int kStageB = PAIR ? kStageBBytes / 2 : kStageBBytes;               // B bytes per stage in this CTA
  constexpr int kStageAll = kStageABytes + kStageB;
  static_assert(static_cast<size_t>(NS) * kStageAll == kRingBytes, "both layouts use the same ring");
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kRingBytes);
  // bars[0..S) full, [S..2S) empty, [2S..2S+2) tmem_full, [2S+2..2S+4) tmem_empty
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * NS + 4);
  uint32_t* s_thr = reinterpret_cast<uint32_t*>(smem + kRingBytes + 256);              // [q_tiles][128] ordered
  uint16_t* s_cnt = reinterpret_cast<uint16_t*>(s_thr + kMaxQTiles * kBM);            // [half][q_tiles][128]
  uint8_t* s_touched = reinterpret_cast<uint8_t*>(s_cnt + kEpiHalves * kMaxQTiles * kBM);  // [q_tiles]

  const int warp = __shfl_sync(0xffffffffu, static_cast<int>(threadIdx.x >> 5), 0);  // warp-uniform for the compiler
  const int lane = threadIdx.x & 31;
  uint32_t cta_rank = 0;
  if constexpr (CL > 1) asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(cta_rank));
  const int unit_id = static_cast<int>(blockIdx.x) / CL;   // cluster index (== CTA index when CL == 1)
  const int n_units = static_cast<int>(gridDim.x) / CL;
  const int64_t i_step = p.visit_stride > 0 ? p.visit_stride : n_units;   // work items between two of this unit's
  constexpr uint16_t kClusterMask = static_cast<uint16_t>((1u << CL) - 1u);
  const uint32_t bar0 = smem_u32(bars);
  auto full_bar = [&](int s) { return bar0 + 8u * s; };
  auto empty_bar = [&](int s) { return bar0 + 8u * (NS + s); };
  auto tfull_bar = [&](int a) { return bar0 + 8u * (2 * NS + a); };
  auto tempty_bar = [&](int a) { return bar0 + 8u * (2 * NS + 2 + a); };
  const bool is_leader = !PAIR || cta_rank == 0;

  if (threadIdx.x == 0) {
    for (int s = 0; s < NS; ++s) {
      mbar_init(full_bar(s), 1);
      // multicast variant: every CTA that receives the multicast must release the slot;
      // pair variant: one multicast commit from the leader's MMA thread releases it in both CTAs
      mbar_init(empty_bar(s), PAIR ? 1 : CL);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(tfull_bar(a), 1);
      // one arrival per epilogue warp that reads the accumulator (of both CTAs for a pair)
      mbar_init(tempty_bar(a), (PAIR ? 2 : 1) * 4 * kEpiHalves);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (threadIdx.x < kMaxQTiles) s_touched[threadIdx.x] = 0;
  if (warp == 1) {
    if constexpr (PAIR) {
      asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                   "r"(kTmemCols)
                   : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    } else {
      asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                   "r"(kTmemCols)
                   : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
  }
  tcgen05_fence_before();
  if constexpr (CL > 1) cluster_sync_all();  // peers' barriers are initialised before anyone signals them
  else __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_slot, 0);

  // surplus units of a pinned schedule idle (start() clamps a first item beyond the end to "done")
  const int64_t i_first = unit_id < i_step ? unit_id : (int64_t(1) << 60);

  if (warp == 0) {
    // ======================= TMA producer (all lanes loop, one elected lane issues) =======================
    {
      if (lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&map_q)) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&map_db)) : "memory");
      }
      __syncwarp();
      int stage = 0;
      uint32_t phase = 0;
      constexpr int kElemsPerStage = BF16 ? 64 : 32;
#ifdef PVDB_BATCH_STATS
      unsigned long long stat_local[16] = {};
      STAT_T(p_begin);
#endif
      VisitSeq<CL> seq;
      for (seq.start(p, i_first, i_step); !seq.done(); seq.next()) {
        const int t = seq.t(), qt = seq.qt(cta_rank);  // a padding query tile loads zeros (out of bounds)
        for (int kb = 0; kb < p.k_blocks; ++kb) {
          STAT_T(p0);
          mbar_wait(empty_bar(stage), phase ^ 1u);
          STAT_ADD(13, clock64() - p0);
          const uint32_t a_dst = smem_u32(stage_base + static_cast<size_t>(stage) * kStageAll);
          const uint32_t b_dst = a_dst + kStageABytes;
          const bool issuer = elect_one();
          if constexpr (PAIR) {
            // both CTAs load their own query slice and their half of the tile's rows into their own
            // shared memory; all bytes are counted on the LEADER's full barrier
            const uint32_t lbar = full_bar(stage) & kPeerBitMask;
            if (issuer) {
              if (is_leader) mbar_expect_tx(full_bar(stage), 2 * kStageAll);
              tma_load_2d_pair(a_dst, &map_q, lbar, kb * kElemsPerStage, qt * kBM);
              tma_load_2d_pair(b_dst, &map_db, lbar, kb * kElemsPerStage,
                               (p.tile_begin + t) * kBN + static_cast<int>(cta_rank) * (kBN / 2));
            }
            __syncwarp();
            if (++stage == NS) {
              stage = 0;
              phase ^= 1u;
            }
            continue;
          }
          if (issuer) {
            mbar_expect_tx(full_bar(stage), kStageAll);
            tma_load_2d(a_dst, &map_q, full_bar(stage), kb * kElemsPerStage, qt * kBM);
            if constexpr (CL == 1) {
              tma_load_2d(b_dst, &map_db, full_bar(stage), kb * kElemsPerStage, (p.tile_begin + t) * kBN);
            } else {
              // this CTA fetches its 1/CL share of the tile's rows and multicasts it to the cluster
              constexpr int kShareRows = kBN / CL;
              tma_load_2d_mc(b_dst + cta_rank * (kStageBBytes / CL), &map_db, full_bar(stage), kb * kElemsPerStage,
                             (p.tile_begin + t) * kBN + static_cast<int>(cta_rank) * kShareRows, kClusterMask);
            }
          }
          __syncwarp();
          if (++stage == NS) {
            stage = 0;
            phase ^= 1u;
          }
        }
      }
#ifdef PVDB_BATCH_STATS
      STAT_ADD(14, clock64() - p_begin);
      if (lane == 0) {
        atomicAdd(&g_batch_stats[13], stat_local[13]);
        atomicAdd(&g_batch_stats[14], stat_local[14]);
      }
#endif
    }
  } else if (warp == 1) {
    // ======================= MMA issuer (all lanes loop, one elected lane issues) =======================
    if (is_leader) {
      constexpr uint32_t idesc = make_idesc(BF16, PAIR ? 2 * kBM : kBM);
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
#ifdef PVDB_BATCH_STATS
      unsigned long long stat_local[16] = {};
      STAT_T(m_begin);
#endif
      VisitSeq<CL> seq;
      for (seq.start(p, i_first, i_step); !seq.done(); seq.next()) {
        STAT_T(m0);
        mbar_wait(tempty_bar(acc), acc_phase ^ 1u);  // epilogue has drained this accumulator
        STAT_ADD(6, clock64() - m0);
        tcgen05_fence_after();
        const uint32_t tmem_d = tmem_base + static_cast<uint32_t>(acc * kBN);
        for (int kb = 0; kb < p.k_blocks; ++kb) {
          STAT_T(m1);
          mbar_wait(full_bar(stage), phase);
          STAT_ADD(7, clock64() - m1);
          tcgen05_fence_after();
          const uint32_t a_addr = smem_u32(stage_base + static_cast<size_t>(stage) * kStageAll);
          const uint64_t da = make_smem_desc(a_addr);
          const uint64_t db = make_smem_desc(a_addr + kStageABytes);
          const bool last_kb = kb + 1 == p.k_blocks;
          if (elect_one()) {
#pragma unroll
            for (int j = 0; j < kKBytes / 32; ++j) {
              // advance 32 bytes of K inside the swizzle atom: +2 in the (addr >> 4) field
              if constexpr (PAIR)
                umma_pair<BF16>(tmem_d, da + static_cast<uint64_t>(2 * j), db + static_cast<uint64_t>(2 * j), idesc,
                                (kb | j) != 0 ? 1u : 0u);
              else
                umma<BF16>(tmem_d, da + static_cast<uint64_t>(2 * j), db + static_cast<uint64_t>(2 * j), idesc,
                           (kb | j) != 0 ? 1u : 0u);
            }
            // ring slot reusable once these MMAs retire (in every CTA the multicast writes to)
            if constexpr (PAIR) tcgen05_commit_pair(empty_bar(stage), kClusterMask);
            else if constexpr (CL == 1) tcgen05_commit(empty_bar(stage));
            else tcgen05_commit_mc(empty_bar(stage), kClusterMask);
            if (last_kb) {
              // accumulator complete (in both CTAs' tensor memory for a pair)
              if constexpr (PAIR) tcgen05_commit_pair(tfull_bar(acc), kClusterMask);
              else tcgen05_commit(tfull_bar(acc));
            }
          }
          __syncwarp();
          if (++stage == NS) {
            stage = 0;
            phase ^= 1u;
          }
        }
        if (++acc == 2) {
          acc = 0;
          acc_phase ^= 1u;
        }
      }
#ifdef PVDB_BATCH_STATS
      STAT_ADD(8, clock64() - m_begin);
      if (lane == 0)
        for (int i = 6; i <= 8; ++i) atomicAdd(&g_batch_stats[i], stat_local[i]);
#endif
    }
  } else if (warp >= 4) {
    // ======================= epilogue: mask + running top-k =======================
    // Two groups of four warps (one warp per TMEM lane quarter in each); group g works on its 128-column
    // half of every accumulator.
    const int ew = warp & 3;                 // the TMEM lane quarter this warp may read
    const int grp = (warp - 4) >> 2;         // epilogue group
    const int ql = ew * 32 + lane;           // query (TMEM lane) owned by this thread
    constexpr int kCols = kBN / kEpiHalves;  // columns one warp examines per visit
    constexpr int kChunks = kCols / 32;      // 32-column chunks per visit and warp
    const int col0 = grp * kCols;            // first column inside the accumulator
    uint16_t* my_cnt = s_cnt + grp * (kMaxQTiles * kBM);
    for (int qt = 0; qt < p.q_tiles; ++qt) {
      if (grp == 0) {
        const bool live = (static_cast<int64_t>(qt) * kBM + ql) < p.nq;
        uint32_t t0 = f32_to_ordered(-INFINITY);
        if (live && p.init_thr != nullptr) {
          // scores equal to the bound must still be admitted: start one ulp below it
          const float b = p.init_thr[static_cast<int64_t>(qt) * kBM + ql];
          if (b > -INFINITY) t0 = f32_to_ordered(b) - 1u;
        }
        s_thr[qt * kBM + ql] = live ? t0 : f32_to_ordered(INFINITY);  // padding queries never collect candidates
      }
      my_cnt[qt * kBM + ql] = 0;
    }
    asm volatile("bar.sync 1, %0;" ::"n"(kEpiThreads) : "memory");  // both groups see the initial thresholds
    uint64_t* cta_pools = p.pools + static_cast<size_t>(blockIdx.x) * p.q_tiles * (kEpiHalves * kBM) * p.pool_cap;
    int acc = 0;
    uint32_t acc_phase = 0;
#ifdef PVDB_BATCH_STATS
    unsigned long long stat_local[16] = {};
    STAT_T(e_begin);
#endif
    // The mask words of a visit's columns (one word per 32-column chunk; lane c < kChunks holds the word of
    // chunk c; the active bitmap is allocated up to capacity, the prefilter only has ceil(rows / 32) words)
    // are fetched ONE VISIT AHEAD: when the epilogue is the pacing stage the accumulator is already full
    // when a warp arrives, and a ~700-cycle L2 round trip would sit on the critical path of every tile.
    auto load_mask = [&](int t) -> uint32_t {
      uint32_t mw = 0u;
      if (lane < kChunks) {
        const int64_t w = ((static_cast<int64_t>(p.tile_begin + t) * kBN + col0) >> 5) + lane;
        mw = __ldg(p.active + w);
        if (p.prefilter != nullptr) mw &= (w < ((p.n_rows + 31) >> 5)) ? __ldg(p.prefilter + w) : 0u;
      }
      return mw;
    };
    VisitSeq<CL> seq;
    seq.start(p, i_first, i_step);
    uint32_t mw_next = seq.done() ? 0u : load_mask(seq.t());
    // state of the work item (one query tile for up to R visits), in registers
    bool ok = false;              // not a padding query tile
    int qt = 0, sidx = 0;         // query tile; index of this thread's query in s_thr / my_cnt
    int64_t gq = 0;
    uint32_t* gthr = p.shared_thr;
    uint32_t thr_in = 0u, g_pub = 0u;
    float thr = -INFINITY;
    int cnt = 0;
    uint64_t* warp_pools = cta_pools;
    while (!seq.done()) {
      const int t = seq.t();
      const bool first = seq.first_in_item(), last = seq.last_in_item();
      const int item_qt = seq.qt(cta_rank);
      const uint32_t cur_mw = mw_next;
      seq.next();
      if (!seq.done()) mw_next = load_mask(seq.t());
      auto release_accumulator = [&]() {
        // all of this warp's TMEM reads of the accumulator are done: hand it back to the MMA warp
        tcgen05_fence_before();
        __syncwarp();
        if (lane == 0) {
          if (is_leader) mbar_arrive(tempty_bar(acc));
          else mbar_arrive_cluster(tempty_bar(acc) & kPeerBitMask);
        }
        if (++acc == 2) {
          acc = 0;
          acc_phase ^= 1u;
        }
      };
      if (first) {
        // ---- a new work item: load the query tile's state
        qt = item_qt;
        ok = qt < p.q_tiles;
        if (ok) {
          sidx = qt * kBM + ql;
          gq = static_cast<int64_t>(qt) * kBM + ql;
          gthr = p.shared_thr + (gq < p.nq ? gq : 0);
          // Shared threshold: every CTA that holds k_sel candidates for a query publishes its k_sel-th
          // score (atomicMax below).  The global k_sel-th best is >= each of them, so anything strictly
          // below the published maximum can be skipped by everybody; without this each of the ~37 CTAs
          // serving a query tile warms its threshold up on its own and collects ~37x more candidates.
          // The load is consumed one visit later (or at the item's end): its L2 round trip hides behind
          // the first tile.
          g_pub = __ldcg(gthr);
          // The threshold of a query is shared by the two groups through shared memory: whoever holds
          // k_sel candidates proves a lower bound of the final k_sel-th best for everybody.
          thr_in = s_thr[sidx];
          thr = ordered_to_f32(thr_in);
          cnt = my_cnt[sidx];
          warp_pools = cta_pools + ((static_cast<size_t>(qt) * kEpiHalves + grp) * kBM + ew * 32) * p.pool_cap;
          if (lane == 0 && ew == 0) s_touched[qt] = 1;   // (the groups may meet different query tiles)
        }
      } else if (ok) {
        // later visits of the item: the other group's bound, and (once it has arrived) the published one
        thr = fmaxf(thr, ordered_to_f32(s_thr[sidx]));
        if (gq < p.nq && g_pub > 1u) thr = fmaxf(thr, ordered_to_f32(g_pub - 1u));  // keep scores >= published
      }
      if (!ok) {
        // padding query tile of an odd count: nothing to select, just recycle the accumulator
        mbar_wait(tfull_bar(acc), acc_phase);
        tcgen05_fence_after();
        release_accumulator();
        continue;
      }
      const int64_t row0 = static_cast<int64_t>(p.tile_begin + t) * kBN + col0;
      STAT_T(e0);
      mbar_wait(tfull_bar(acc), acc_phase);
      STAT_ADD(1, clock64() - e0);
      tcgen05_fence_after();
      const uint32_t taddr0 = tmem_base + (static_cast<uint32_t>(ew * 32) << 16) + static_cast<uint32_t>(acc * kBN + col0);
      uint32_t va[32], vb[32];
      tmem_ld_32x32(taddr0, va);  // first chunk on its way before anything else is looked at
      STAT_ADD(9, 1);
#ifdef PVDB_BATCH_STATS
      const int cnt_before_visit = cnt;
      int pruned_away = 0;
#endif
      if (p.dump != nullptr) {
        // seed pass: the masked scores themselves are wanted (seed_threshold_kernel selects from them)
        float* drow = p.dump + static_cast<size_t>(gq) * p.dump_ld + (static_cast<size_t>(t) * kBN + col0);
#pragma unroll
        for (int cb = 0; cb < kChunks; ++cb) {
          const uint32_t mw = __shfl_sync(0xffffffffu, cur_mw, cb);
          tmem_ld_wait(va);
#pragma unroll
          for (int j = 0; j < 32; ++j) vb[j] = ((mw >> j) & 1u) ? va[j] : 0xff800000u;
          if (cb + 1 < kChunks) tmem_ld_32x32(taddr0 + static_cast<uint32_t>((cb + 1) * 32), va);
          uint4* d4 = reinterpret_cast<uint4*>(drow + cb * 32);
#pragma unroll
          for (int j = 0; j < 8; ++j) d4[j] = make_uint4(vb[4 * j], vb[4 * j + 1], vb[4 * j + 2], vb[4 * j + 3]);
        }
      } else {
#pragma unroll 1
      for (int cb = 0; cb < kChunks; cb += 2) {
        // mask words of the two 32-column chunks of this step (warp-uniform)
        const uint32_t mw0 = __shfl_sync(0xffffffffu, cur_mw, cb), mw1 = __shfl_sync(0xffffffffu, cur_mw, cb + 1);
        // chunk cb is in flight into `va`; start chunk cb+1 into `vb` before examining `va`
        STAT_T(w0);
        tmem_ld_wait(va);
        STAT_ADD(11, clock64() - w0);
        tmem_ld_32x32(taddr0 + static_cast<uint32_t>((cb + 1) * 32), vb);
        STAT_T(c0);
        if (mw0 != 0u) scan_chunk(va, mw0, thr, warp_pools, lane, cnt, static_cast<uint32_t>(row0) + cb * 32);
        STAT_ADD(12, clock64() - c0);
        STAT_T(w1);
        tmem_ld_wait(vb);
        STAT_ADD(11, clock64() - w1);
        if (cb + 2 < kChunks) tmem_ld_32x32(taddr0 + static_cast<uint32_t>((cb + 2) * 32), va);
        STAT_T(c1);
        if (mw1 != 0u) scan_chunk(vb, mw1, thr, warp_pools, lane, cnt, static_cast<uint32_t>(row0) + (cb + 1) * 32);
        STAT_ADD(12, clock64() - c1);
        // pools that could overflow during the next 64 columns are pruned now (warp co-operative)
        unsigned need = __ballot_sync(0xffffffffu, cnt > p.pool_cap - 64);
        if (need) __syncwarp();
        STAT_T(e1);
        while (need) {
          const int src = __ffs(need) - 1;
          need &= need - 1;
          const int c = __shfl_sync(0xffffffffu, cnt, src);
          int nc;
          float nt;
          prune_pool<NI>(warp_pools, src, c, p.k_sel, lane, nc, nt);
          STAT_ADD(3, 1);
          if (lane == src) {
#ifdef PVDB_BATCH_STATS
            pruned_away += c - nc;
#endif
            cnt = nc;
            if (nt > thr) {
              thr = nt;
              // The other group may be a visit ahead or behind, so a bound proven here is shared
              // NON-strictly (one ulp lower): a row with exactly the k_sel-th score must not be dropped
              // there, it could have the lower row number and win the tie.
              atomicMax(&s_thr[sidx], f32_to_ordered(nt) - 1u);
            }
            if (c >= p.k_sel) atomicMax(gthr, f32_to_ordered(nt));
          }
          __syncwarp();
        }
        STAT_ADD(2, clock64() - e1);
      }
      }  // select (not dump)
#ifdef PVDB_BATCH_STATS
      {
        int appended = cnt + pruned_away - cnt_before_visit;
        for (int o = 16; o > 0; o >>= 1) appended += __shfl_xor_sync(0xffffffffu, appended, o);
        STAT_ADD(4, appended);
      }
#endif
      release_accumulator();
      if (last) {
        // ---- the item ends: hand the state back (bounds taken from elsewhere travel on, non-strictly)
        const uint32_t thr_out = f32_to_ordered(thr) - 1u;
        if (thr_out > thr_in) atomicMax(&s_thr[sidx], thr_out);
        my_cnt[sidx] = static_cast<uint16_t>(cnt);
      }
    }
    STAT_ADD(0, clock64() - e_begin);
    // all visits done: the pools stay as they are (finalize filters and sorts); publish their counts
    asm volatile("bar.sync 1, %0;" ::"n"(kEpiThreads) : "memory");  // the epilogue warps agree on s_touched
    for (int qt = 0; qt < p.q_tiles; ++qt) {
      if (!s_touched[qt] || p.dump != nullptr) continue;
      const size_t u = static_cast<size_t>(blockIdx.x) * p.q_tiles + qt;
      p.counts[(u * kEpiHalves + grp) * kBM + ql] = my_cnt[qt * kBM + ql];
      if (warp == 4 && lane == 0) p.touched[u] = 1;
    }
#ifdef PVDB_BATCH_STATS
    if (lane == 0)
      for (int i = 0; i <= 4; ++i) atomicAdd(&g_batch_stats[i], stat_local[i]);
    if (lane == 0) atomicAdd(&g_batch_stats[9], stat_local[9]);
    if (lane == 0) atomicAdd(&g_batch_stats[11], stat_local[11]);
    if (lane == 0) atomicAdd(&g_batch_stats[12], stat_local[12]);
    if (lane == 0 && warp == 4 && blockIdx.x == 0) atomicAdd(&g_batch_stats[10], 1ull);
#endif
  }

  tcgen05_fence_before();
  // with clusters nobody may leave while a peer can still multicast into, or signal, this CTA
  if constexpr (CL > 1) cluster_sync_all();
  else __syncthreads();
  if (warp == 1) {
    tcgen05_fence_after();
    if constexpr (PAIR)
      asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(kTmemCols) : "memory");
    else
      asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(kTmemCols) : "memory");
  }
}

// ---------------------------------------------------------------------------- radix select helper
// hist[256] counts values per 8-bit digit; called by ONE full warp.  Finds the digit (bin) that
// holds the kk-th largest value (bins are scanned from 255 down) and the rank of that value inside
// the bin.  Requires sum(hist) >= kk.
__device__ __forceinline__ void warp_find_bin_desc(const unsigned* hist, int kk, int lane, int& bin_out, int& rank_out) {
  // lane l owns bins 255 - 8l ... 248 - 8l
  unsigned mine = 0u;
#pragma unroll
  for (int j = 0; j < 8; ++j) mine += hist[255 - 8 * lane - j];
  unsigned incl = mine;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const unsigned up = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += up;
  }
  const unsigned hit = __ballot_sync(0xffffffffu, incl >= static_cast<unsigned>(kk));
  const int owner = __ffs(hit) - 1;
  int bin = 255 - 8 * lane;
  unsigned cum = incl - mine;
  if (lane == owner) {
    for (int j = 0; j < 8; ++j, --bin) {
      const unsigned h = hist[bin];
      if (cum + h >= static_cast<unsigned>(kk)) break;
      cum += h;
    }
  }
  bin_out = __shfl_sync(0xffffffffu, bin, owner);
  rank_out = kk - static_cast<int>(__shfl_sync(0xffffffffu, cum, owner));
}

// ---------------------------------------------------------------------------- finalize
// One block per query.  The candidate pools of every (CTA, half) that met the query's tile are
// streamed through a filter -- a key survives if its score reaches the best lower bound known for the
// query's k_sel-th best (published threshold, sample-pass bound, or the k_sel-th key of what has been
// merged so far) -- into a shared-memory buffer that is sorted (block bitonic network) and cut to
// k_sel whenever it could overflow.  The best k_sel rows are then re-scored exactly in fp32 and the
// top k written.  Keys are unique, so the result does not depend on the order of arrival.
constexpr int kFinalThreads = 256;
constexpr int kFinalCap = 4096;  // keys held in shared memory (32 KB)

// Exactness guard (final merge only).  The tensor-core pass ranks rows by a LOW-PRECISION score
// lowp(r) whose distance to the exact fp32 score is bounded for unit vectors by the input rounding:
//   |exact(r) - lowp(r)| <= ||q - round(q)|| * ||v_r|| + ||round(q)|| * ||v_r - round(v_r)|| + accumulation
// (Cauchy-Schwarz on each operand's rounding error).  Every row that is NOT among the k_sel kept
// candidates has lowp <= L, the k_sel-th best low-precision score, hence exact <= L + eps.  If the
// re-scored k-th best candidate is strictly above L + eps, no outside row can enter the top k and the
// result equals the exact scan's.  Otherwise the query is FLAGGED and re-run on the exact scan path
// (api.cu).  eps uses the query's measured rounding norm and the store's tracked worst row.
struct GuardParams {
  int mode;                   // 0 off; 1 tf32 pass / fp32 rows; 2 bf16 pass / fp32 rows; 3 bf16 pass / bf16-only store
  const float* qeps;          // [nq][4]: ||q - tf32(q)||, ||q - bf16(q)||, ||q||
  const uint32_t* err_words;  // store: max ||v - tf32(v)||^2, max ||v - bf16(v)||^2 (float bits)
  float acc_slop;             // accumulation-order allowance
  unsigned* flag_count;       // flagged queries: count and list (indices into the whole call's batch)
  int* flag_list;
  int64_t q_offset;
};

__device__ __forceinline__ void block_sort_desc(uint64_t* keys, int n_pow2) {
  for (int size = 2; size <= n_pow2; size <<= 1) {
    for (int stride = size >> 1; stride > 0; stride >>= 1) {
      __syncthreads();
      for (int i = threadIdx.x; i < (n_pow2 >> 1); i += blockDim.x) {
        const int lo = 2 * i - (i & (stride - 1));
        const int hi = lo + stride;
        const bool desc = (lo & size) == 0;
        const uint64_t a = keys[lo], b = keys[hi];
        if ((a < b) == desc) {
          keys[lo] = b;
          keys[hi] = a;
        }
      }
    }
  }
  __syncthreads();
}

// sort keys[0..n) descending (n <= kFinalCap); entries past n up to the next power of two are zeroed
__device__ __forceinline__ void block_sort_prefix(uint64_t* keys, int n) {
  int n_pow2 = 2;
  while (n_pow2 < n) n_pow2 <<= 1;
  __syncthreads();
  for (int i = n + threadIdx.x; i < n_pow2; i += blockDim.x) keys[i] = 0ull;
  block_sort_desc(keys, n_pow2);
}

// Keep the best `k_sel` of keys[0..n) (n <= kFinalCap, n > k_sel): radix select (four 8-bit passes over
// the order-preserving score word, shared-memory histogram) finds the score of the k_sel-th best key,
// then every key with at least that score is compacted to the front (unsorted).  Ties on the score
// word are all kept, so the returned count can exceed k_sel by the number of ties; the caller falls
// back to a full sort when that number is large.  Returns the new count; *bound = that score word.
__device__ __forceinline__ int block_select_top(uint64_t* keys, int n, int k_sel, unsigned* hist, uint32_t* s_word,
                                                int* s_int, uint32_t* bound) {
  static_assert(kFinalThreads == 256, "one histogram bin per thread");
  static_assert(kFinalCap % kFinalThreads == 0, "keys are held in registers during compaction");
  const int lane = threadIdx.x & 31;
  uint32_t prefix = 0u, mask = 0u;
  int kk = k_sel;
  for (int shift = 24; shift >= 0; shift -= 8) {
    hist[threadIdx.x] = 0u;
    __syncthreads();
    for (int i = threadIdx.x; i < n; i += kFinalThreads) {
      const uint32_t x = static_cast<uint32_t>(keys[i] >> 32);
      if ((x & mask) == prefix) atomicAdd(&hist[(x >> shift) & 255u], 1u);
    }
    __syncthreads();
    if (threadIdx.x < 32) {
      int bin, rank;
      warp_find_bin_desc(hist, kk, lane, bin, rank);
      if (lane == 0) {
        *s_word = prefix | (static_cast<uint32_t>(bin) << shift);
        *s_int = rank;
      }
    }
    __syncthreads();
    prefix = *s_word;
    kk = *s_int;
    mask |= 255u << shift;
  }
  // compaction: everybody reads its keys first, then the survivors are re-packed from slot 0
  constexpr int kPer = kFinalCap / kFinalThreads;
  uint64_t mine[kPer];
#pragma unroll
  for (int j = 0; j < kPer; ++j) {
    const int i = threadIdx.x + j * kFinalThreads;
    mine[j] = (i < n) ? keys[i] : 0ull;
  }
  if (threadIdx.x == 0) *s_int = 0;
  __syncthreads();
#pragma unroll
  for (int j = 0; j < kPer; ++j)
    if (mine[j] != 0ull && static_cast<uint32_t>(mine[j] >> 32) >= prefix) keys[atomicAdd(s_int, 1)] = mine[j];
  __syncthreads();
  *bound = prefix;
  const int kept = *s_int;
  __syncthreads();  // s_int is reused by the caller
  return kept;
}

__global__ void __launch_bounds__(kFinalThreads)
finalize_batch_kernel(const uint64_t* __restrict__ pools, const uint16_t* __restrict__ counts,
                      const uint8_t* __restrict__ touched, int n_ctas, int pool_cap, int k_sel, int q_tiles,
                      int64_t nq, int k, const float* __restrict__ qn, int ldq, const float* __restrict__ f32,
                      int ld32, const __nv_bfloat16* __restrict__ b16, int ld16, int rescore, int64_t row_base,
                      float* __restrict__ out_scores,
                      int64_t* __restrict__ out_rows, const uint64_t* __restrict__ carry_in,
                      uint64_t* __restrict__ carry_out, float* __restrict__ thr_out,
                      const uint32_t* __restrict__ shared_thr, const float* __restrict__ init_thr,
                      const GuardParams guard) {
  __shared__ uint64_t keys[kFinalCap];
  __shared__ int s_lists[kNumSMs];
  __shared__ int s_nlists;
  __shared__ int s_n;
  __shared__ unsigned s_hist[256];
  __shared__ uint32_t s_word;
  __shared__ int s_int;
  const int64_t q = blockIdx.x;
  const int qt = static_cast<int>(q / kBM);
  const int ql = static_cast<int>(q % kBM);
  if (threadIdx.x == 0) {
    int n = 0;
    for (int b = 0; b < n_ctas; ++b)
      if (touched[static_cast<size_t>(b) * q_tiles + qt]) s_lists[n++] = b;
    s_nlists = n;
  }
  // lower bound (ordered image) every surviving key must reach
  uint32_t lb = 0u;
  {
    const uint32_t g = shared_thr[q];
    if (g > 1u) lb = g;
    if (init_thr != nullptr) {
      const float b = init_thr[q];
      if (b > -INFINITY) lb = max(lb, f32_to_ordered(b));
    }
  }
  int n_keys = 0;  // keys[0..n_keys) hold the survivors so far (block-uniform)
  if (carry_in != nullptr) {
    // the sample pass already produced a sorted (zero padded) list of k_sel keys for this query
    for (int j = threadIdx.x; j < k_sel; j += blockDim.x) keys[j] = carry_in[q * k_sel + j];
    n_keys = k_sel;
  }
  if (threadIdx.x == 0) s_n = n_keys;
  __syncthreads();
  const int n_pools = s_nlists * kEpiHalves;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  // Cut the buffer to the best k_sel keys and raise the bound (cheap select; a full sort only when a
  // crowd of keys ties on the score word, e.g. many copies of one vector).
  auto cut = [&]() {
    if (n_keys > k_sel) {
      uint32_t bound;
      int kept_now = block_select_top(keys, n_keys, k_sel, s_hist, &s_word, &s_int, &bound);
      if (kept_now > k_sel + 64) {
        block_sort_prefix(keys, kept_now);
        kept_now = k_sel;
        bound = static_cast<uint32_t>(keys[k_sel - 1] >> 32);
        __syncthreads();
      }
      n_keys = kept_now;
      lb = max(lb, bound);
    }
    if (threadIdx.x == 0) s_n = n_keys;
    __syncthreads();
  };
  int c = 0;
  while (c < n_pools) {
    // every pool taken this round may contribute up to pool_cap keys
    const int take = min(n_pools - c, (kFinalCap - n_keys) / pool_cap);
    if (take == 0) {
      cut();
      continue;
    }
    // one warp per pool, only the slots in use
    for (int pi = c + warp; pi < c + take; pi += kFinalThreads / 32) {
      const size_t u = static_cast<size_t>(s_lists[pi / kEpiHalves]) * q_tiles + qt;
      const size_t slot = (u * kEpiHalves + (pi % kEpiHalves)) * kBM + ql;
      const int cnt = static_cast<int>(counts[slot]);
      // the 32 pools of an epilogue warp are interleaved (pool_slot)
      const uint64_t* wpool = pools + (slot & ~size_t(31)) * pool_cap;
      for (int j = lane; j < cnt; j += 32) {
        const uint64_t key = wpool[pool_slot(ql & 31, j)];
        if (key != 0ull && static_cast<uint32_t>(key >> 32) >= lb) keys[atomicAdd(&s_n, 1)] = key;
      }
    }
    __syncthreads();
    n_keys = s_n;
    __syncthreads();  // nobody appends again (next round) before everyone has read the count
    c += take;
  }
  cut();
  block_sort_prefix(keys, n_keys);  // at most k_sel + 64 keys
  const int kept = min(n_keys, k_sel);
  // keys[0..kept) sorted descending by tensor-core score
  if (carry_out != nullptr) {
    // sample pass: hand the merged list and its k_sel-th score (a proven lower bound of the final
    // k_sel-th best) to the main pass instead of producing results
    for (int j = threadIdx.x; j < k_sel; j += blockDim.x) carry_out[q * k_sel + j] = (j < kept) ? keys[j] : 0ull;
    if (threadIdx.x == 0) {
      const uint64_t kth = (kept >= k_sel) ? keys[k_sel - 1] : 0ull;
      thr_out[q] = kth ? key_score(kth) : -INFINITY;
    }
    return;
  }
  // weakest kept candidate's low-precision score: rows outside the list score no more than this
  // (an empty k_sel-th slot means every eligible row is in the list: nothing outside)
  const uint64_t weakest = n_keys >= k_sel ? keys[k_sel - 1] : 0ull;
  const bool have_outside = weakest != 0ull;
  const float lowp_bound = have_outside ? key_score(weakest) : -INFINITY;
  __syncthreads();  // everybody has read the low-precision key before re-scoring overwrites it
  if (rescore) {
    // exact fp32-accumulated dot product of the fp32 query with each surviving row (one warp per
    // candidate): against the fp32 matrix when the store keeps one, else against the bf16 mirror --
    // the same arithmetic as the single-query scan of that store, so a query gets the same answer
    // alone and inside a batch
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const float4* q4 = reinterpret_cast<const float4*>(qn + q * ldq);
    for (int cnd = warp; cnd < kept; cnd += kFinalThreads / 32) {
      const uint64_t key = keys[cnd];
      if (key == 0ull) continue;
      const uint32_t row = key_row(key);
      float acc = 0.f;
      if (f32 != nullptr) {
        const float4* v4 = reinterpret_cast<const float4*>(f32 + static_cast<size_t>(row) * ld32);
        for (int i = lane; i < (ld32 >> 2); i += 32) {
          const float4 a = __ldg(v4 + i);
          const float4 b = q4[i];
          acc = fmaf(a.x, b.x, acc);
          acc = fmaf(a.y, b.y, acc);
          acc = fmaf(a.z, b.z, acc);
          acc = fmaf(a.w, b.w, acc);
        }
      } else {
        const uint4* v8 = reinterpret_cast<const uint4*>(b16 + static_cast<size_t>(row) * ld16);
        for (int i = lane; i < (ld16 >> 3); i += 32) {
          const uint4 a = __ldg(v8 + i);
          const float4 b0 = q4[2 * i], b1 = q4[2 * i + 1];
          acc = fmaf(__uint_as_float(a.x << 16), b0.x, acc);
          acc = fmaf(__uint_as_float(a.x & 0xffff0000u), b0.y, acc);
          acc = fmaf(__uint_as_float(a.y << 16), b0.z, acc);
          acc = fmaf(__uint_as_float(a.y & 0xffff0000u), b0.w, acc);
          acc = fmaf(__uint_as_float(a.z << 16), b1.x, acc);
          acc = fmaf(__uint_as_float(a.z & 0xffff0000u), b1.y, acc);
          acc = fmaf(__uint_as_float(a.w << 16), b1.z, acc);
          acc = fmaf(__uint_as_float(a.w & 0xffff0000u), b1.w, acc);
        }
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
      if (lane == 0) keys[cnd] = make_key(acc, row);
    }
    block_sort_prefix(keys, kept);
    if (guard.mode != 0 && threadIdx.x == 0 && have_outside && kept >= k) {
      const float* e = guard.qeps + q * 4;
      const float qn_norm = fmaxf(e[2], 1.f);
      const float ev_tf = sqrtf(__uint_as_float(guard.err_words[0])), ev_bf = sqrtf(__uint_as_float(guard.err_words[1]));
      float eps = guard.acc_slop;
      if (guard.mode == 1) eps += e[0] * 1.0001f + qn_norm * ev_tf * 1.0001f;   // rows have unit norm
      else if (guard.mode == 2) eps += e[1] * 1.0001f + qn_norm * ev_bf * 1.0001f;
      else eps += e[1] * 1.0001f;                                              // the stored rows are exact
      const float exact_kth = key_score(keys[k - 1]);
      if (!(exact_kth > lowp_bound + eps)) guard.flag_list[atomicAdd(guard.flag_count, 1u)] = static_cast<int>(guard.q_offset + q);
    }
  }
  for (int j = threadIdx.x; j < k; j += blockDim.x) {
    const uint64_t key = (j < kept) ? keys[j] : 0ull;
    out_scores[q * k + j] = key ? key_score(key) : -INFINITY;
    out_rows[q * k + j] = key ? row_base + static_cast<int64_t>(key_row(key)) : -1ll;
  }
}

// ---------------------------------------------------------------------------- seed thresholds
// One block per query: the k_sel-th largest of the query's `n` dumped scores (masked ones are -inf) is
// a proven lower bound of its final k_sel-th best tensor-core score -- the selection passes start from
// it instead of -inf.  Radix select over the order-preserving 32-bit image of the scores: four 8-bit
// passes over a shared-memory copy, each narrowing the prefix of the k-th largest value.
constexpr int kSeedTilesMax = 64;
constexpr int kSeedColsMax = kSeedTilesMax * kBN;  // 16384 scores = 64 KB of (dynamic) shared memory
constexpr int kSeedThreads = 256;

__global__ void __launch_bounds__(kSeedThreads)
seed_threshold_kernel(const float* __restrict__ dump, int dump_ld, int n, int k_sel, float* __restrict__ thr_out) {
  extern __shared__ uint32_t v[];  // n scores
  __shared__ unsigned hist[256];
  __shared__ uint32_t s_prefix;
  __shared__ int s_k;
  const int64_t q = blockIdx.x;
  const int lane = threadIdx.x & 31;
  if (n < k_sel) {
    if (threadIdx.x == 0) thr_out[q] = -INFINITY;
    return;
  }
  for (int i = threadIdx.x; i < n; i += kSeedThreads) v[i] = f32_to_ordered(dump[q * dump_ld + i]);
  uint32_t prefix = 0u, mask = 0u;
  int kk = k_sel;  // rank (1 = largest) of the wanted value among those that match the prefix
  for (int shift = 24; shift >= 0; shift -= 8) {
    hist[threadIdx.x] = 0u;  // kSeedThreads == 256 bins
    __syncthreads();
    for (int i = threadIdx.x; i < n; i += kSeedThreads) {
      const uint32_t x = v[i];
      if ((x & mask) == prefix) atomicAdd(&hist[(x >> shift) & 255u], 1u);
    }
    __syncthreads();
    if (threadIdx.x < 32) {
      int bin, rank;
      warp_find_bin_desc(hist, kk, lane, bin, rank);
      if (lane == 0) {
        s_prefix = prefix | (static_cast<uint32_t>(bin) << shift);
        s_k = rank;
      }
    }
    __syncthreads();
    prefix = s_prefix;
    kk = s_k;
    mask |= 255u << shift;
  }
  if (threadIdx.x == 0) thr_out[q] = prefix > f32_to_ordered(-INFINITY) ? ordered_to_f32(prefix) : -INFINITY;
}

// ---------------------------------------------------------------------------- host side
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(ptr);
    (void)cudaGetLastError();
  }
  return fn;
}

bool batch_path_available() { return get_encode_fn() != nullptr; }

// 2-D row-major matrix [rows][inner] with `ld` elements between rows; box = 128 bytes x box_rows
static int encode_map(CUtensorMap* map, bool bf16, const void* base, int inner, int64_t rows, int ld, int box_rows) {
  EncodeTiledFn fn = get_encode_fn();
  if (!fn) return fail(PVDB_ERR_CUDA, "cuTensorMapEncodeTiled is not available from this driver");
  const cuuint64_t gdim[2] = {static_cast<cuuint64_t>(inner), static_cast<cuuint64_t>(rows)};
  const cuuint64_t gstride[1] = {static_cast<cuuint64_t>(ld) * (bf16 ? 2u : 4u)};
  const cuuint32_t box[2] = {static_cast<cuuint32_t>(bf16 ? 64 : 32), static_cast<cuuint32_t>(box_rows)};
  const cuuint32_t estride[2] = {1, 1};
  CUresult r = fn(map, bf16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2,
                  const_cast<void*>(base), gdim, gstride, box, estride, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail(PVDB_ERR_CUDA, "cuTensorMapEncodeTiled failed with CUresult %d", static_cast<int>(r));
  return PVDB_OK;
}

// Kernel variant for (element type, cluster size, pair MMA, pool size).
template <bool BF16, int CL, bool PAIR>
static const void* batch_kernel_ptr(int pool_cap) {
  return pool_cap == 128 ? reinterpret_cast<const void*>(batch_topk_kernel<BF16, 4, CL, PAIR>)
                         : reinterpret_cast<const void*>(batch_topk_kernel<BF16, 8, CL, PAIR>);
}

template <bool BF16>
static const void* batch_kernel_cl(int cl, bool pair, int pool_cap) {
  if (pair) return batch_kernel_ptr<BF16, 2, true>(pool_cap);
  switch (cl) {
    case 8: return batch_kernel_ptr<BF16, 8, false>(pool_cap);
    case 4: return batch_kernel_ptr<BF16, 4, false>(pool_cap);
    case 2: return batch_kernel_ptr<BF16, 2, false>(pool_cap);
    default: return batch_kernel_ptr<BF16, 1, false>(pool_cap);
  }
}

static const void* batch_kernel(bool bf16, int cl, bool pair, int pool_cap) {
  return bf16 ? batch_kernel_cl<true>(cl, pair, pool_cap) : batch_kernel_cl<false>(cl, pair, pool_cap);
}

static void batch_launch_config(cudaLaunchConfig_t& cfg, cudaLaunchAttribute* attr, int cl, int grid, cudaStream_t st) {
  cfg = cudaLaunchConfig_t{};
  cfg.gridDim = dim3(static_cast<unsigned>(grid));
  cfg.blockDim = dim3(kBatchThreads);
  cfg.dynamicSmemBytes = kBatchSmem;
  cfg.stream = st;
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = static_cast<unsigned>(cl);
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
}

// How many clusters of `cl` CTAs of this kernel the device can hold at once (one CTA per SM; a
// cluster must fit a GPC, so sizes above 2 may leave a few SMs unused).  The static visit schedule
// needs every unit resident, so the grid never exceeds this.
static int batch_max_units(const void* kern, int cl, int* out) {
  // Function attributes and occupancy belong to a (device, kernel) pair: a process that opens stores
  // on two GPUs must opt each device into the large dynamic shared memory separately.
  struct Entry { int device; const void* kern; int units; };
  static std::mutex mu;
  static std::vector<Entry> cache;
  int device = 0;
  PVDB_CUDA(cudaGetDevice(&device));
  std::lock_guard<std::mutex> g(mu);
  for (auto& e : cache)
    if (e.device == device && e.kern == kern) {
      *out = e.units;
      return PVDB_OK;
    }
  PVDB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(kBatchSmem)));
  if (cl > 8) PVDB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
  cudaLaunchConfig_t cfg;
  cudaLaunchAttribute attr[1];
  batch_launch_config(cfg, attr, cl, cl * (kNumSMs / cl), nullptr);
  int n = 0;
  PVDB_CUDA(cudaOccupancyMaxActiveClusters(&n, kern, &cfg));
  if (n < 1) return fail(PVDB_ERR_CUDA, "batch: the device cannot hold one cluster of %d CTAs", cl);
  n = std::min(n, kNumSMs / cl);
  cache.push_back({device, kern, n});
  *out = n;
  return PVDB_OK;
}

static int launch_batch(const void* kern, int cl, const CUtensorMap& mq, const CUtensorMap& mdb, const BatchParams& p,
                        int grid, cudaStream_t st) {
  cudaLaunchConfig_t cfg;
  cudaLaunchAttribute attr[1];
  batch_launch_config(cfg, attr, cl, grid, st);
  void* args[3] = {const_cast<CUtensorMap*>(&mq), const_cast<CUtensorMap*>(&mdb), const_cast<BatchParams*>(&p)};
  PVDB_CUDA(cudaLaunchKernelExC(&cfg, kern, args));
  PVDB_LAUNCH_CHECK();
  return PVDB_OK;
}

int64_t batch_query_rows(int64_t nq) { return (nq + kBM - 1) / kBM * kBM; }

// candidates kept per query for a top-k request
static int batch_k_sel(int k, bool use_bf16, bool rescore) {
  if (!rescore) return k;
  const int base = use_bf16 ? kSlackBF16 : kSlackTF32;
  const int slack = k <= 32 ? base : std::max(base, std::min(kSlackLargeK, kMaxSel - k));
  return k + slack;
}

int batch_max_k(bool use_bf16, bool rescore) { return kMaxSel - (rescore ? (use_bf16 ? kSlackBF16 : kSlackTF32) : 0); }

int search_batch(pvdb_store* s, bool use_bf16, const float* d_qn, const __nv_bfloat16* d_qn16, int64_t nq_total,
                 int k, const uint32_t* d_pref, bool no_rescore, const float* d_qeps, unsigned* d_flag_count,
                 int* d_flag_list, float* d_out_scores, int64_t* d_out_rows, cudaStream_t st) {
  // candidates are always re-scored in fp32 arithmetic unless the caller asks for the raw tensor-core
  // scores: against the fp32 matrix when there is one, else against the bf16 mirror
  const bool rescore = !no_rescore;
  const int k_sel = batch_k_sel(k, use_bf16, rescore);
  if (k_sel > kMaxSel) return fail(PVDB_ERR_UNSUPPORTED, "batch: k=%d too large for the fused tensor-core path", k);

  // Two query tiles or more: 2-CTA clusters share every database tile through TMA multicast (each
  // CTA fetches half of the tile's rows), which cuts the L2 -> shared-memory traffic per CTA from
  // 48 KB to 32 KB per K block.  A single query tile has nobody to share with.
  // Clusters of CL CTAs take CL consecutive query tiles of the same database tile; each CTA fetches
  // 1/CL of the tile's rows and multicasts it.  L2 -> SM traffic per CTA and K block is 16 KB (its
  // query slice) + 32/CL KB; the L2 delivers ~43 B/clk per SM, the tensor core consumes a K block in
  // 512 clk (tf32) -- CL = 1: 94 B/clk needed, 2: 64, 4: 48, 8: 40.
  int cl_max = getenv("PVDB_BATCH_NO_CLUSTER") ? 1 : kClusterDefault;
  if (const char* e = getenv("PVDB_BATCH_CLUSTER")) cl_max = std::max(1, atoi(e));
  // cta_group::2 MMAs (one M=256 instruction per CTA pair); PVDB_BATCH_PAIR=0 keeps the multicast variant
  const char* pair_env = getenv("PVDB_BATCH_PAIR");
  const bool pair_mma = pair_env ? atoi(pair_env) != 0 : kPairDefault;
  // database tiles per work item (VisitSeq); 0 = choose per pass
  int tile_block_env = 0;
  if (const char* e = getenv("PVDB_BATCH_TILE_BLOCK")) tile_block_env = std::min(kMaxTileBlock, std::max(1, atoi(e)));
  const void* db_ptr = use_bf16 ? s->bf16.ptr : s->f32.ptr;
  const int db_ld = use_bf16 ? s->ld_bf16 : s->ld_f32;

  const int64_t max_q = static_cast<int64_t>(kMaxQTiles) * kBM;
  const int total_tiles = static_cast<int>((s->rows + kBN - 1) / kBN);
  for (int64_t q0 = 0; q0 < nq_total; q0 += max_q) {
    const int64_t nq = std::min(max_q, nq_total - q0);
    GuardParams guard{};
    if (rescore && d_qeps != nullptr) {
      guard.mode = !use_bf16 ? 1 : (s->f32.ptr != nullptr ? 2 : 3);
      guard.qeps = d_qeps + q0 * 4;
      guard.err_words = s->d_err_words;
      // fp32 accumulation inside the tensor core (one rounding per K step of 8 / 16 elements, |partial
      // sums| <= 1) plus the re-scoring's own summation order: a generous (dim / 4 + 16) ulps of 1
      guard.acc_slop = (static_cast<float>(s->dim) / 4.f + 16.f) * 1.1920929e-7f;
      guard.flag_count = d_flag_count;
      guard.flag_list = d_flag_list;
      guard.q_offset = q0;
    }
    BatchParams p{};
    p.nq = nq;
    p.n_rows = s->rows;
    p.k_blocks = use_bf16 ? (s->dim + 63) / 64 : (s->dim + 31) / 32;
    p.q_tiles = static_cast<int>((nq + kBM - 1) / kBM);
    p.k_sel = k_sel;
    // the epilogue prunes when fewer than 64 free slots remain; leave at least 32 slots between
    // prunes (cap >= k_sel + 96)
    p.pool_cap = k_sel <= 32 ? 128 : 256;
    p.active = static_cast<const uint32_t*>(s->active.ptr);
    p.prefilter = d_pref;

    // Sample pass: the first 1/16 of the tiles is searched on its own; the k_sel-th best score it
    // finds for a query is a proven lower bound of that query's final k_sel-th best, so the main
    // pass starts every (CTA, query) state at that threshold instead of warming each one up from
    // -inf (which costs ~k_sel * ln(rows/k_sel) pool insertions per state).  Both passes feed the
    // same final merge, so the result is the exact top k either way.
    int sample_tiles = total_tiles / kSampleFraction;
    if (const char* e = getenv("PVDB_BATCH_SAMPLE")) sample_tiles = atoi(e) > 0 ? total_tiles / atoi(e) : 0;
    int64_t min_sample_visits = 4 * kNumSMs;  // below this the sample pass is too small to pay off
    if (const char* e = getenv("PVDB_BATCH_SAMPLE_MINV")) min_sample_visits = atoll(e);
    if (static_cast<int64_t>(sample_tiles) * p.q_tiles < min_sample_visits) sample_tiles = 0;

    const int grid_max = kNumSMs;  // pools / flags are sized for a full grid
    const size_t n_pools = static_cast<size_t>(grid_max) * p.q_tiles * kEpiHalves * kBM;
    const size_t pool_bytes = n_pools * p.pool_cap * sizeof(uint64_t);
    const size_t count_bytes = (n_pools * sizeof(uint16_t) + 255) & ~size_t(255);
    const size_t touched_bytes = (static_cast<size_t>(grid_max) * p.q_tiles + 255) & ~size_t(255);
    const size_t thr_bytes = (static_cast<size_t>(nq) * sizeof(uint32_t) + 255) & ~size_t(255);
    const size_t carry_bytes = sample_tiles ? ((static_cast<size_t>(nq) * k_sel * sizeof(uint64_t) + 255) & ~size_t(255)) : 0;
    const size_t init_bytes = sample_tiles ? thr_bytes : 0;
    // Seed pass: the first tiles (up to 8192 rows) are multiplied once more with the scores written
    // out -- a plain GEMM, no selection -- and their k_sel-th best per query starts every threshold
    // (seed_threshold_kernel).  That removes the cold start (every score of the first tiles appended,
    // pools pruned over and over): the selection passes begin at a pass rate of k_sel / seed rows.
    int seed_tiles = std::min(32, total_tiles / 8);  // these tiles are multiplied twice (<= 1/8 of a small store)
    if (const char* e = getenv("PVDB_BATCH_SEED_TILES")) seed_tiles = std::min(kSeedTilesMax, atoi(e));
    const bool seed = seed_tiles >= 2 && getenv("PVDB_BATCH_NO_SEED") == nullptr;
    const int seed_cols = seed_tiles * kBN;
    const size_t seed_thr_bytes = seed ? thr_bytes : 0;
    const size_t dump_bytes = seed ? static_cast<size_t>(p.q_tiles) * kBM * seed_cols * sizeof(float) : 0;
    PVDB_TRY(s->d_misc.ensure(pool_bytes + touched_bytes + thr_bytes + carry_bytes + init_bytes + count_bytes +
                              seed_thr_bytes + dump_bytes));
    unsigned char* base = static_cast<unsigned char*>(s->d_misc.ptr);
    p.pools = reinterpret_cast<uint64_t*>(base);
    p.touched = base + pool_bytes;
    p.shared_thr = reinterpret_cast<uint32_t*>(p.touched + touched_bytes);
    uint64_t* carry = reinterpret_cast<uint64_t*>(base + pool_bytes + touched_bytes + thr_bytes);
    float* init_thr = reinterpret_cast<float*>(base + pool_bytes + touched_bytes + thr_bytes + carry_bytes);
    p.counts = reinterpret_cast<uint16_t*>(base + pool_bytes + touched_bytes + thr_bytes + carry_bytes + init_bytes);
    float* seed_thr = reinterpret_cast<float*>(reinterpret_cast<unsigned char*>(p.counts) + count_bytes);
    float* dump = reinterpret_cast<float*>(reinterpret_cast<unsigned char*>(seed_thr) + seed_thr_bytes);

    CUtensorMap mq;
    // the caller pads the prepared queries with zero rows to whole TMA boxes (batch_query_rows)
    const int64_t nq_box = batch_query_rows(nq);
    // K extent = whole boxes: the query rows are zero padded up to ldq, so the database map below may
    // let the last box of a row run into the next row's first values (finite; times zero)
    const int k_ext = p.k_blocks * (use_bf16 ? 64 : 32);
    if (use_bf16) PVDB_TRY(encode_map(&mq, true, d_qn16 + q0 * s->ldq, k_ext, nq_box, s->ldq, kBM));
    else PVDB_TRY(encode_map(&mq, false, d_qn + q0 * s->ldq, k_ext, nq_box, s->ldq, kBM));

    auto run_pass = [&](int tile_begin, int n_tiles, const float* thr_in, const uint64_t* carry_in, uint64_t* carry_out,
                        float* thr_out, bool pin_query_tiles, bool dump_scores = false) -> int {
      p.tile_begin = tile_begin;
      p.n_tiles = n_tiles;
      p.init_thr = thr_in;
      p.dump = dump_scores ? dump : nullptr;
      p.dump_ld = seed_cols;
      // largest allowed cluster that does not pad the query tiles by more than a quarter (three query
      // tiles run unclustered: 300 queries over 10M x 768 took 9.4-10.7 ms that way, 11.9-12.7 ms as
      // two clusters with a padding tile)
      int cl = 1;
      for (int c = 2; c <= cl_max && c <= 8; c <<= 1)
        if (p.q_tiles >= c && ((p.q_tiles + c - 1) / c) * c * 4 <= p.q_tiles * 5) cl = c;
      const bool pair = pair_mma && cl >= 2;
      if (pair) cl = 2;
      const void* kern = batch_kernel(use_bf16, cl, pair, p.pool_cap);
      int max_units = 0;
      PVDB_TRY(batch_max_units(kern, cl, &max_units));
      CUtensorMap mdb;  // box = the rows one CTA fetches per K block
      PVDB_TRY(encode_map(&mdb, use_bf16, db_ptr, k_ext, s->capacity, db_ld, kBN / cl));
      const int n_qp = (p.q_tiles + cl - 1) / cl;
      const int64_t n_visits = static_cast<int64_t>(n_tiles) * n_qp;
      // Work items of R consecutive tiles for one query-tile group (VisitSeq): as large as leaves every
      // unit >= 16 items (the last wave of items is the load imbalance), at most kMaxTileBlock.
      int tile_block = static_cast<int>(std::min<int64_t>(kMaxTileBlock, n_visits / (16 * static_cast<int64_t>(max_units))));
      if (tile_block_env > 0) tile_block = tile_block_env;
      p.tile_block = std::max(1, std::min(tile_block, n_tiles));
      const int64_t n_items = static_cast<int64_t>((n_tiles + p.tile_block - 1) / p.tile_block) * n_qp;
      const int grid = cl * static_cast<int>(std::min<int64_t>(n_items, max_units));
      // pinned schedule: stride = (query-tile pairs) x (units per pair), so a unit never changes pair
      const int n_units = grid / cl;
      p.visit_stride = (pin_query_tiles && n_units >= n_qp) ? n_qp * (n_units / n_qp) : 0;
      PVDB_CUDA(cudaMemsetAsync(p.touched, 0, touched_bytes + thr_bytes, st));
      PVDB_TRY(launch_batch(kern, cl, mq, mdb, p, grid, st));
      if (dump_scores) {
        const int n = static_cast<int>(std::min<int64_t>(static_cast<int64_t>(n_tiles) * kBN, s->rows - static_cast<int64_t>(tile_begin) * kBN));
        const size_t seed_smem = static_cast<size_t>(seed_cols) * sizeof(uint32_t);
        if (seed_smem > 48 * 1024)
          PVDB_CUDA(cudaFuncSetAttribute(seed_threshold_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         static_cast<int>(kSeedColsMax * sizeof(uint32_t))));
        seed_threshold_kernel<<<static_cast<unsigned>(nq), kSeedThreads, seed_smem, st>>>(dump, seed_cols, n, p.k_sel,
                                                                                          thr_out);
        PVDB_LAUNCH_CHECK();
        return PVDB_OK;
      }
      finalize_batch_kernel<<<static_cast<unsigned>(nq), kFinalThreads, 0, st>>>(
          p.pools, p.counts, p.touched, grid, p.pool_cap, p.k_sel, p.q_tiles, nq, k, d_qn + q0 * s->ldq, s->ldq,
          static_cast<const float*>(s->f32.ptr), s->ld_f32, static_cast<const __nv_bfloat16*>(s->bf16.ptr),
          s->ld_bf16, rescore ? 1 : 0, s->row_base, d_out_scores + q0 * k,
          d_out_rows + q0 * k, carry_in, carry_out, thr_out, p.shared_thr, thr_in,
          carry_out == nullptr ? guard : GuardParams{});
      PVDB_LAUNCH_CHECK();
      return PVDB_OK;
    };
    const float* thr0 = nullptr;
    if (seed) {
      PVDB_TRY(run_pass(0, seed_tiles, nullptr, nullptr, nullptr, seed_thr, false, true));
      thr0 = seed_thr;
    }
    if (sample_tiles > 0) {
      PVDB_TRY(run_pass(0, sample_tiles, thr0, nullptr, carry, init_thr, true));
      PVDB_TRY(run_pass(sample_tiles, total_tiles - sample_tiles, init_thr, carry, nullptr, nullptr, false));
    } else {
      // small problem, single pass: pinning each unit to one query tile (pair) keeps it to ONE
      // threshold state; the few surplus units that idle cost less than several states per CTA
      PVDB_TRY(run_pass(0, total_tiles, thr0, nullptr, nullptr, nullptr, true));
    }
  }
  return PVDB_OK;
}

}  // namespace pvdb

#ifdef PVDB_BATCH_STATS
extern "C" int pvdb_debug_batch_stats(unsigned long long* out, int n, int reset) {
  unsigned long long zero[16] = {};
  if (out && n > 0)
    PVDB_CUDA(cudaMemcpyFromSymbol(out, pvdb::g_batch_stats, sizeof(unsigned long long) * (n < 16 ? n : 16)));
  if (reset) PVDB_CUDA(cudaMemcpyToSymbol(pvdb::g_batch_stats, zero, sizeof(zero)));
  return PVDB_OK;
}
#endif
