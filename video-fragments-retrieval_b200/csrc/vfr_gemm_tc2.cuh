// The split-operand tensor-core GEMM of vfr_gemm_tc.cuh as a PERSISTENT CTA-PAIR kernel (tcgen05 cta_group::2):
//   C[M,N] = A[M,K] * B[N,K]^T,  A.B^T ~= Ah.Bh^T + Al.Bh^T + Ah.Bl^T  (same packed operands, same epilogue functors).
//
// Why.  Measured over the 66 294 working tiles of a 37 888-query K3 (tools/k3_ab.py, round 2), a 256 x 256 tile of the
// one-CTA kernel takes 108.9 k cycles for 55.3 k cycles of MMA: 29.4 k are the epilogue (LSTM cell, exposed: the tile's
// accumulators fill all 512 TMEM columns, nothing can run under them), 17.9 k main-loop stalls on a 3-stage ring,
// 4.3 k set-up + first-operand latency per tile, 2 k tear-down.  Here
//  * two CTAs of a cluster (one TPC) share every tile: each holds 128 of the 256 rows, so a tile's accumulator is 256
//    TMEM columns per CTA and the OTHER 256 columns take the next tile - the epilogue of tile i runs under the MMAs of
//    tile i + 1 (8 epilogue warps per CTA, two per TMEM lane quarter, 128 columns each);
//  * each CTA stages only its half of B (tcgen05.mma.cta_group::2 reads both halves), so a K chunk is 32 KB per CTA instead
//    of 64 KB - 6 stages in the same shared memory, same bytes per MMA cycle;
//  * CTAs are persistent (one pair per TPC, static round-robin over the tiles, N fastest so that the pairs working at the
//    same time share the A rows in L2): barriers, TMEM and tensor-map fetches are set up once per kernel, and the ring
//    never drains between tiles.
// Protocol (as CUTLASS's 2-SM pipelines): the LEADER CTA (cluster rank 0) issues every MMA; both CTAs' TMA loads
// (cta_group::2) complete on the leader's `full` barrier, which the leader's producer arms with the bytes of both;
// tcgen05.commit multicasts `empty` (stage free) and `acc_full` (accumulator complete) to both CTAs; the epilogue warps of
// both CTAs arrive on the leader's `acc_empty`.
#pragma once
#include "vfr_gemm_tc.cuh"
#include <algorithm>

namespace vfr {

// K chunks of 64 columns: rows of 128 B (SWIZZLE_128B), i.e. every row of a TMA box is one full L2 line.  With chunks of 32
// (64-byte rows, 6 stages) the MMA issuer of a 37 888-query K3 step waited for operands 46 % of its time at only 25 B per
// clock and SM (tools/k3_ab.py, round 2).
constexpr int G2_BK = 64;
constexpr int G2_STAGES = 3;
constexpr int G2_SUB = 128 * G2_BK * 2;                   // one [128 x 64] 16-bit box = 16 KB
constexpr int G2_STAGE = 4 * G2_SUB;                      // Ah Al Bh Bl (this CTA's halves) = 64 KB
constexpr uint32_t G2_SMEM = G2_STAGES * G2_STAGE + 1024 + 512;
constexpr int G2_EPI_WARPS = GT_THREADS / 32 - 2;         // 8

// K-major SWIZZLE_128B operand tile: rows of 128 B, 8-row groups 1024 B apart
__device__ __forceinline__ uint64_t g2_desc(const void* smem_ptr) {
  const uint32_t addr = smem_u32(smem_ptr);
  uint64_t d = 0;
  d |= (uint64_t)((addr >> 4) & 0x3fff);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}
__device__ __forceinline__ uint32_t g2_cta_rank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void g2_cluster_sync() {
  asm volatile("barrier.cluster.arrive.release.aligned;\nbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// shared::cluster address of the same shared-memory location in CTA `rank` of the cluster
__device__ __forceinline__ uint32_t g2_mapa(const void* p, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(smem_u32(p)), "r"(rank));
  return r;
}
__device__ __forceinline__ void g2_arrive_remote(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
// TMA tile load of a CTA pair: data into THIS CTA's shared memory, completion bytes on the barrier at `bar_cluster_addr`
// (the leader's)
__device__ __forceinline__ void g2_tma_load(void* dst, const CUtensorMap* map, int c0, int c1, uint32_t bar_cluster_addr) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(
          smem_u32(dst)),
      "l"(map), "r"(bar_cluster_addr), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void g2_mma(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(d_tmem),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc)
      : "memory");
}
// arrives (once the MMAs issued so far have completed) on the barrier at this shared-memory offset in BOTH CTAs
__device__ __forceinline__ void g2_commit_both(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
                   smem_u32(bar)),
               "h"((uint16_t)3)
               : "memory");
}

// One lane of a CONVERGED warp; deterministic (the same lane for the same mask every time), so the MMAs and the commits
// that track them come from one thread.
__device__ __forceinline__ bool g2_elect() {
  uint32_t pred;
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "elect.sync _|p, 0xffffffff;\n"
      "selp.u32 %0, 1, 0, p;\n"
      "}\n"
      : "=r"(pred));
  return pred != 0;
}
// The 12 MMAs of one 64-column K chunk (4 k steps x {Ah.Bh, Al.Bh, Ah.Bl}, in that order) as ONE asm block issued by the
// elected lane of a converged warp.  Issued one by one from `if (lane == 0)` code, the compiler wraps every tcgen05.mma in an
// ELECT / BRA.U.ANY loop and re-derives its descriptors (~15 instructions, ~120 cycles of issue per MMA - as long as the
// MMA runs, so the issuer never gets ahead of the tensor pipe and every barrier wait of its own becomes a pipe bubble);
// here the operands reach uniform registers once per chunk and the UTCHMMAs follow each other directly.
// acc: does the first MMA accumulate?
__device__ __forceinline__ void g2_mma_chunk(uint32_t d_tmem, uint64_t ah, uint64_t al, uint64_t bh, uint64_t bl, uint32_t idesc,
                                             uint32_t acc) {
  static_assert(G2_BK == 64, "four k steps of 16 per chunk");
  asm volatile(
      "{\n"
      ".reg .pred p, t, e;\n"
      ".reg .b64 ah1, al1, bh1, bl1, ah2, al2, bh2, bl2, ah3, al3, bh3, bl3;\n"
      "elect.sync _|e, 0xffffffff;\n"
      "setp.ne.b32 p, %6, 0;\n"
      "setp.eq.b32 t, 0, 0;\n"
      "add.u64 ah1, %1, 2;\n"
      "add.u64 al1, %2, 2;\n"
      "add.u64 bh1, %3, 2;\n"
      "add.u64 bl1, %4, 2;\n"
      "add.u64 ah2, %1, 4;\n"
      "add.u64 al2, %2, 4;\n"
      "add.u64 bh2, %3, 4;\n"
      "add.u64 bl2, %4, 4;\n"
      "add.u64 ah3, %1, 6;\n"
      "add.u64 al3, %2, 6;\n"
      "add.u64 bh3, %3, 6;\n"
      "add.u64 bl3, %4, 6;\n"
      "@e tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %3, %5, p;\n"
      "@e tcgen05.mma.cta_group::2.kind::f16 [%0], %2, %3, %5, t;\n"
      "@e tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %4, %5, t;\n"
      "@e tcgen05.mma.cta_group::2.kind::f16 [%0], ah1, bh1, %5, t;\n"
      "@e tcgen05.mma.cta_group::2.kind::f16 [%0], al1, bh1, %5, t;\n"
      "@e tcgen05.mma.cta_group::2.kind::f16 [%0], ah1, bl1, %5, t;\n"
      "@e tcgen05.mma.cta_group::2.kind::f16 [%0], ah2, bh2, %5, t;\n"
      "@e tcgen05.mma.cta_group::2.kind::f16 [%0], al2, bh2, %5, t;\n"
      "@e tcgen05.mma.cta_group::2.kind::f16 [%0], ah2, bl2, %5, t;\n"
      "@e tcgen05.mma.cta_group::2.kind::f16 [%0], ah3, bh3, %5, t;\n"
      "@e tcgen05.mma.cta_group::2.kind::f16 [%0], al3, bh3, %5, t;\n"
      "@e tcgen05.mma.cta_group::2.kind::f16 [%0], ah3, bl3, %5, t;\n"
      "}\n" ::"r"(d_tmem),
      "l"(ah), "l"(al), "l"(bh), "l"(bl), "r"(idesc), "r"(acc)
      : "memory");
}

struct G2Tile { int z, m0, n0, M; };

// Dependencies between CONSECUTIVE launches of the kernel that overlap in time (programmatic dependent launch: launch
// t + 1 starts filling SMs as the CTAs of launch t exit, it does not wait for the whole grid).  Every epilogue warp of a
// tile adds 1 to cur[z][row block] once its stores are visible; a tile of the next launch waits until the row block it
// reads - and, per problem, the row block `src_block[z]` (-1: none) other rows of its operand are copied from - has
// collected G2_DEP_DONE of them.  prev == nullptr: nothing to wait for.  cur == nullptr: nothing to publish.
struct G2Deps {
  const int* prev;         // [2][n_blocks] counters of the previous launch
  int* cur;                // [2][n_blocks] counters of this launch (zero at launch)
  const int* prev_limit;   // [2] row limits of the previous launch (row blocks at or past them had no tile there)
  int n_blocks;
  int src_block[2];
  // cur[2 n_blocks] / prev[2 n_blocks]: one more counter, over ALL tiles of the launch - once the previous launch is seen
  // complete a CTA stops looking at row blocks (an acquire load per tile is an L2 round trip in front of the TMA loads)
};
__device__ __forceinline__ int g2_ld_acquire(const int* p) {
  int v;
  asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void g2_dep_wait_one(const int* p, int want) {
  unsigned long long t0 = 0;
  for (uint32_t spin = 0; g2_ld_acquire(p) < want; ++spin) {
    __nanosleep(64);
    if ((spin & 0xfff) == 0xfff) {
      unsigned long long now;
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
      if (t0 == 0) t0 = now;
      else if (now - t0 > 4000000000ull) {
        printf("vfr: gemm_tc2 dependency wait timed out (block %d thread %d)\n", blockIdx.x, threadIdx.x);
        __trap();
      }
    }
  }
}
// rows of block `mb` of problem z (and of its copy source) written by the previous launch are visible after this
__device__ __forceinline__ void g2_dep_wait(const G2Deps& d, int z, int mb, int n_tiles_n) {
  if (d.prev == nullptr) return;
  const int done = n_tiles_n * 2 * (GT_THREADS / 32 - 2);
  const int* row = d.prev + z * d.n_blocks;
  if (mb * GT_BM < d.prev_limit[z]) g2_dep_wait_one(row + mb, done);
  const int sb = z ? d.src_block[1] : d.src_block[0];
  if (sb >= 0 && sb != mb) g2_dep_wait_one(row + sb, done);
  asm volatile("fence.proxy.async;" ::: "memory");       // what follows may be a TMA (async proxy) read of those rows
}
// has the previous launch finished altogether?  (sticky per thread: `all_done`)
__device__ __forceinline__ bool g2_dep_all_done(const G2Deps& d, int n_tiles_n, bool& all_done) {
  if (d.prev == nullptr || all_done) return true;
  const int per_tile = 2 * (GT_THREADS / 32 - 2);
  const int want = (((d.prev_limit[0] + GT_BM - 1) / GT_BM) + ((d.prev_limit[1] + GT_BM - 1) / GT_BM)) * n_tiles_n * per_tile;
  if (g2_ld_acquire(d.prev + 2 * d.n_blocks) >= want) {
    asm volatile("fence.proxy.async;" ::: "memory");
    all_done = true;
  }
  return all_done;
}
struct G2NoRow {};
template <class E, class = void>
struct G2PreOf { using type = G2NoRow; };
template <class E>
struct G2PreOf<E, std::void_t<typename E::Pre>> { using type = typename E::Pre; };
template <class E, class = void>
struct G2RowOf { using type = G2NoRow; };
template <class E>
struct G2RowOf<E, std::void_t<typename E::Row>> { using type = typename E::Row; };

// SEG: K-segmented accumulation (see gemm_tc_kernel) - every seg_chunks K chunks are one work item with its own TMEM
// buffer; the epilogue warps add item f to the fp32 partial sums of the tile in seg_buf (round-to-nearest adds; each
// thread re-reads only what it wrote itself) while the MMAs of item f + 1 fill the other buffer, and the fused epilogue
// functor runs on the last item of a tile.  Batch 1 only.
template <class Epi, bool SEG>
__global__ void __launch_bounds__(GT_THREADS, 1)
gemm_tc2_kernel(const __grid_constant__ GemmTcMaps maps, int batch, int M_all, const int* __restrict__ m_limit, int k_chunks,
                int lo_a, int lo_b, uint32_t fmt, int N_all, int seg_chunks, float* __restrict__ seg_buf, int64_t seg_ld,
                const G2Deps deps, long long* __restrict__ dbg, Epi epi) {
  extern __shared__ uint8_t g2_smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(g2_smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + G2_STAGES * G2_STAGE);
  uint64_t* full = bars;                       // [S]  (the leader's are used)
  uint64_t* empty = bars + G2_STAGES;          // [S]  per CTA
  uint64_t* acc_full = bars + 2 * G2_STAGES;   // [2]  per CTA
  uint64_t* acc_empty = acc_full + 2;          // [2]  (the leader's are used)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = g2_cta_rank();
  const bool leader = rank == 0;
  // a dependent launch (if the host asked for one) may take this CTA's SM as soon as it exits
  if (deps.cur != nullptr) asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  // development aid (VFR_GEMM_DBG = device address of int64 [8]), cycles summed over the LEADER CTAs of all pairs:
  // 0 MMA issuer waiting for a free accumulator, 1 waiting for operands, 2 its whole tile loop, 3 epilogue warp 2 waiting
  // for a complete accumulator, 4 its epilogue work, 5 producer waiting for free stages, 6 tiles, 7 pairs
  const bool trace = dbg != nullptr && leader;
  long long w0 = 0, w1 = 0;

  if (threadIdx.x == 0) {
    for (int s = 0; s < G2_STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
    for (int b = 0; b < 2; ++b) { mbar_init(&acc_full[b], 1); mbar_init(&acc_empty[b], 2 * G2_EPI_WARPS); }
    fence_barrier_init();
  }
  g2_cluster_sync();                  // the peer's barriers exist before anything signals them
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;");
  }
  gt_fence_before();
  __syncthreads();
  gt_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  // tile list: problems z = 0 .. batch-1 back to back, inside a problem N fastest; tile t of the list belongs to pair
  // t mod (pairs).  Both CTAs of a pair (and all their warps) walk the same list.
  const int n_tiles_n = (N_all + GT_BN - 1) / GT_BN;
  const int M0 = m_limit ? min(M_all, m_limit[0]) : M_all;
  const int M1 = batch > 1 ? (m_limit ? min(M_all, m_limit[1]) : M_all) : 0;
  const int tiles0 = (M0 + GT_BM - 1) / GT_BM * n_tiles_n, tiles1 = (M1 + GT_BM - 1) / GT_BM * n_tiles_n;
  const int total = tiles0 + tiles1;
  const int n_seg = SEG ? (k_chunks + seg_chunks - 1) / seg_chunks : 1;
  const int pair = blockIdx.x >> 1, n_pairs = gridDim.x >> 1;
  auto decode = [&](int t) {
    G2Tile tl;
    tl.z = t < tiles0 ? 0 : 1;
    const int r = t - (tl.z ? tiles0 : 0);
    tl.m0 = (r / n_tiles_n) * GT_BM;
    tl.n0 = (r % n_tiles_n) * GT_BN;
    tl.M = tl.z ? M1 : M0;
    return tl;
  };

  if (warp == 0) {
    // ================= TMA producer (both CTAs: own 128 rows of A, own half of the B rows) =================
    if (lane == 0) {
      const uint32_t full_leader0 = g2_mapa(&full[0], 0);     // (+ 8 s: no runtime-indexed local array)
      int it = 0;
      bool prev_done = false;
      for (int t = pair; t < total; t += n_pairs) {
        const G2Tile tl = decode(t);
        const CUtensorMap* ma = tl.z ? &maps.a[1] : &maps.a[0];
        const CUtensorMap* mb = tl.z ? &maps.b[1] : &maps.b[0];
        // the last N tile of a matrix issues narrower MMAs (N rounded up to 32); the pair splits THAT N in halves
        const int n_mma = min(GT_BN, (N_all - tl.n0 + 31) / 32 * 32);
        const int row_a = tl.m0 + (int)rank * 128, row_b = tl.n0 + (int)rank * (n_mma >> 1);
        if (!g2_dep_all_done(deps, n_tiles_n, prev_done)) g2_dep_wait(deps, tl.z, tl.m0 / GT_BM, n_tiles_n);
        for (int c = 0; c < k_chunks; ++c, ++it) {
          const int s = it % G2_STAGES;
          if (trace) { const long long t0 = clock64(); gt_wait(&empty[s], ((it / G2_STAGES) & 1) ^ 1, 32); w0 += clock64() - t0; }
          else gt_wait(&empty[s], ((it / G2_STAGES) & 1) ^ 1, 32);
          uint8_t* st = smem + s * G2_STAGE;
          if (leader) mbar_expect_tx(&full[s], 2 * G2_STAGE);
          const int kc = c * G2_BK;
          g2_tma_load(st + 0 * G2_SUB, ma, kc, row_a, full_leader0 + 8u * (uint32_t)s);          // Ah
          g2_tma_load(st + 1 * G2_SUB, ma, lo_a + kc, row_a, full_leader0 + 8u * (uint32_t)s);   // Al
          g2_tma_load(st + 2 * G2_SUB, mb, kc, row_b, full_leader0 + 8u * (uint32_t)s);          // Bh (this CTA's half of the N rows)
          g2_tma_load(st + 3 * G2_SUB, mb, lo_b + kc, row_b, full_leader0 + 8u * (uint32_t)s);   // Bl
        }
      }
      if (trace) atomicAdd(reinterpret_cast<unsigned long long*>(dbg + 5), (unsigned long long)w0);
    }
  } else if (warp == 1) {
    // ================= MMA issuer (leader CTA only) =================
    if (leader) {
      // (the whole warp runs this loop converged; the tcgen05 instructions come from its elected lane - see g2_mma_chunk)
      int it = 0, ti = 0;
      const long long t_loop = trace ? clock64() : 0;
      for (int t = pair; t < total; t += n_pairs) {
        const G2Tile tl = decode(t);
        const int n_mma = min(GT_BN, (N_all - tl.n0 + 31) / 32 * 32);
        // kind::f16, D = fp32, M = 256 (the pair), N = n_mma
        const uint32_t idesc = (1u << 4) | fmt | ((uint32_t)(n_mma >> 3) << 17) | ((uint32_t)(256 >> 4) << 24);
        int c = 0;
        for (int f = 0; f < n_seg; ++f, ++ti) {
          const int buf = ti & 1;
          if (trace) { const long long t0 = clock64(); gt_wait(&acc_empty[buf], ((ti >> 1) & 1) ^ 1, 32); w0 += clock64() - t0; }
          else gt_wait(&acc_empty[buf], ((ti >> 1) & 1) ^ 1, 32);   // both CTAs' epilogues have drained this buffer
          gt_fence_after();
          const uint32_t d = tmem_base + (uint32_t)buf * 256;
          const int c_end = SEG ? min(k_chunks, c + seg_chunks) : k_chunks;
          for (bool first = true; c < c_end; ++c, ++it, first = false) {
            const int s = it % G2_STAGES;
            if (trace) { const long long t0 = clock64(); gt_wait(&full[s], (it / G2_STAGES) & 1, 32); w1 += clock64() - t0; }
            else gt_wait(&full[s], (it / G2_STAGES) & 1, 32);
            gt_fence_after();
            uint8_t* st = smem + s * G2_STAGE;
            const uint64_t ah = g2_desc(st), al = g2_desc(st + G2_SUB), bh = g2_desc(st + 2 * G2_SUB), bl = g2_desc(st + 3 * G2_SUB);
            g2_mma_chunk(d, ah, al, bh, bl, idesc, first ? 0u : 1u);
            if (g2_elect()) g2_commit_both(&empty[s]);             // the stage is free in BOTH CTAs once these MMAs have read it
          }
          if (g2_elect()) g2_commit_both(&acc_full[buf]);
          __syncwarp();
        }
      }
      if (trace && lane == 0) {
        atomicAdd(reinterpret_cast<unsigned long long*>(dbg + 0), (unsigned long long)w0);
        atomicAdd(reinterpret_cast<unsigned long long*>(dbg + 1), (unsigned long long)w1);
        atomicAdd(reinterpret_cast<unsigned long long*>(dbg + 2), (unsigned long long)(clock64() - t_loop));
        atomicAdd(reinterpret_cast<unsigned long long*>(dbg + 6), (unsigned long long)ti);
        atomicAdd(reinterpret_cast<unsigned long long*>(dbg + 7), 1ull);
      }
    }
  } else {
    // ================= epilogue: 8 warps per CTA, two per TMEM lane quarter (128 columns each) =================
    const int quarter = warp & 3, half = (warp - 2) >> 2;
    const uint32_t acc_empty_leader0 = g2_mapa(&acc_empty[0], 0);
    bool prev_done_e = false;
    int ti = 0;
    for (int t = pair; t < total; t += n_pairs) {
      const G2Tile tl = decode(t);
      const int m = tl.m0 + (int)rank * 128 + quarter * 32 + lane;
      for (int f = 0; f < n_seg; ++f, ++ti) {
        const int buf = ti & 1;
        // (SEG: tile-major fp32 partial sums private to this kernel - a thread only re-reads what it wrote, and the 32
        //  lanes of a warp access 512 contiguous bytes per instruction.  They do not depend on the accumulator, so the
        //  first two column groups are requested before the wait and the rest two groups ahead.)
        float4* pb = nullptr;
        float4 o[2][4];
        const bool add_prev = SEG && f > 0 && m < tl.M;
        if constexpr (SEG) {
          pb = reinterpret_cast<float4*>(seg_buf) + (((int64_t)t * 2 + rank) * G2_EPI_WARPS + (warp - 2)) * 8 * 128 + lane;
          if (add_prev) {
#pragma unroll
            for (int j = 0; j < 4; ++j) { o[0][j] = pb[32 * j]; o[1][j] = pb[128 + 32 * j]; }
          }
        }
        // (functors with a `Pre`: the global operands of all 8 column groups are in flight before the accumulator is)
        typename G2PreOf<Epi>::type pre[8];
        if (deps.prev != nullptr && f == 0 && !prev_done_e) {      // (this tile's global operands come from the previous launch too)
          if (lane == 0 && !g2_dep_all_done(deps, n_tiles_n, prev_done_e)) g2_dep_wait(deps, tl.z, tl.m0 / GT_BM, n_tiles_n);
          prev_done_e = __shfl_sync(0xffffffffu, prev_done_e ? 1 : 0, 0) != 0;
        }
        if constexpr (gt_has_pre<Epi>::value && !SEG) {
          if (m < tl.M) {
#pragma unroll
            for (int c = 0; c < 8; ++c) pre[c] = epi.prefetch(tl.z, m, tl.n0 + half * 128 + c * 16);
          }
        }
        long long t_e = 0;
        if (trace && warp == 2) { const long long t0 = clock64(); gt_wait(&acc_full[buf], (ti >> 1) & 1, 128); t_e = clock64(); w0 += t_e - t0; }
        else gt_wait(&acc_full[buf], (ti >> 1) & 1, 128);
        gt_fence_after();
        const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(buf * 256 + half * 128);
        const bool last = f == n_seg - 1;
        typename G2RowOf<Epi>::type rowctx{};
        if constexpr (gt_has_row<Epi>::value) if (last && m < tl.M) rowctx = epi.row(tl.z, m);
        auto group = [&](int c, auto slot_c) {
          constexpr int slot = decltype(slot_c)::value;
          float v[16];
          gt_ld16(taddr + c * 16, v);
          const int nn = tl.n0 + half * 128 + c * 16;
          if (m < tl.M) {
            if constexpr (SEG) {
              if (add_prev) {
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                  const float4 q = o[slot][j];
                  v[4 * j] += q.x; v[4 * j + 1] += q.y; v[4 * j + 2] += q.z; v[4 * j + 3] += q.w;
                }
                if (c + 2 < 8) {
#pragma unroll
                  for (int j = 0; j < 4; ++j) o[slot][j] = pb[(c + 2) * 128 + 32 * j];
                }
              }
              if (!last) {
#pragma unroll
                for (int j = 0; j < 4; ++j) pb[c * 128 + 32 * j] = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
              }
            }
            if (last) {
              if constexpr (gt_has_row<Epi>::value) epi(tl.z, m, nn, v, rowctx);
              else if constexpr (gt_has_pre<Epi>::value && !SEG) epi(tl.z, m, nn, v, pre[c]);
              else if constexpr (gt_has_pre<Epi>::value) epi(tl.z, m, nn, v, epi.prefetch(tl.z, m, nn));
              else epi(tl.z, m, nn, v);
            }
          }
          if constexpr (gt_has_after<Epi>::value) if (last) epi.after(tl.z, m - lane, nn, lane);
        };
        if constexpr (SEG) {
#pragma unroll 1
          for (int c = 0; c < 8; c += 2) {
            group(c, std::integral_constant<int, 0>{});
            group(c + 1, std::integral_constant<int, 1>{});
          }
        } else {
#pragma unroll
          for (int c = 0; c < 8; c += 2) {          // (unrolled: pre[c] stays in registers)
            group(c, std::integral_constant<int, 0>{});
            group(c + 1, std::integral_constant<int, 1>{});
          }
        }
        gt_fence_before();
        __syncwarp();
        if (lane == 0) g2_arrive_remote(acc_empty_leader0 + 8u * (uint32_t)buf);
        if (deps.cur != nullptr && last) {
          // this warp's part of the tile is written: visible to every later reader (generic and TMA) before the count
          __threadfence();
          asm volatile("fence.proxy.async;" ::: "memory");
          __syncwarp();
          if (lane == 0) {
            asm volatile("red.release.gpu.global.add.s32 [%0], 1;" ::"l"(deps.cur + tl.z * deps.n_blocks + tl.m0 / GT_BM) : "memory");
            asm volatile("red.release.gpu.global.add.s32 [%0], 1;" ::"l"(deps.cur + 2 * deps.n_blocks) : "memory");
          }
        }
        if (trace && warp == 2) w1 += clock64() - t_e;
      }
    }
    if (trace && warp == 2 && lane == 0) {
      atomicAdd(reinterpret_cast<unsigned long long*>(dbg + 3), (unsigned long long)w0);
      atomicAdd(reinterpret_cast<unsigned long long*>(dbg + 4), (unsigned long long)w1);
    }
  }
  gt_fence_before();
  __syncthreads();
  g2_cluster_sync();                  // the peer may still read this CTA's shared memory / signal its barriers
  if (warp == 1) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512));
}

// 0 = the one-CTA kernel, 1 = the CTA-pair kernel where it applies (default)
static inline int g2_enabled() {
  const char* e = getenv("VFR_GEMM2");
  return e ? atoi(e) : 1;
}

template <class Epi, bool SEG>
static int launch_gemm_tc2_impl(const GemmTcMaps& maps, int pairs, int batch, int M, int N, int kp, Epi epi, cudaStream_t st,
                                const int* m_limit, bool f16, int lo_a, int lo_b, int seg_k, float* seg_buf, int64_t seg_ld,
                                const G2Deps& deps, bool overlap_prev) {
  VFR_CUDA(cudaFuncSetAttribute(gemm_tc2_kernel<Epi, SEG>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)G2_SMEM));
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)(2 * pairs), 1, 1);
  cfg.blockDim = dim3(GT_THREADS, 1, 1);
  cfg.dynamicSmemBytes = G2_SMEM;
  cfg.stream = st;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  // overlap_prev: this launch may start while the previous kernel of the stream is still running (its CTAs wait for their
  // operands through `deps`, the kernel never executes griddepcontrol.wait)
  attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[1].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = overlap_prev ? 2 : 1;
  VFR_CUDA(cudaLaunchKernelEx(&cfg, gemm_tc2_kernel<Epi, SEG>, maps, batch, M, m_limit, kp / G2_BK, lo_a, lo_b,
                              f16 ? GT_FMT_F16 : GT_FMT_BF16, N, seg_k / G2_BK, seg_buf, seg_ld, deps, gt_dbg_ptr(), epi));
  return check_launch("gemm_tc2_kernel");
}

template <class Epi>
static int launch_gemm_tc2(const void* const* a, const void* const* b, int batch, int M, int N, int kp, int64_t lda, int64_t ldb,
                           Epi epi, cudaStream_t st, const int* m_limit, bool f16, int lo_a, int lo_b, int seg_k,
                           float* seg_buf, int64_t seg_ld, const G2Deps* deps_in, bool overlap_prev) {
  G2Deps deps{};
  deps.src_block[0] = deps.src_block[1] = -1;
  if (deps_in) deps = *deps_in;
  GemmTcMaps maps;
  for (int z = 0; z < batch; ++z) {
    int rc = gt_make_map(&maps.a[z], a[z], (uint64_t)M, (uint64_t)lda, 128, f16, G2_BK);
    if (rc) return rc;
    rc = gt_make_map(&maps.b[z], b[z], (uint64_t)N, (uint64_t)ldb, 128, f16, G2_BK);
    if (rc) return rc;
  }
  if (batch == 1) { maps.a[1] = maps.a[0]; maps.b[1] = maps.b[0]; }
  static int n_sms = 0;
  if (!n_sms) {
    int dev = 0;
    VFR_CUDA(cudaGetDevice(&dev));
    VFR_CUDA(cudaDeviceGetAttribute(&n_sms, cudaDevAttrMultiProcessorCount, dev));
  }
  const int64_t tiles = (int64_t)batch * ((M + GT_BM - 1) / GT_BM) * ((N + GT_BN - 1) / GT_BN);
  const int pairs = (int)std::max<int64_t>(1, std::min<int64_t>(n_sms / 2, tiles));
  if (seg_buf) return launch_gemm_tc2_impl<Epi, true>(maps, pairs, batch, M, N, kp, epi, st, m_limit, f16, lo_a, lo_b, seg_k, seg_buf, seg_ld, deps, overlap_prev);
  return launch_gemm_tc2_impl<Epi, false>(maps, pairs, batch, M, N, kp, epi, st, m_limit, f16, lo_a, lo_b, 0, nullptr, 0, deps, overlap_prev);
}

}  // namespace vfr
