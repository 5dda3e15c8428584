// Split-bf16 tensor-core GEMM  C[M,N] = A[M,K] * B[N,K]^T  at fp32 accuracy (tcgen05 / TMEM / TMA).
//
// Each fp32 operand x is stored as a bf16 pair (hi = bf16(x), lo = bf16(x - hi)); a packed row is
// [hi(0..Kp) | lo(0..Kp)] bf16, Kp = K rounded up to 64.  The product is accumulated in fp32 in TMEM as
//     A.B^T ~= Ah.Bh^T + Al.Bh^T + Ah.Bl^T        (relative error ~2^-17 / sqrt(K) per dot product)
// One CTA computes a 256 x 256 output tile as two 128-row accumulators (2 x 256 TMEM columns), so a
// K-chunk of 32 needs (256 + 256) rows x (hi + lo) x 64 B = 64 KB of operands for 12 MMAs of
// 128x256x16 - the same bytes-per-MAC as a cta_group::2 pair, without a cluster.  Operands arrive by
// TMA (SWIZZLE_64B boxes) through a 3-stage mbarrier ring; one thread issues the MMAs; 8 epilogue warps
// read the accumulators (tcgen05.ld 32x32b.x16, thread = output row) and apply a fused epilogue functor.
#pragma once
#include "vfr_common.cuh"
#include <cuda.h>
#include <cuda_bf16.h>
#include <stdlib.h>
#include <type_traits>

namespace vfr {

constexpr int GT_BM = 256, GT_BN = 256, GT_BK = 32, GT_STAGES = 3;
constexpr int GT_THREADS = 320;
constexpr int GT_A_SUB = 128 * GT_BK * 2;                 // one [128 x 32] bf16 box = 8 KB
constexpr int GT_B_BOX = GT_BN * GT_BK * 2;               // one [256 x 32] bf16 box = 16 KB
constexpr int GT_STAGE = 4 * GT_A_SUB + 2 * GT_B_BOX;     // Ah0 Ah1 Al0 Al1 Bh Bl = 64 KB
constexpr uint32_t GT_SMEM = GT_STAGES * GT_STAGE + 1024 + 256;

struct GemmTcMaps {          // up to two problems per launch (blockIdx.z)
  CUtensorMap a[2];
  CUtensorMap b[2];
};

__device__ __forceinline__ void gt_tma_load(void* dst, const CUtensorMap* map, int c0, int c1, uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(
          smem_u32(dst)),
      "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void gt_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void gt_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void gt_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void gt_mma(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(d_tmem),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc)
      : "memory");
}
// K-major SWIZZLE_64B operand tile: rows of 64 B, 8-row groups 512 B apart
__device__ __forceinline__ uint64_t gt_desc(const void* smem_ptr) {
  const uint32_t addr = smem_u32(smem_ptr);
  uint64_t d = 0;
  d |= (uint64_t)((addr >> 4) & 0x3fff);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(512 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)4 << 61;                         // SWIZZLE_64B
  return d;
}
__device__ __forceinline__ void gt_wait(uint64_t* bar, uint32_t parity, unsigned ns) {
  uint32_t done = 0;
  unsigned long long t0 = 0;
  for (uint32_t spin = 0;; ++spin) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(done)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    if (done) return;
    __nanosleep(ns);
    if ((spin & 0xfff) == 0xfff) {
      unsigned long long now;
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
      if (t0 == 0) t0 = now;
      else if (now - t0 > 4000000000ull) {
        printf("vfr: gemm_tc mbarrier wait timed out (block %d,%d,%d thread %d)\n", blockIdx.x, blockIdx.y, blockIdx.z,
               threadIdx.x);
        __trap();
      }
    }
  }
}
__device__ __forceinline__ void gt_ld16(uint32_t taddr, float (&v)[16]) {
  uint32_t r[16];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}

// the same load split in two: issue (the registers are NOT valid yet) and wait (which also orders every later use of
// the registers after the wait) - global loads of the epilogue can be put in flight between the two
__device__ __forceinline__ void gt_ld16_issue(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr));
}
__device__ __forceinline__ void gt_ld16_wait(uint32_t (&r)[16]) {
  asm volatile("tcgen05.wait::ld.sync.aligned;"
               : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]), "+r"(r[8]),
                 "+r"(r[9]), "+r"(r[10]), "+r"(r[11]), "+r"(r[12]), "+r"(r[13]), "+r"(r[14]), "+r"(r[15])
               :
               : "memory");
}
// an epilogue functor that declares `struct Pre` and `Pre prefetch(z, m, n0) const` gets its global operands of column
// group c + 1 requested while group c is computed:  operator()(z, m, n0, v, pre)
template <class E, class = void>
struct gt_has_pre : std::false_type {};
template <class E>
struct gt_has_pre<E, std::void_t<typename E::Pre>> : std::true_type {};

// an epilogue functor that declares `struct Row` and `Row row(z, m) const` gets its per-row operands (scales, gathered
// pointers ...) computed once per tile row instead of once per 16-column group:  operator()(z, m, n0, v, row)
// (CTA-pair kernel only)
template <class E, class = void>
struct gt_has_row : std::false_type {};
template <class E>
struct gt_has_row<E, std::void_t<typename E::Row>> : std::true_type {};

// an epilogue functor that declares `struct After` gets  after(z, m_warp, n0, lane)  called by ALL lanes of the warp after
// every 16-column group of the final item (m_warp = the warp's first row; rows past M included) - warp-cooperative work
// that follows from the group just written (CTA-pair kernel only)
template <class E, class = void>
struct gt_has_after : std::false_type {};
template <class E>
struct gt_has_after<E, std::void_t<typename E::After>> : std::true_type {};

constexpr uint32_t GT_FMT_BF16 = (1u << 7) | (1u << 10);
constexpr uint32_t GT_FMT_F16 = 0u;

// split an fp32 value into its bf16 hi / lo pair
__device__ __forceinline__ void split2(float x, __nv_bfloat16& hi, __nv_bfloat16& lo) {
  hi = __float2bfloat16_rn(x);
  lo = __float2bfloat16_rn(x - __bfloat162float(hi));
}

// Epi: struct with  __device__ void operator()(int z, int m, int n0, const float (&v)[16]) const
//      called for 16 consecutive output columns n0..n0+15 of row m (m < M guaranteed, columns not).
template <class Epi>
__global__ void __launch_bounds__(GT_THREADS, 1)
gemm_tc_kernel(const __grid_constant__ GemmTcMaps maps, int M_all, const int* __restrict__ m_limit, int k_chunks, int lo_a,
               int lo_b, uint32_t fmt, int flush_chunks, float* __restrict__ flush_buf, int64_t flush_ld, int N_all, int pre_mode, long long* __restrict__ dbg, Epi epi) {
  extern __shared__ uint8_t gt_smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(gt_smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + GT_STAGES * GT_STAGE);
  uint64_t* full = bars;
  uint64_t* empty = bars + GT_STAGES;
  uint64_t* acc_full = bars + 2 * GT_STAGES;
  uint64_t* acc_empty = acc_full + 1;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 1);
  // K-segmented accumulation (flush_chunks > 0): the tensor core adds into its fp32 accumulator with truncation, a bias
  // that grows with the number of accumulation steps (~1.5e-5 relative at K = 8192).  Every flush_chunks K-chunks the
  // epilogue warps therefore drain the accumulator into an fp32 buffer (round-to-nearest adds) and the MMAs restart from
  // zero; the fused epilogue functor sees the sum of the segments.
  const int n_seg = (flush_chunks > 0) ? (k_chunks + flush_chunks - 1) / flush_chunks : 1;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int z = blockIdx.z;
  const int m0 = blockIdx.y * GT_BM, n0 = blockIdx.x * GT_BN;
  // development aid (VFR_GEMM_DBG = device address of int64 [8]): cycles summed over all tiles that do work -
  // 0 set-up (barriers, TMEM), 1 wait for the first operands, 2 main loop (first operands -> last MMA issued),
  // 3 epilogue (accumulator complete -> last store, one epilogue warp), 4 whole tile, 7 number of tiles
  const bool trace = dbg != nullptr;
  long long t_a = 0, t_b = 0;
  if (trace) t_a = clock64();
  const long long t_begin = t_a;
  // rows of problem z that are live in this launch (device-side count: no host sync); a tile past it retires
  const int M = m_limit ? min(M_all, m_limit[z]) : M_all;
  if (m0 >= M) return;
  const CUtensorMap* ma = &maps.a[z];
  const CUtensorMap* mb = &maps.b[z];

  if (threadIdx.x == 0) {
    for (int s = 0; s < GT_STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
    mbar_init(acc_full, 1);
    mbar_init(acc_empty, GT_THREADS / 32 - 2);
    fence_barrier_init();
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  gt_fence_before();
  __syncthreads();
  gt_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  if (trace && threadIdx.x == 32) { t_b = clock64(); atomicAdd(reinterpret_cast<unsigned long long*>(dbg), (unsigned long long)(t_b - t_a)); }

  if (warp == 0) {
    if (lane == 0) {
      for (int c = 0; c < k_chunks; ++c) {
        const int s = c % GT_STAGES;
        gt_wait(&empty[s], ((c / GT_STAGES) & 1) ^ 1, 64);
        uint8_t* st = smem + s * GT_STAGE;
        mbar_expect_tx(&full[s], GT_STAGE);
        const int kc = c * GT_BK;
        gt_tma_load(st + 0 * GT_A_SUB, ma, kc, m0, &full[s]);              // Ah rows m0..+127
        gt_tma_load(st + 1 * GT_A_SUB, ma, kc, m0 + 128, &full[s]);        // Ah rows m0+128..
        gt_tma_load(st + 2 * GT_A_SUB, ma, lo_a + kc, m0, &full[s]);       // Al
        gt_tma_load(st + 3 * GT_A_SUB, ma, lo_a + kc, m0 + 128, &full[s]);
        gt_tma_load(st + 4 * GT_A_SUB, mb, kc, n0, &full[s]);              // Bh
        gt_tma_load(st + 4 * GT_A_SUB + GT_B_BOX, mb, lo_b + kc, n0, &full[s]);  // Bl
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      // kind::f16: D = fp32; A / B formats from `fmt` (bits 7 / 10: 1 = bf16, 0 = fp16 - the split-fp16 operands)
      // (the last N tile of a matrix whose N is not a multiple of 256 issues narrower MMAs: N rounded up to 16 -
      //  4H = 4000 gate rows leave 160 of the 16th tile's 256 columns)
      const int n_mma = min(GT_BN, (N_all - n0 + 15) / 16 * 16);
      const uint32_t idesc = (1u << 4) | fmt | ((uint32_t)(n_mma >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
      int seg = 0;
      for (int c = 0; c < k_chunks; ++c) {
        const int s = c % GT_STAGES;
        const int cs = (flush_chunks > 0) ? c % flush_chunks : c;     // chunk index inside its accumulation segment
        if (cs == 0 && c > 0) {                                       // the epilogue has drained the previous segment
          gt_wait(acc_empty, (seg - 1) & 1, 32);
          gt_fence_after();
        }
        gt_wait(&full[s], (c / GT_STAGES) & 1, 32);
        gt_fence_after();
        if (trace && c == 0) { t_a = clock64(); atomicAdd(reinterpret_cast<unsigned long long*>(dbg + 1), (unsigned long long)(t_a - t_b)); }
        uint8_t* st = smem + s * GT_STAGE;
        const uint64_t bh = gt_desc(st + 4 * GT_A_SUB), bl = gt_desc(st + 4 * GT_A_SUB + GT_B_BOX);
#pragma unroll
        for (int sub = 0; sub < 2; ++sub) {
          const uint64_t ah = gt_desc(st + sub * GT_A_SUB), al = gt_desc(st + (2 + sub) * GT_A_SUB);
          const uint32_t d = tmem_base + (uint32_t)sub * 256;
#pragma unroll
          for (int k = 0; k < GT_BK / 16; ++k) {
            gt_mma(d, ah + 2 * k, bh + 2 * k, idesc, (cs | k) ? 1u : 0u);
            gt_mma(d, al + 2 * k, bh + 2 * k, idesc, 1u);
            gt_mma(d, ah + 2 * k, bl + 2 * k, idesc, 1u);
          }
        }
        gt_commit(&empty[s]);
        if (c == k_chunks - 1 || (flush_chunks > 0 && cs == flush_chunks - 1)) {
          gt_commit(acc_full);
          ++seg;
        }
      }
      if (trace) atomicAdd(reinterpret_cast<unsigned long long*>(dbg + 2), (unsigned long long)(clock64() - t_a));
    }
  } else {
    const int ew = warp - 2;
    const int quarter = warp & 3;
    const int sub = ew >> 2;
    const int m = m0 + sub * 128 + quarter * 32 + lane;
    const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)sub * 256;
    bool piped = false;
    if constexpr (gt_has_pre<Epi>::value) piped = pre_mode != 0;
    if constexpr (gt_has_pre<Epi>::value) if (piped) {
      // software-pipelined epilogue (single accumulation segment, opt-in: see gt_pre_mode): the functor's global operands
      // of the next column group and the TMEM load of this one are in flight together
      typename Epi::Pre cur{}, nxt{};
      if (m < M) cur = epi.prefetch(z, m, n0);
      gt_wait(acc_full, 0, 256);
      gt_fence_after();
      for (int c = 0; c < GT_BN / 16; ++c) {
        uint32_t r[16];
        gt_ld16_issue(taddr + c * 16, r);
        if (c + 1 < GT_BN / 16 && m < M) nxt = epi.prefetch(z, m, n0 + (c + 1) * 16);
        gt_ld16_wait(r);
        float v[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
        if (m < M) epi(z, m, n0 + c * 16, v, cur);
        cur = nxt;
      }
    }
    if (!piped)
    for (int f = 0; f < n_seg; ++f) {
      gt_wait(acc_full, f & 1, 256);
      gt_fence_after();
      if (trace && f == n_seg - 1 && threadIdx.x == 64) t_b = clock64();
      const bool last = f == n_seg - 1;
      for (int c = 0; c < GT_BN / 16; ++c) {
        float v[16];
        gt_ld16(taddr + c * 16, v);
        if (n_seg > 1 && m < M) {
          float4* pb = reinterpret_cast<float4*>(flush_buf + (int64_t)m * flush_ld + n0 + c * 16);
          if (f > 0) {
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const float4 o = pb[j];
              v[4 * j] += o.x; v[4 * j + 1] += o.y; v[4 * j + 2] += o.z; v[4 * j + 3] += o.w;
            }
          }
          if (!last) {
#pragma unroll
            for (int j = 0; j < 4; ++j) pb[j] = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
          }
        }
        if (last && m < M) {
          if constexpr (gt_has_pre<Epi>::value) epi(z, m, n0 + c * 16, v, epi.prefetch(z, m, n0 + c * 16));
          else epi(z, m, n0 + c * 16, v);
        }
      }
      if (!last) {                       // the accumulator has been read: the MMAs of the next segment may overwrite it
        gt_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(acc_empty);
      }
    }
  }
  if (trace && threadIdx.x == 64) atomicAdd(reinterpret_cast<unsigned long long*>(dbg + 3), (unsigned long long)(clock64() - t_b));
  gt_fence_before();
  __syncthreads();
  if (trace && threadIdx.x == 0) {
    atomicAdd(reinterpret_cast<unsigned long long*>(dbg + 4), (unsigned long long)(clock64() - t_begin));
    atomicAdd(reinterpret_cast<unsigned long long*>(dbg + 7), 1ull);
  }
  if (warp == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512));
}

// ---- host helpers ---------------------------------------------------------------------------------
typedef CUresult (*GtEncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                               const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                               CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static inline GtEncodeFn gt_encode_fn() {
  static GtEncodeFn fn = nullptr;
  if (!fn) {
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<GtEncodeFn>(ptr);
  }
  return fn;
}

// packed operand [rows, ld_elems] bf16 (ld_elems >= 2*kp); box = [32 k] x [box_rows]; rows beyond `rows` read as 0
static inline int gt_make_map(CUtensorMap* map, const void* base, uint64_t rows, uint64_t ld_elems, uint32_t box_rows,
                              bool f16 = false, uint32_t box_k = GT_BK) {
  GtEncodeFn enc = gt_encode_fn();
  VFR_REQUIRE(enc, VFR_ERR_CUDA, "cuTensorMapEncodeTiled entry point not available");
  cuuint64_t gdim[2] = {(cuuint64_t)ld_elems, (cuuint64_t)rows};
  cuuint64_t gstride[1] = {(cuuint64_t)ld_elems * 2};
  cuuint32_t box[2] = {box_k, box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(map, f16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base),
                   gdim, gstride, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, box_k * 2 == 128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  VFR_REQUIRE(r == CUDA_SUCCESS, VFR_ERR_CUDA, "cuTensorMapEncodeTiled failed with CUresult %d", (int)r);
  return VFR_OK;
}

// VFR_GEMM_PRE=1 turns the software-pipelined epilogue ON.  Measured on B200 (tools/k3_ab.py, K3 of 37 888 queries, same
// process): pipelined 35.2 - 36.4 ms, plain 34.35 ms - the epilogue warps already overlap one another's global loads, and
// the extra live registers cost more than the hidden latency gains.  Off by default; kept for shapes with fewer warps.
static inline int gt_pre_mode() {
  const char* e = getenv("VFR_GEMM_PRE");
  return e ? atoi(e) : 0;
}

static inline long long* gt_dbg_ptr() {
  const char* e = getenv("VFR_GEMM_DBG");
  return e ? reinterpret_cast<long long*>(strtoull(e, nullptr, 0)) : nullptr;
}

// packed operand width: K rounded up to 64 (one K chunk of the CTA-pair kernel; two of the one-CTA kernel)
static inline int gt_kp(int k) { return (k + 63) / 64 * 64; }

struct G2Deps;
template <class Epi>
static int launch_gemm_tc2(const void* const* a, const void* const* b, int batch, int M, int N, int kp, int64_t lda, int64_t ldb,
                           Epi epi, cudaStream_t st, const int* m_limit, bool f16, int lo_a, int lo_b, int seg_k,
                           float* seg_buf, int64_t seg_ld, const G2Deps* deps_in, bool overlap_prev);
static inline int g2_enabled();

// A: packed [M, lda] (lda >= 2*kp), B: packed [N, ldb]; batch <= 2 problems with identical shapes.
// lo_a / lo_b (default kp): column distance between the hi and the lo half of a row of A / B - a K-segment of a wider
// packed operand is used by passing its first column's address, its own kp and the wide operand's hi -> lo distance.
// f16: the operands are split-fp16 (pre-scaled by powers of two into fp16's range) instead of split-bf16.
template <class Epi>
static int launch_gemm_tc(const void* const* a, const void* const* b, int batch, int M, int N, int kp, int64_t lda,
                          int64_t ldb, Epi epi, cudaStream_t st, const int* m_limit = nullptr, bool f16 = false,
                          int lo_a = 0, int lo_b = 0, int flush_k = 0, float* flush_buf = nullptr, int64_t flush_ld = 0) {
  if (M <= 0 || N <= 0 || kp <= 0) return VFR_OK;
  if (lo_a == 0) lo_a = kp;
  if (lo_b == 0) lo_b = kp;
  VFR_REQUIRE(batch >= 1 && batch <= 2 && kp % GT_BK == 0 && lda >= lo_a + kp && ldb >= lo_b + kp && lda % 8 == 0 &&
                  ldb % 8 == 0 && lo_a % 8 == 0 && lo_b % 8 == 0,
              VFR_ERR_INVALID, "launch_gemm_tc: bad operand layout");
  // K-segmented accumulation: flush_buf fp32 [>= M rounded up to 256 rows, flush_ld = N rounded up to 256] (batch 1 only),
  // 16-byte aligned (the CTA-pair kernel uses it tile by tile: 256 KB per 256 x 256 tile)
  VFR_REQUIRE(!flush_buf || (batch == 1 && flush_k >= GT_BK && flush_k % GT_BK == 0 && flush_ld % 4 == 0 &&
                             flush_ld >= (N + GT_BN - 1) / GT_BN * GT_BN),
              VFR_ERR_INVALID, "launch_gemm_tc: bad flush buffer");
  // the persistent CTA-pair kernel (vfr_gemm_tc2.cuh) serves every operand whose K is a whole number of its 64-column chunks
  if (kp % 64 == 0 && (!flush_buf || flush_k % 64 == 0) && g2_enabled())
    return launch_gemm_tc2(a, b, batch, M, N, kp, lda, ldb, epi, st, m_limit, f16, lo_a, lo_b, flush_buf ? flush_k : 0, flush_buf,
                           flush_ld, nullptr, false);
  GemmTcMaps maps;
  for (int z = 0; z < batch; ++z) {
    int rc = gt_make_map(&maps.a[z], a[z], (uint64_t)M, (uint64_t)lda, 128, f16);
    if (rc) return rc;
    rc = gt_make_map(&maps.b[z], b[z], (uint64_t)N, (uint64_t)ldb, GT_BN, f16);
    if (rc) return rc;
  }
  if (batch == 1) { maps.a[1] = maps.a[0]; maps.b[1] = maps.b[0]; }
  VFR_CUDA(cudaFuncSetAttribute(gemm_tc_kernel<Epi>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)GT_SMEM));
  dim3 grid((N + GT_BN - 1) / GT_BN, (M + GT_BM - 1) / GT_BM, batch);
  gemm_tc_kernel<Epi><<<grid, GT_THREADS, GT_SMEM, st>>>(maps, M, m_limit, kp / GT_BK, lo_a, lo_b, f16 ? GT_FMT_F16 : GT_FMT_BF16,
                                                         flush_buf ? flush_k / GT_BK : 0, flush_buf, flush_ld, N, gt_pre_mode(), gt_dbg_ptr(), epi);
  return check_launch("gemm_tc_kernel");
}

}  // namespace vfr

#include "vfr_gemm_tc2.cuh"
