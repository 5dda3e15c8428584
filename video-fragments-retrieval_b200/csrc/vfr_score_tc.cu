// K4 (tensor-core path): query x clip squared-L2 as a split-bf16 GEMM on tcgen05, fused with the
// distance -> moment-mean -> top-k epilogue.  Same contract as vfr_score_topk / vfr_score_full
// (reference model/evaluate.py:49-58,71-80) at the north-star tolerance (fp32 scores within 1e-5).
//
// Arithmetic.  With q = qh + ql, v' = -2v = vh + vl (bf16 hi/lo pairs, 16 mantissa bits each side):
//     ||v - q + eps||^2 = nq + [ nv - 2 q.v ]        nq = |q|^2 - 2 eps sum(q)
//                                                    nv = |v|^2 + 2 eps sum(v) + D eps^2
//     nv - 2 q.v  ~=  qh.vh + ql.vh + qh.vl + (1,1,1).(nv0,nv1,nv2)
// i.e. three bf16 MMAs accumulated in fp32 in TMEM; nv rides along as three extra K columns (bf16
// triple = 24 bits) so the epilogue gets `nv - 2 q.v` straight from the accumulator.  Dropped term
// ql.vl and the bf16x2 representation error are ~2^-17 |q||v| / sqrt(D): ~1e-6 of d^2 unless the
// expansion cancels (near-duplicate embeddings).  Those pairs (d^2 < |q|^2 / 2) are recomputed with
// the exact fp32 direct-difference form of vfr_score.cu, so the result keeps fp32 accuracy there too.
//
// Pipeline (one CTA per SM, 320 threads, warp-specialised):
//   warp 0      TMA producer : A (128 queries x [hi|lo] x 128 k, resident) once, then B chunks
//               [240 clip slots x 64 k] (30 KB, SWIZZLE_128B) through a 5-stage mbarrier ring
//   warp 1      MMA issuer   : tcgen05.mma cta_group::1 kind::f16, M=128, N=240, K=16 per instruction,
//               21 instructions per tile (D=100), accumulators double-buffered in TMEM (2 x 256 cols)
//   warps 2..9  epilogue     : tcgen05.ld 32x32b -> one thread = one query row; +nq, clamp, sqrt.approx,
//               prefix sums of a video's 6 clip distances, 21 compares against tau*len, rare append
//               to the thread-private candidate list (vfr_topk.cuh).  Two warps share a TMEM lane
//               quarter and split the tile's 40 videos.
// Bank layout: fixed slots of S=6 clips per video (shorter videos zero-padded and masked), rows of
// 256 bf16 = [hi segment 128 | lo segment 128]; tile t = slots [240 t, 240 t + 240).
#include "vfr_common.cuh"
#include "vfr_topk.cuh"
#include <cuda.h>
#include <cuda_bf16.h>
#include <math_constants.h>

namespace vfr {

constexpr int TC_S = 6;                 // clip slots per video
constexpr int TC_N = 240;               // clip slots (MMA N) per tile
constexpr int TC_VID = TC_N / TC_S;     // 40 videos per tile
constexpr int TC_M = 128;               // queries per tile
constexpr int TC_KSEG = 128;            // bf16 columns per segment (hi / lo)
constexpr int TC_ROW = 2 * TC_KSEG;     // 256 bf16 per packed row
constexpr int TC_STAGES = 5;
constexpr int TC_A_CHUNK = TC_M * 128;  // bytes of one [128 x 64 bf16] A chunk
constexpr int TC_B_CHUNK = TC_N * 128;  // bytes of one [240 x 64 bf16] B chunk = 30720
constexpr int TC_THREADS = 320;
constexpr int TC_EPI_WARPS = 8;
constexpr int TC_SUB = 24;              // columns per epilogue sub-block (4 videos)
constexpr int TC_SUBS = (TC_N / 2) / TC_SUB;   // 5 sub-blocks per column half
constexpr int TC_CAP_HI = CAP - (TC_SUB / TC_S) * 21;  // at most 84 appends between compaction checks
constexpr uint32_t TC_SMEM = 4 * TC_A_CHUNK + TC_STAGES * TC_B_CHUNK + 1024 /*align*/ + 256 /*barriers*/;

enum TcMode { TC_TOPK = 0, TC_FULL = 1 };

struct TcParams {
  const float* nq;            // [Qpad]  |q|^2 - 2 eps sum(q)
  const float* gq;            // [Qpad]  guard threshold |q|^2 / 2
  const uint8_t* nseg;        // [Vpad]  clips per video (0 for padding videos)
  int uniform;                // every video has exactly TC_S clips
  int64_t n_videos;
  int64_t n_queries;
  int n_qtiles;
  int n_tiles;
  int tiles_per_split;
  int ksteps;                 // 16-wide k steps per segment (ceil((D+3)/16))
  int n_terms;                // 3 = split-bf16 (fp32 accuracy), 1 = plain bf16
  // exact fallback operands
  const float* bank;          // fp32 [C, D]
  const float* queries;       // fp32 [Q, D]
  const int32_t* vid_off;     // [V+1]
  const int64_t* mom_off;     // [V+1]
  int dim;
  // TOPK
  int k;
  unsigned long long* cand;
  int32_t* cand_cnt;
  int n_parts;
  unsigned* tau_g;            // [Qpad] shared per-query threshold
  // FULL
  float* out_full;
  int64_t m_total;
};

// ---------------------------------------------------------------------------------------------
// packing kernels
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void split3(float x, __nv_bfloat16& a, __nv_bfloat16& b, __nv_bfloat16& c) {
  a = __float2bfloat16_rn(x);
  const float r1 = x - __bfloat162float(a);
  b = __float2bfloat16_rn(r1);
  c = __float2bfloat16_rn(r1 - __bfloat162float(b));
}

// one warp per clip slot
__global__ void tc_pack_bank_kernel(const float* __restrict__ bank, const int32_t* __restrict__ vid_off,
                                    int64_t n_videos, int64_t n_slots_pad, int dim, int n_terms,
                                    __nv_bfloat16* __restrict__ out, uint8_t* __restrict__ nseg) {
  const int lane = threadIdx.x & 31;
  const int64_t slot = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (slot >= n_slots_pad) return;
  const int64_t v = slot / TC_S;
  const int c = (int)(slot % TC_S);
  int n = 0;
  if (v < n_videos) n = vid_off[v + 1] - vid_off[v];
  if (c == 0 && lane == 0) nseg[v] = (uint8_t)n;
  __nv_bfloat16* row = out + slot * TC_ROW;
  const __nv_bfloat16 zero = __float2bfloat16_rn(0.f);
  if (c >= n) {
    for (int k = lane; k < TC_ROW; k += 32) row[k] = zero;
    return;
  }
  const float* src = bank + (int64_t)(vid_off[v] + c) * dim;
  double ss = 0.0, sm = 0.0;
  for (int k = lane; k < TC_KSEG; k += 32) {
    __nv_bfloat16 hi = zero, lo = zero;
    if (k < dim) {
      const float x = src[k];
      ss += (double)x * x;
      sm += (double)x;
      const float y = -2.f * x;
      hi = __float2bfloat16_rn(y);
      if (n_terms == 3) lo = __float2bfloat16_rn(y - __bfloat162float(hi));
    }
    if (k < dim || k >= dim + 3) row[k] = hi;   // columns dim..dim+2 (the nv triple) are written by lane 0 below
    row[TC_KSEG + k] = lo;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    ss += __shfl_xor_sync(0xffffffffu, ss, o);
    sm += __shfl_xor_sync(0xffffffffu, sm, o);
  }
  if (lane == 0) {
    const double eps = (double)VFR_PAIRWISE_EPS;
    const float nv = (float)(ss + 2.0 * eps * sm + (double)dim * eps * eps);
    __nv_bfloat16 a, b, c3;
    split3(nv, a, b, c3);
    row[dim] = a;
    row[dim + 1] = b;
    row[dim + 2] = c3;
  }
}

// one warp per query row
__global__ void tc_pack_query_kernel(const float* __restrict__ q, int64_t n_queries, int64_t n_rows_pad, int dim,
                                     int n_terms, __nv_bfloat16* __restrict__ out, float* __restrict__ nq,
                                     float* __restrict__ gq) {
  const int lane = threadIdx.x & 31;
  const int64_t r = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (r >= n_rows_pad) return;
  __nv_bfloat16* row = out + r * TC_ROW;
  const __nv_bfloat16 zero = __float2bfloat16_rn(0.f);
  if (r >= n_queries) {
    for (int k = lane; k < TC_ROW; k += 32) row[k] = zero;
    if (lane == 0) { nq[r] = 0.f; gq[r] = 0.f; }
    return;
  }
  const float* src = q + r * dim;
  double ss = 0.0, sm = 0.0;
  for (int k = lane; k < TC_KSEG; k += 32) {
    __nv_bfloat16 hi = zero, lo = zero;
    if (k < dim) {
      const float x = src[k];
      ss += (double)x * x;
      sm += (double)x;
      hi = __float2bfloat16_rn(x);
      if (n_terms == 3) lo = __float2bfloat16_rn(x - __bfloat162float(hi));
    } else if (k < dim + 3) {
      hi = __float2bfloat16_rn(1.f);   // multiplies the nv triple of the bank row
    }
    row[k] = hi;
    row[TC_KSEG + k] = lo;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    ss += __shfl_xor_sync(0xffffffffu, ss, o);
    sm += __shfl_xor_sync(0xffffffffu, sm, o);
  }
  if (lane == 0) {
    nq[r] = (float)(ss - 2.0 * (double)VFR_PAIRWISE_EPS * sm);
    gq[r] = (n_terms == 3) ? (float)(0.5 * ss) : -1.f;   // plain-bf16 mode has no exact fallback
  }
}

// ---------------------------------------------------------------------------------------------
// tcgen05 / TMA primitives
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* map, int c0, int c1, uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(
          smem_u32(dst)),
      "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}
// mbarrier wait with a watchdog: a protocol bug traps (launch failure) instead of hanging the GPU
__device__ __forceinline__ void mbar_wait_wd(uint64_t* bar, uint32_t parity) {
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
    if ((spin & 0x3ff) == 0x3ff) {
      unsigned long long now;
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
      if (t0 == 0) t0 = now;
      else if (now - t0 > 4000000000ull) {   // 4 s
        printf("vfr: mbarrier wait timed out (block %d thread %d)\n", blockIdx.x, threadIdx.x);
        __trap();
      }
    }
  }
}
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void tc_mma(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(d_tmem),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc)
      : "memory");
}
// K-major, SWIZZLE_128B operand tile: rows of 128 B, 8-row groups 1024 B apart
__device__ __forceinline__ uint64_t make_desc(const void* smem_ptr) {
  const uint32_t addr = smem_u32(smem_ptr);
  uint64_t d = 0;
  d |= (uint64_t)((addr >> 4) & 0x3fff);          // start address
  d |= (uint64_t)1 << 16;                         // leading byte offset (ignored for swizzled K-major)
  d |= (uint64_t)(1024 >> 4) << 32;               // stride byte offset between 8-row groups
  d |= (uint64_t)1 << 46;                         // descriptor version (Blackwell)
  d |= (uint64_t)2 << 61;                         // SWIZZLE_128B
  return d;
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float (&v)[24], int off) {
  uint32_t r[16];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr));
#pragma unroll
  for (int i = 0; i < 16; ++i) v[off + i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, float (&v)[24], int off) {
  uint32_t r[8];
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
               : "r"(taddr));
#pragma unroll
  for (int i = 0; i < 8; ++i) v[off + i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ float sqrt_approx(float x) {
  float y;
  asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// exact fp32 distance of one (query, clip) pair: the arithmetic of vfr_score.cu (direct-difference form)
__device__ __noinline__ float exact_distance(const float* __restrict__ vr, const float* __restrict__ qr, int dim) {
  float acc = 0.f, sv = 0.f, sq = 0.f;
  for (int k = 0; k < dim; ++k) {
    const float d = __fsub_rn(vr[k], qr[k]);
    acc = __fmaf_rn(d, d, acc);
  }
  for (int k = 0; k < dim; ++k) sv = __fadd_rn(sv, vr[k]);
  for (int k = 0; k < dim; ++k) sq = __fadd_rn(sq, qr[k]);
  const float corr = __fmaf_rn(2.f * VFR_PAIRWISE_EPS, __fsub_rn(sv, sq), (float)dim * VFR_PAIRWISE_EPS * VFR_PAIRWISE_EPS);
  return __fsqrt_rn(fmaxf(__fadd_rn(acc, corr), 0.f));
}

// cold paths of the epilogue, kept out of line so the hot loop stays small in the instruction cache
__device__ __noinline__ int append_video(unsigned long long* list, int cnt, float tau, int n, unsigned mbase, float d0,
                                         float d1, float d2, float d3, float d4, float d5) {
  const float d[TC_S] = {d0, d1, d2, d3, d4, d5};
  const float rcp[TC_S] = {1.f, 0.5f, 1.f / 3.f, 0.25f, 0.2f, 1.f / 6.f};
#pragma unroll
  for (int s = 0; s < TC_S; ++s) {
    float run = 0.f;
#pragma unroll
    for (int e = s; e < TC_S; ++e) {
      run += d[e];
      const float score = run * rcp[e - s];
      if (e < n && score <= tau)
        list[cnt++] = ((unsigned long long)__float_as_uint(score) << 32) | (mbase + (unsigned)moment_index(n, s, e));
    }
  }
  return cnt;
}

__device__ __noinline__ void write_video_scores(float* o, int n, float d0, float d1, float d2, float d3, float d4,
                                                float d5) {
  const float d[TC_S] = {d0, d1, d2, d3, d4, d5};
  const float rcp[TC_S] = {1.f, 0.5f, 1.f / 3.f, 0.25f, 0.2f, 1.f / 6.f};
#pragma unroll
  for (int s = 0; s < TC_S; ++s) {
    float run = 0.f;
#pragma unroll
    for (int e = s; e < TC_S; ++e) {
      run += d[e];
      if (e < n) o[moment_index(n, s, e)] = run * rcp[e - s];
    }
  }
}

__device__ __noinline__ void compact_lists_noinline(unsigned long long* list, int& cnt, float& tau, int k, bool need,
                                                    int lane) {
  compact_lists(list, cnt, tau, k, need, lane);
}

// mbarrier wait with nanosleep back-off (single-lane producer / MMA threads must not steal issue slots
// from the epilogue warps of their SM sub-partition) and a watchdog: a protocol bug traps instead of hanging
__device__ __forceinline__ void mbar_wait_sleep(uint64_t* bar, uint32_t parity, unsigned ns) {
  // try_wait with a suspend-time hint: the thread sleeps in hardware until the phase completes (or the
  // hint expires), so single-lane producer / MMA threads do not steal issue slots from the epilogue warps
  uint32_t done = 0;
  unsigned long long t0 = 0;
  (void)ns;
  for (uint32_t spin = 0;; ++spin) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(done)
        : "r"(smem_u32(bar)), "r"(parity), "r"(1000000u)
        : "memory");
    if (done) return;
    if ((spin & 0x3f) == 0x3f) {
      unsigned long long now;
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
      if (t0 == 0) t0 = now;
      else if (now - t0 > 4000000000ull) {   // 4 s
        printf("vfr: mbarrier wait timed out (block %d thread %d)\n", blockIdx.x, threadIdx.x);
        __trap();
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------
// the kernel
// ---------------------------------------------------------------------------------------------
template <int MODE, bool SPLIT>
__global__ void __launch_bounds__(TC_THREADS, 1)
score_tc_kernel(const __grid_constant__ CUtensorMap tm_a, const __grid_constant__ CUtensorMap tm_b, const TcParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* smem_a = smem;                                   // 4 chunks: hi0, hi1, lo0, lo1
  uint8_t* smem_b = smem + 4 * TC_A_CHUNK;                  // TC_STAGES chunks
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_b + TC_STAGES * TC_B_CHUNK);
  uint64_t* full = bars;                        // [TC_STAGES]
  uint64_t* empty = bars + TC_STAGES;           // [TC_STAGES]
  uint64_t* a_full = bars + 2 * TC_STAGES;      // [1]
  uint64_t* tmem_full = a_full + 1;             // [2]
  uint64_t* tmem_empty = tmem_full + 2;         // [2]
  uint32_t* tmem_base_slot = reinterpret_cast<uint32_t*>(tmem_empty + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int qtile = blockIdx.x % p.n_qtiles;
  const int split = blockIdx.x / p.n_qtiles;
  const int tile_begin = split * p.tiles_per_split;
  const int tile_end = min(tile_begin + p.tiles_per_split, p.n_tiles);
  const int n_my_tiles = max(tile_end - tile_begin, 0);
  const int seg_chunks = (p.ksteps > 4) ? 2 : 1;           // 64-column chunks per segment
  const int b_chunks = seg_chunks * (p.n_terms == 3 ? 2 : 1);

  if (threadIdx.x == 0) {
    for (int s = 0; s < TC_STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
    mbar_init(a_full, 1);
    for (int b = 0; b < 2; ++b) { mbar_init(&tmem_full[b], 1); mbar_init(&tmem_empty[b], TC_EPI_WARPS); }
    fence_barrier_init();
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_base_slot)),
                 "r"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_base_slot;

  if (warp == 0) {
    // ================= TMA producer =================
    if (lane == 0 && n_my_tiles > 0) {
      const int a_chunks = seg_chunks * (p.n_terms == 3 ? 2 : 1);
      mbar_expect_tx(a_full, (uint32_t)a_chunks * TC_A_CHUNK);
      for (int s = 0; s < (p.n_terms == 3 ? 2 : 1); ++s)
        for (int c = 0; c < seg_chunks; ++c)
          tma_load_2d(smem_a + (s * 2 + c) * TC_A_CHUNK, &tm_a, s * TC_KSEG + c * 64, qtile * TC_M, a_full);
      int it = 0;
      for (int t = 0; t < n_my_tiles; ++t) {
        const int row = (tile_begin + t) * TC_N;
        for (int c = 0; c < b_chunks; ++c, ++it) {
          const int s = it % TC_STAGES;
          mbar_wait_sleep(&empty[s], ((it / TC_STAGES) & 1) ^ 1, 64);
          mbar_expect_tx(&full[s], TC_B_CHUNK);
          const int seg = c / seg_chunks, cc = c % seg_chunks;
          tma_load_2d(smem_b + s * TC_B_CHUNK, &tm_b, seg * TC_KSEG + cc * 64, row, &full[s]);
        }
      }
    }
  } else if (warp == 1) {
    // ================= MMA issuer =================
    if (lane == 0 && n_my_tiles > 0) {
      const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(TC_N >> 3) << 17) | ((uint32_t)(TC_M >> 4) << 24);
      mbar_wait_sleep(a_full, 0, 64);
      tc_fence_after();
      int it = 0;
      for (int t = 0; t < n_my_tiles; ++t) {
        const int buf = t & 1;
        mbar_wait_sleep(&tmem_empty[buf], ((t >> 1) & 1) ^ 1, 32);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)buf * 256;
        uint32_t accumulate = 0;
        for (int c = 0; c < b_chunks; ++c, ++it) {
          const int s = it % TC_STAGES;
          mbar_wait_sleep(&full[s], (it / TC_STAGES) & 1, 32);
          tc_fence_after();
          const int seg = c / seg_chunks, cc = c % seg_chunks;
          const int ks = (cc == 0) ? min(p.ksteps, 4) : (p.ksteps - 4);
          const uint64_t bdesc = make_desc(smem_b + s * TC_B_CHUNK);
          // B hi chunk pairs with A hi and A lo; B lo chunk pairs with A hi only
          const int n_a = (seg == 0 && p.n_terms == 3) ? 2 : 1;
          for (int a = 0; a < n_a; ++a) {
            const uint64_t adesc = make_desc(smem_a + (a * 2 + cc) * TC_A_CHUNK);
            for (int k = 0; k < ks; ++k) {
              tc_mma(d_tmem, adesc + (uint64_t)(k * 2), bdesc + (uint64_t)(k * 2), idesc, accumulate);
              accumulate = 1;
            }
          }
          tc_commit(&empty[s]);          // smem stage reusable once these MMAs have read it
        }
        tc_commit(&tmem_full[buf]);      // accumulator of this tile complete
      }
    }
  } else {
    // ================= epilogue =================
    const int ew = warp - 2;
    const int quarter = warp & 3;                 // TMEM lane quarter this warp may touch
    const int half = ew >> 2;                     // which 120 columns of the tile
    const int64_t q_global = (int64_t)qtile * TC_M + quarter * 32 + lane;
    const bool q_valid = q_global < p.n_queries;
    const float nq = p.nq[q_global];
    const float gq = p.gq[q_global];
    const int part = split * 2 + half;
    unsigned long long* my_list = nullptr;
    int my_cnt = 0;
    float my_tau = CUDART_INF_F;
    float thr = CUDART_INF_F;     // tau * (1 + 2e-6): slightly inclusive filter threshold
    if (MODE == TC_TOPK) my_list = p.cand + ((int64_t)q_global * p.n_parts + part) * CAP;
    const uint32_t lane_addr = tmem_base + ((uint32_t)(quarter * 32) << 16);
    constexpr int VSUB = TC_SUB / TC_S;           // 4 videos per sub-block

    for (int t = 0; t < n_my_tiles; ++t) {
      const int buf = t & 1;
      const int tile = tile_begin + t;
      if (MODE == TC_TOPK) {
        const float tg = tau_fetch(p.tau_g + q_global);
        if (tg < my_tau) {
          my_tau = tg;
          thr = my_tau * 1.000002f;
        }
      }
      mbar_wait_sleep(&tmem_full[buf], (t >> 1) & 1, 20);
      tc_fence_after();
      for (int sb = 0; sb < TC_SUBS; ++sb) {
        float acc[TC_SUB];
        const uint32_t col = (uint32_t)(buf * 256 + half * (TC_N / 2) + sb * TC_SUB);
        tmem_ld16(lane_addr + col, acc, 0);
        tmem_ld8(lane_addr + col + 16, acc, 16);
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        const int64_t v_first = (int64_t)tile * TC_VID + half * (TC_VID / 2) + sb * VSUB;
        // ---- hot path: 24 distances, 4 x 21 threshold tests; everything else is behind `flags` ----
        float d[VSUB][TC_S];
        unsigned flags = 0;                       // bit j: video j has a candidate; bit 4+j: needs the exact fallback
#pragma unroll
        for (int j = 0; j < VSUB; ++j) {
          bool ex = false;
#pragma unroll
          for (int i = 0; i < TC_S; ++i) {
            const float d2 = acc[j * TC_S + i] + nq;
            ex |= d2 < gq;
            // split mode: every d2 < gq (negatives included) is redone exactly below, so no clamp
            d[j][i] = sqrt_approx(SPLIT ? d2 : fmaxf(d2, 0.f));
          }
          flags |= ex ? (16u << j) : 0u;
        }
        // ragged banks / the padded tail of the last tile: mask the clip slots a video does not have
        const bool ragged = (!p.uniform) || (v_first + VSUB > p.n_videos);
        int nn[VSUB];
#pragma unroll
        for (int j = 0; j < VSUB; ++j) nn[j] = TC_S;
        if (ragged) {
#pragma unroll
          for (int j = 0; j < VSUB; ++j) {
            const int64_t v = v_first + j;
            nn[j] = (v < p.n_videos) ? (int)p.nseg[v] : 0;
#pragma unroll
            for (int i = 0; i < TC_S; ++i)
              if (i >= nn[j]) d[j][i] = CUDART_INF_F;
          }
        }
        if (flags >> 4) {
          // near-duplicate embeddings: the GEMM expansion cancels; redo those clips with the exact form
          if (q_valid) {
            const float* qr = p.queries + q_global * p.dim;
#pragma unroll
            for (int j = 0; j < VSUB; ++j) {
              if ((flags >> (4 + j)) & 1u) {
                const int64_t v = v_first + j;
                if (v < p.n_videos) {
                  const int c0 = p.vid_off[v];
#pragma unroll
                  for (int i = 0; i < TC_S; ++i)
                    if (i < nn[j] && (acc[j * TC_S + i] + nq) < gq)
                      d[j][i] = exact_distance(p.bank + (int64_t)(c0 + i) * p.dim, qr, p.dim);
                }
              }
            }
          }
        }
        if (MODE == TC_FULL) {
          if (q_valid) {
#pragma unroll
            for (int j = 0; j < VSUB; ++j)
              if (nn[j] > 0)
                write_video_scores(p.out_full + q_global * p.m_total + p.mom_off[v_first + j], nn[j], d[j][0], d[j][1],
                                   d[j][2], d[j][3], d[j][4], d[j][5]);
          }
        } else {
          // A moment (s, e) passes  mean(d[s..e]) <= tau  iff  sum_{i=s..e} (d_i - tau) <= 0, so a video has a
          // candidate iff the MINIMUM-SUM SUBARRAY of x_i = d_i - tau is <= 0: Kadane's recurrence, 4 ops per
          // clip instead of 15 adds + 21 compares per video.  (tau is padded by 2e-6; the exact test is
          // redone on the cold path.)
          float best[VSUB];
#pragma unroll
          for (int j = 0; j < VSUB; ++j) {
            float cur = d[j][0] - thr;
            best[j] = cur;
#pragma unroll
            for (int i = 1; i < TC_S; ++i) {
              cur = (d[j][i] - thr) + fminf(cur, 0.f);
              best[j] = fminf(best[j], cur);
            }
          }
          if (fminf(fminf(best[0], best[1]), fminf(best[2], best[3])) <= 0.f) {
#pragma unroll
            for (int j = 0; j < VSUB; ++j) flags |= (best[j] <= 0.f) ? (1u << j) : 0u;
          }
          if (flags & 15u) {
#pragma unroll
            for (int j = 0; j < VSUB; ++j)
              if ((flags >> j) & 1u)
                my_cnt = append_video(my_list, my_cnt, my_tau, nn[j], (unsigned)p.mom_off[v_first + j], d[j][0], d[j][1],
                                      d[j][2], d[j][3], d[j][4], d[j][5]);
          }
          if (__any_sync(0xffffffffu, my_cnt > TC_CAP_HI)) {
            const float tau_before = my_tau;
            compact_lists_noinline(my_list, my_cnt, my_tau, p.k, my_cnt > TC_CAP_HI, lane);
            if (my_tau != tau_before) {
              tau_publish(p.tau_g + q_global, my_tau);
              thr = my_tau * 1.000002f;
            }
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tmem_empty[buf]);
    }
    if (MODE == TC_TOPK) p.cand_cnt[q_global * p.n_parts + part] = my_cnt;
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512));
  }
}

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(ptr);
  }
  return fn;
}

static int make_map(CUtensorMap* map, const void* base, uint64_t rows, uint32_t box_rows) {
  EncodeTiledFn enc = get_encode();
  VFR_REQUIRE(enc, VFR_ERR_CUDA, "cuTensorMapEncodeTiled entry point not available");
  cuuint64_t gdim[2] = {(cuuint64_t)TC_ROW, (cuuint64_t)rows};
  cuuint64_t gstride[1] = {(cuuint64_t)TC_ROW * 2};
  cuuint32_t box[2] = {64, box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), gdim, gstride, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  VFR_REQUIRE(r == CUDA_SUCCESS, VFR_ERR_CUDA, "cuTensorMapEncodeTiled failed with CUresult %d", (int)r);
  return VFR_OK;
}

static int64_t tc_tiles(int64_t n_videos) { return (n_videos + TC_VID - 1) / TC_VID; }
static int64_t tc_qtiles(int64_t n_queries) { return (n_queries + TC_M - 1) / TC_M; }

// Bank splits per query tile (one CTA per SM).  Every split adds two candidate lists per query, and the
// filter threshold of a list only tightens with the moments that list has seen, so prefer the FEWEST
// splits that still keep >= 85 % of the SMs busy; with >= #SM query tiles there is no split at all.
static int tc_split(int64_t n_queries, int64_t n_tiles, int n_split) {
  if (n_split > 0) return (int)(n_split < n_tiles ? n_split : n_tiles);
  const int64_t qt = tc_qtiles(n_queries);
  int sms = 148, dev = 0;
  if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  int best = 1;
  double best_util = 0.0;
  for (int ns = 1; ns <= 64 && ns <= n_tiles; ++ns) {
    const int64_t ctas = ns * qt;
    const int64_t waves = (ctas + sms - 1) / sms;
    const double util = (double)ctas / (double)(waves * sms);
    if (util >= 0.85) return ns;
    if (util > best_util) { best_util = util; best = ns; }
  }
  return best;
}

}  // namespace vfr

using namespace vfr;

extern "C" size_t vfr_tc_bank_bytes(int64_t n_videos) {
  if (n_videos <= 0) return 0;
  const size_t slots = (size_t)tc_tiles(n_videos) * TC_N;
  const size_t vpad = (size_t)tc_tiles(n_videos) * TC_VID;
  return slots * TC_ROW * 2 + ((vpad + 15) / 16) * 16;
}

extern "C" int vfr_tc_bank_pack(const float* bank, const int32_t* vid_off, int64_t n_videos, int n_max, int dim,
                                int n_terms, void* packed, vfr_stream_t stream) {
  VFR_REQUIRE(bank && vid_off && packed, VFR_ERR_INVALID, "vfr_tc_bank_pack: null pointer");
  VFR_REQUIRE(n_videos > 0 && n_max >= 1 && n_max <= TC_S, VFR_ERR_UNSUPPORTED,
              "vfr_tc_bank_pack: the tensor-core path holds videos of at most %d clips (got n_max=%d)", TC_S, n_max);
  VFR_REQUIRE(dim >= 1 && dim + 3 <= TC_KSEG, VFR_ERR_UNSUPPORTED, "vfr_tc_bank_pack: dim=%d must be <= %d", dim, TC_KSEG - 3);
  VFR_REQUIRE(n_terms == 1 || n_terms == 3, VFR_ERR_INVALID, "vfr_tc_bank_pack: n_terms must be 1 or 3");
  const int64_t slots = tc_tiles(n_videos) * TC_N;
  __nv_bfloat16* out = reinterpret_cast<__nv_bfloat16*>(packed);
  uint8_t* nseg = reinterpret_cast<uint8_t*>(out + slots * TC_ROW);
  const int64_t blocks = (slots + 7) / 8;
  VFR_REQUIRE(blocks < (int64_t(1) << 31), VFR_ERR_UNSUPPORTED, "bank too large");
  tc_pack_bank_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(bank, vid_off, n_videos, slots, dim, n_terms, out, nseg);
  return check_launch("tc_pack_bank_kernel");
}

extern "C" size_t vfr_tc_query_bytes(int64_t n_queries) {
  if (n_queries <= 0) return 0;
  const size_t rows = (size_t)tc_qtiles(n_queries) * TC_M;
  return rows * TC_ROW * 2 + rows * 2 * sizeof(float);
}

extern "C" int vfr_tc_query_pack(const float* queries, int64_t n_queries, int dim, int n_terms, void* packed,
                                 vfr_stream_t stream) {
  VFR_REQUIRE(queries && packed, VFR_ERR_INVALID, "vfr_tc_query_pack: null pointer");
  VFR_REQUIRE(n_queries > 0 && dim >= 1 && dim + 3 <= TC_KSEG, VFR_ERR_UNSUPPORTED, "vfr_tc_query_pack: bad shape");
  VFR_REQUIRE(n_terms == 1 || n_terms == 3, VFR_ERR_INVALID, "vfr_tc_query_pack: n_terms must be 1 or 3");
  const int64_t rows = tc_qtiles(n_queries) * TC_M;
  __nv_bfloat16* out = reinterpret_cast<__nv_bfloat16*>(packed);
  float* nq = reinterpret_cast<float*>(out + rows * TC_ROW);
  float* gq = nq + rows;
  tc_pack_query_kernel<<<(unsigned)((rows + 7) / 8), 256, 0, (cudaStream_t)stream>>>(queries, n_queries, rows, dim,
                                                                                     n_terms, out, nq, gq);
  return check_launch("tc_pack_query_kernel");
}

static int tc_fill(TcParams& p, CUtensorMap& ma, CUtensorMap& mb, const void* bank_packed, const float* bank,
                   const int32_t* vid_off, const int64_t* mom_off, int64_t n_videos, int uniform, int dim, int n_terms,
                   const void* query_packed, const float* queries, int64_t n_queries) {
  VFR_REQUIRE(bank_packed && bank && vid_off && mom_off && query_packed && queries, VFR_ERR_INVALID, "score_tc: null pointer");
  VFR_REQUIRE(n_videos > 0 && n_queries > 0, VFR_ERR_INVALID, "score_tc: empty bank or batch");
  VFR_REQUIRE(dim >= 1 && dim + 3 <= TC_KSEG, VFR_ERR_UNSUPPORTED, "score_tc: dim=%d", dim);
  VFR_REQUIRE(n_terms == 1 || n_terms == 3, VFR_ERR_INVALID, "score_tc: n_terms");
  const int64_t tiles = tc_tiles(n_videos), qtiles = tc_qtiles(n_queries);
  VFR_REQUIRE(tiles < (int64_t(1) << 31) / TC_N && qtiles < (1 << 24), VFR_ERR_UNSUPPORTED, "score_tc: too large");
  const int64_t slots = tiles * TC_N, qrows = qtiles * TC_M;
  int rc = make_map(&ma, query_packed, (uint64_t)qrows, TC_M);
  if (rc) return rc;
  rc = make_map(&mb, bank_packed, (uint64_t)slots, TC_N);
  if (rc) return rc;
  p = TcParams{};
  const __nv_bfloat16* qp = reinterpret_cast<const __nv_bfloat16*>(query_packed);
  p.nq = reinterpret_cast<const float*>(qp + qrows * TC_ROW);
  p.gq = p.nq + qrows;
  p.nseg = reinterpret_cast<const uint8_t*>(reinterpret_cast<const __nv_bfloat16*>(bank_packed) + slots * TC_ROW);
  p.uniform = uniform;
  p.n_videos = n_videos;
  p.n_queries = n_queries;
  p.n_qtiles = (int)qtiles;
  p.n_tiles = (int)tiles;
  p.ksteps = (dim + 3 + 15) / 16;
  p.n_terms = n_terms;
  p.bank = bank;
  p.queries = queries;
  p.vid_off = vid_off;
  p.mom_off = mom_off;
  p.dim = dim;
  return VFR_OK;
}

extern "C" size_t vfr_score_topk_tc_bytes(int64_t n_queries, int64_t n_videos, int n_split) {
  if (n_queries <= 0 || n_videos <= 0) return 0;
  const int ns = tc_split(n_queries, tc_tiles(n_videos), n_split);
  const size_t qpad = (size_t)tc_qtiles(n_queries) * TC_M;
  const size_t parts = (size_t)ns * 2;
  return qpad * parts * CAP * sizeof(unsigned long long) + qpad * parts * sizeof(int32_t) + qpad * sizeof(unsigned);
}

extern "C" int vfr_score_topk_tc(const void* bank_packed, const float* bank, const int32_t* vid_off,
                                 const int64_t* mom_off, int64_t n_videos, int uniform, int dim, int n_terms,
                                 const void* query_packed, const float* queries, int64_t n_queries, int k,
                                 int64_t id_base, float* out_scores, int64_t* out_ids, void* workspace, int n_split,
                                 vfr_stream_t stream) {
  TcParams p;
  CUtensorMap ma, mb;
  int rc = tc_fill(p, ma, mb, bank_packed, bank, vid_off, mom_off, n_videos, uniform, dim, n_terms, query_packed,
                   queries, n_queries);
  if (rc) return rc;
  VFR_REQUIRE(out_scores && out_ids && workspace, VFR_ERR_INVALID, "vfr_score_topk_tc: null pointer");
  VFR_REQUIRE(k >= 1 && k <= VFR_TOPK_MAX, VFR_ERR_UNSUPPORTED, "k=%d not in [1,%d]", k, VFR_TOPK_MAX);
  const int ns_req = tc_split(n_queries, p.n_tiles, n_split);
  p.tiles_per_split = (p.n_tiles + ns_req - 1) / ns_req;
  const int ns = (p.n_tiles + p.tiles_per_split - 1) / p.tiles_per_split;
  const size_t qpad = (size_t)p.n_qtiles * TC_M;
  p.k = k;
  p.n_parts = ns * 2;
  p.cand = reinterpret_cast<unsigned long long*>(workspace);
  p.cand_cnt = reinterpret_cast<int32_t*>(p.cand + qpad * (size_t)ns_req * 2 * CAP);
  p.tau_g = reinterpret_cast<unsigned*>(p.cand_cnt + qpad * (size_t)ns_req * 2);
  cudaStream_t st = (cudaStream_t)stream;
  rc = launch_fill_u32(p.tau_g, 0x7f800000u, qpad, st);
  if (rc) return rc;
  VFR_CUDA(cudaFuncSetAttribute(score_tc_kernel<TC_TOPK, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TC_SMEM));
  VFR_CUDA(cudaFuncSetAttribute(score_tc_kernel<TC_TOPK, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TC_SMEM));
  if (n_terms == 3) score_tc_kernel<TC_TOPK, true><<<(unsigned)(p.n_qtiles * ns), TC_THREADS, TC_SMEM, st>>>(ma, mb, p);
  else score_tc_kernel<TC_TOPK, false><<<(unsigned)(p.n_qtiles * ns), TC_THREADS, TC_SMEM, st>>>(ma, mb, p);
  rc = check_launch("score_tc_kernel<TOPK>");
  if (rc) return rc;
  return launch_topk_finish(p.cand, p.cand_cnt, p.n_parts, k, id_base, n_queries, out_scores, out_ids, st);
}

extern "C" int vfr_score_full_tc(const void* bank_packed, const float* bank, const int32_t* vid_off,
                                 const int64_t* mom_off, int64_t n_videos, int uniform, int dim, int n_terms,
                                 const void* query_packed, const float* queries, int64_t n_queries, float* out,
                                 int64_t m_total, vfr_stream_t stream) {
  TcParams p;
  CUtensorMap ma, mb;
  int rc = tc_fill(p, ma, mb, bank_packed, bank, vid_off, mom_off, n_videos, uniform, dim, n_terms, query_packed,
                   queries, n_queries);
  if (rc) return rc;
  VFR_REQUIRE(out && m_total > 0, VFR_ERR_INVALID, "vfr_score_full_tc: bad output");
  const int ns_req = tc_split(n_queries, p.n_tiles, 0);
  p.tiles_per_split = (p.n_tiles + ns_req - 1) / ns_req;
  const int ns = (p.n_tiles + p.tiles_per_split - 1) / p.tiles_per_split;
  p.out_full = out;
  p.m_total = m_total;
  cudaStream_t st = (cudaStream_t)stream;
  VFR_CUDA(cudaFuncSetAttribute(score_tc_kernel<TC_FULL, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TC_SMEM));
  VFR_CUDA(cudaFuncSetAttribute(score_tc_kernel<TC_FULL, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TC_SMEM));
  if (n_terms == 3) score_tc_kernel<TC_FULL, true><<<(unsigned)(p.n_qtiles * ns), TC_THREADS, TC_SMEM, st>>>(ma, mb, p);
  else score_tc_kernel<TC_FULL, false><<<(unsigned)(p.n_qtiles * ns), TC_THREADS, TC_SMEM, st>>>(ma, mb, p);
  return check_launch("score_tc_kernel<FULL>");
}
