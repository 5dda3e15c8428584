// K4 (filter + refine): the k best moments of every query over the whole bank, bit-identical to the
// exact-fp32 engine (vfr_score_topk), at ONE fp16 tensor-core pass per (query, clip).
// Replaces reference model/evaluate.py:49-58 (distance + moment means) and :71-80 (argsort, top-k).
//
// Idea.  A moment's score is the MEAN of its clips' distances (evaluate.py:58), so it is >= the distance
// of its closest clip.  Single clips are moments themselves, hence the k-th best moment score tau_k is
// <= c_k, the k-th smallest clip distance of the bank.  Every moment of the final top-k therefore
// belongs to a video that owns a clip with distance <= c_k: the search over 21 M moments reduces to
//   stage 1 (filter)  find, per query, all clips whose squared distance is <= c_k^2        [tensor cores]
//   stage 2 (refine)  score all moments of the <= ~k videos owning those clips exactly      [CUDA cores]
// Stage 1 needs neither the sqrt nor the moment means: it is a GEMM whose epilogue is one 3-input min
// per two accumulator elements and one compare per 64.
//
// Stage 1 arithmetic.  fp16 operands (11-bit significands), exact products, fp32 accumulation in TMEM:
//   acc[r][c] = sum_k A[r][k] B[c][k]        A[r] = [ q_r 2^sq_r | 2^(sq_r+t) x3 ]   B[c] = [ -2 v_c 2^sb | nv_c 2^(sb-t) as an fp16 triple ]
//   (q, v here are CENTRED on the bank mean - d^2 is translation invariant - so a common offset of the embeddings does not inflate |q||v|)
//             = 2^(sq_r+sb) ( |v_c + eps|^2 - 2 q_r.v_c )  =  2^(sq_r+sb) ( d^2 - nq_r )       (+ rounding)
// with power-of-two scales (per bank: sb, t; per query: sq_r) that keep the operands in fp16's normal
// range.  The rounding error of the approximate d^2 is bounded RIGOROUSLY per query by
//   E_r = 1.05 * 2^-10 * |q_r| * 2 max|v|  + (sub-normal, triple, accumulation and key-rounding terms)
// and the filter keeps every clip with  d2~ <= tau2 + 2 E_r  where tau2 is the k-th smallest d2~ seen so
// far: if S is the set of the k smallest d2~, their exact d^2 are <= tau2 + E, so c_k^2 <= tau2 + E, and a
// clip with exact d^2 <= c_k^2 has d2~ <= tau2 + 2E.  No exact candidate can be lost; stage 2 recomputes
// everything that survives with the arithmetic of vfr_score.cu (sequential fp32 FADD/FFMA over k,
// IEEE sqrt and division), so scores, ids and tie order equal the exact engine's bit for bit.
//
// Stage 1 pipeline (one CTA per SM, 352 threads, warp-specialised; the default scheme, NB = 8 below):
//   warp 0       TMA producer: the CTA's query tiles (R x 128 rows, resident) once, then bank tiles of 256
//                clip rows as two [256 x 64] fp16 boxes (SWIZZLE_128B) through an mbarrier ring
//   warps 1, 2   MMA issuers, one per query tile and TMEM accumulator: tcgen05.mma cta_group::1 kind::f16, M=128,
//                N=256, K=16; 7 per (tile, query tile) at D=100, committed right after they are issued; a ring
//                stage is free when BOTH have released it (R = 1 and the other schemes: warp 1 alone, 320 threads)
//   warps 3..10  epilogue, two sets of four warps (one warp per TMEM lane quarter); set s owns TMEM buffer
//                s, so a thread sees ALL 256 columns of ONE query row and keeps ONE candidate list:
//                tcgen05.ld 32x32b.x32 double-buffered in registers (the next 64 columns load while the
//                current 64 go through an FMNMX3 tree and one compare against the query's threshold);
//                the rare hit appends (d2~, clip id) to the list (band-preserving radix-select
//                compaction, cf. vfr_topk.cuh); the TMEM buffer is released after the last load
// R = 2 query tiles share every bank tile in shared memory, halving the L2 -> SM traffic per MMA
// (set s = query tile s); with R = 1 the two sets take the bank tiles of even / odd parity.
#include "vfr_common.cuh"
#include "vfr_topk.cuh"
#include <algorithm>
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <math_constants.h>
#include <stdlib.h>

namespace vfr {

constexpr int SL_N = 256;                 // clip rows per bank tile (MMA N)
constexpr int SL_M = 128;                 // query rows per MMA
constexpr int SL_ROW = 128;               // fp16 columns per packed row when D + 3 <= 128 ...
constexpr int SL_MAXROW = 1088;           // ... larger D: rows of ceil((D + 3) / 64) * 64 columns, D <= 1085 (K-streaming kernel)
constexpr int SL_A_CHUNK = SL_M * 128;    // bytes of one [128 x 64 fp16] box
constexpr int SL_B_CHUNK = SL_N * 128;    // bytes of one [256 x 64 fp16] box = 32 KB
constexpr int SL_SET_WARPS = 4;           // one epilogue warp per TMEM lane quarter ...
constexpr int SL_THREADS = 64 + 2 * SL_SET_WARPS * 32;   // ... in two sets (one per TMEM buffer): 320 threads
constexpr int SL_THREADS2 = SL_THREADS + 32;             // the two-issuer variant (NB = 5) has one more non-epilogue warp
constexpr int SL_CAP = 1024;              // candidate slots per (query, part) list
constexpr int SL_CAP_HI = SL_CAP - SL_N;  // a tile can append at most SL_N keys to a list
constexpr int SL_PF_SHARE = 8;            // every SL_PF_SHARE-th CTA prefetches a given bank tile into L2
constexpr int SL_QPAD = 512;              // packed query rows are padded to a multiple of this (R x 128 rows x CTA pair)

struct SlBankMeta {          // written by the bank pack kernels, read by the query pack kernel
  unsigned vmax_abs_bits;    // max |v_k|                    (fp32 bit patterns: non-negative, so uint order)
  unsigned vnorm_max_bits;   // max ||v_c||
  unsigned nv_max_bits;      // max nv_c = |v_c + eps|^2
  unsigned vsum_abs_max_bits;// max_c sum_k |v_ck|
  int sb;                    // operand scale exponent of the bank
  int t;                     // exponent of the nv columns
  int pad[2];
  double colsum[SL_MAXROW];  // column sums of the bank (pack-time scratch)
  float center[SL_MAXROW];   // bank mean: both operands are centred on it (d^2 is translation invariant), which
                             // keeps |q'||v'| - and with it the error bound - small for embeddings with a common offset
};

struct SlParams {
  const float4* qmeta;       // [Qpad] {nq, scale = 2^(sq+sb), 1/scale, band2 = 2E}
  int64_t n_clips;
  int64_t n_queries;
  int n_qgroups;             // CTAs along the query dimension (R query tiles each)
  int n_tiles;
  int tiles_per_split;
  int ksteps;                // 16-wide k steps (ceil((D+3)/16))
  int n_chunks;              // 64-column chunks per packed row (K-streaming kernel, D + 3 > 128)
  int k;
  unsigned long long* cand;  // [Qpad][n_parts][SL_CAP]  (d2~ bits << 32 | clip id)
  int32_t* cand_cnt;         // [Qpad][n_parts]
  int n_parts;
  unsigned* tau_g;           // [Qpad] the threshold the filter uses: min(tau_cert, the sampled starting threshold)
  unsigned* tau_cert;        // [Qpad] CERTIFIED part of it: k-th smallest d2~ of keys really seen (or a bound put from outside)
  int tile_stride;           // bank tile of scan position i = i * tile_stride (1; the sample pass strides over the bank)
  int cap_trig;              // list length that triggers the FIRST compaction of a list (<= SL_CAP_HI)
  int cap_step;              // later ones: this many keys above what the previous compaction kept
  float* samp;               // sample pass, optional export: [Qpad][n_parts][SL_J] ascending upper bounds of EXACT d^2
  int sample_j;              // sample pass: the starting threshold is the sample_j-th smallest sampled minimum
  float* tau_part;           // [Qpad][n_parts] ceil(k / n_parts)-th smallest d2~ of each list (inf until compacted)
  int tile_lo;               // first bank tile of this launch (the scan may be split into several launches)
  int resume;                // != 0: the candidate lists continue from a previous launch
  int wait_mode;             // mbarrier wait flavour (see sl_wait)
  long long* dbg;            // optional timeline of CTA 0 (development aid): [9 roles][256 events][4] + 8 counters
  int32_t* flags;            // [Qpad] != 0: a candidate list overflowed / scales out of range (see vfr.h)
  unsigned long long* stats; // [2] {warp-level compaction events, lists compacted} since the lists were started (vfr_sel_stats)
  int prod_dbg;              // development aid: the producer's request times replace epilogue warp 7's slots of the timeline
  int rot_step;              // tiles by which consecutive CTA groups are rotated against each other (0: off; VFR_SEL_ROT)
  int prefetch;              // bank tiles the producer asks L2 for ahead of its loads (0: off; VFR_SEL_PF)
  const uint32_t* qpack;     // the packed query rows themselves ([Qpad][128] fp16 as 64 words): the NB = 3 kernel keeps them in TMEM
};

// ---------------------------------------------------------------------------------------------
// packing
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void atomic_max_pos(unsigned* dst, float v) { atomicMax(dst, __float_as_uint(fmaxf(v, 0.f))); }

// column sums of the bank -> its mean (the centre)
__global__ void sl_bank_colsum_kernel(const float* __restrict__ bank, int64_t n_clips, int dim, SlBankMeta* meta) {
  for (int k = threadIdx.x; k < dim; k += blockDim.x) {   // a block walks a strided set of rows, coalesced over k
    double s = 0.0;
    for (int64_t c = blockIdx.x; c < n_clips; c += gridDim.x) s += (double)bank[c * dim + k];
    atomicAdd(&meta->colsum[k], s);
  }
}
__global__ void sl_bank_center_kernel(int64_t n_clips, int dim, SlBankMeta* meta) {
  for (int k = threadIdx.x; k < SL_MAXROW; k += blockDim.x)
    meta->center[k] = (k < dim) ? (float)(meta->colsum[k] / (double)n_clips) : 0.f;
}

// one warp per clip row: maxima the scales and the error bound are derived from (centred values)
__global__ void sl_bank_stats_kernel(const float* __restrict__ bank, int64_t n_clips, int dim, SlBankMeta* meta) {
  const int lane = threadIdx.x & 31;
  const int64_t warp0 = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int64_t nwarps = (int64_t)gridDim.x * (blockDim.x >> 5);
  float m_abs = 0.f, m_norm = 0.f, m_nv = 0.f, m_sum = 0.f;
  for (int64_t c = warp0; c < n_clips; c += nwarps) {
    const float* src = bank + c * dim;
    double ss = 0.0, sm = 0.0;
    float sa = 0.f, ma = 0.f;
    for (int k = lane; k < dim; k += 32) {
      const float x = __fsub_rn(src[k], meta->center[k]);
      ss += (double)x * x;
      sm += (double)x;
      sa += fabsf(x);
      ma = fmaxf(ma, fabsf(x));
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      ss += __shfl_xor_sync(0xffffffffu, ss, o);
      sm += __shfl_xor_sync(0xffffffffu, sm, o);
      sa += __shfl_xor_sync(0xffffffffu, sa, o);
      ma = fmaxf(ma, __shfl_xor_sync(0xffffffffu, ma, o));
    }
    const double eps = (double)VFR_PAIRWISE_EPS;
    const double nv = ss + 2.0 * eps * sm + (double)dim * eps * eps;
    m_abs = fmaxf(m_abs, ma);
    m_norm = fmaxf(m_norm, (float)(sqrt(ss) * 1.0000002));
    m_nv = fmaxf(m_nv, (float)(nv * 1.0000002));
    m_sum = fmaxf(m_sum, sa * 1.00001f);
  }
  if (lane == 0) {
    atomic_max_pos(&meta->vmax_abs_bits, m_abs);
    atomic_max_pos(&meta->vnorm_max_bits, m_norm);
    atomic_max_pos(&meta->nv_max_bits, m_nv);
    atomic_max_pos(&meta->vsum_abs_max_bits, m_sum);
  }
}

__global__ void sl_bank_scales_kernel(SlBankMeta* meta) {
  const float vmax = __uint_as_float(meta->vmax_abs_bits);
  const float nvmax = __uint_as_float(meta->nv_max_bits);
  int sb = 0, t = 0;
  if (vmax > 0.f && vmax < CUDART_INF_F) sb = 6 - ilogbf(vmax);          // 2 |v| 2^sb < 2^8
  sb = max(-60, min(60, sb));
  if (nvmax > 0.f && nvmax < CUDART_INF_F) t = ilogbf(nvmax) + sb - 13;  // nv 2^(sb-t) < 2^14
  t = max(-120, min(120, t));
  meta->sb = sb;
  meta->t = t;
}

// x (>= 0, < 2^15) as the sum of three fp16 values
__device__ __forceinline__ void split3_h(double x, __half& a, __half& b, __half& c) {
  a = __double2half(x);
  const double r1 = x - (double)__half2float(a);
  b = __double2half(r1);
  c = __double2half(r1 - (double)__half2float(b));
}

// one warp per packed row (rows >= n_clips are padding: zero operands, nv = 2^15)
__global__ void sl_bank_pack_kernel(const float* __restrict__ bank, int64_t n_clips, int64_t n_rows_pad, int dim, int pitch,
                                    const SlBankMeta* __restrict__ meta, __half* __restrict__ out) {
  const int lane = threadIdx.x & 31;
  const int64_t c = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (c >= n_rows_pad) return;
  __half* row = out + c * pitch;
  const __half zero = __float2half_rn(0.f);
  if (c >= n_clips) {
    for (int k = lane; k < pitch; k += 32) row[k] = (k == dim) ? __float2half_rn(32768.f) : zero;
    return;
  }
  const int sb = meta->sb, t = meta->t;
  const float* src = bank + c * dim;
  double ss = 0.0, sm = 0.0;
  for (int k = lane; k < pitch; k += 32) {
    __half h = zero;
    if (k < dim) {
      const float x = __fsub_rn(src[k], meta->center[k]);
      ss += (double)x * x;
      sm += (double)x;
      h = __float2half_rn(scalbnf(-2.f * x, sb));
    }
    if (k < dim || k >= dim + 3) row[k] = h;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    ss += __shfl_xor_sync(0xffffffffu, ss, o);
    sm += __shfl_xor_sync(0xffffffffu, sm, o);
  }
  if (lane == 0) {
    const double eps = (double)VFR_PAIRWISE_EPS;
    const double nv = ss + 2.0 * eps * sm + (double)dim * eps * eps;
    __half a, b, c3;
    split3_h(scalbn(nv, sb - t), a, b, c3);
    row[dim] = a;
    row[dim + 1] = b;
    row[dim + 2] = c3;
  }
}

// one warp per query row
__global__ void sl_query_pack_kernel(const float* __restrict__ q, int64_t n_queries, int64_t n_rows_pad, int dim, int pitch,
                                     const SlBankMeta* __restrict__ meta, __half* __restrict__ out,
                                     float4* __restrict__ qmeta, int32_t* __restrict__ flags) {
  const int lane = threadIdx.x & 31;
  const int64_t r = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (r >= n_rows_pad) return;
  __half* row = out + r * pitch;
  const __half zero = __float2half_rn(0.f);
  if (r >= n_queries) {
    for (int k = lane; k < pitch; k += 32) row[k] = zero;
    if (lane == 0) { qmeta[r] = make_float4(0.f, 1.f, 1.f, 0.f); flags[r] = 0; }
    return;
  }
  const float* src = q + r * dim;
  double ss = 0.0, sm = 0.0;
  float sa = 0.f, ma = 0.f;
  for (int k = lane; k < dim; k += 32) {
    const float x = __fsub_rn(src[k], meta->center[k]);
    ss += (double)x * x;
    sm += (double)x;
    sa += fabsf(x);
    ma = fmaxf(ma, fabsf(x));
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    ss += __shfl_xor_sync(0xffffffffu, ss, o);
    sm += __shfl_xor_sync(0xffffffffu, sm, o);
    sa += __shfl_xor_sync(0xffffffffu, sa, o);
    ma = fmaxf(ma, __shfl_xor_sync(0xffffffffu, ma, o));
  }
  const int sb = meta->sb, t = meta->t;
  int sq = 0, bad = 0;
  if (ma > 0.f && ma < CUDART_INF_F) sq = 7 - ilogbf(ma);      // |q| 2^sq < 2^8
  if (!(ma < CUDART_INF_F) || !(ss < 1e300)) bad = 1;          // inf / nan in the query
  sq = min(sq, 14 - t);                                          // the nv multiplier 2^(sq+t) must fit fp16
  if (sq + t < -24 || sq + sb > 100 || sq + sb < -100) { bad = 1; sq = 0; }
  for (int k = lane; k < pitch; k += 32) {
    __half h = zero;
    if (k < dim) h = __float2half_rn(scalbnf(__fsub_rn(src[k], meta->center[k]), sq));
    else if (k < dim + 3) h = bad ? zero : __float2half_rn(scalbnf(1.f, sq + t));
    row[k] = h;
  }
  if (lane == 0) {
    const double eps = (double)VFR_PAIRWISE_EPS;
    const double nq = ss - 2.0 * eps * sm;
    const double qn = sqrt(ss);
    const double vn = (double)__uint_as_float(meta->vnorm_max_bits);
    const double nvm = (double)__uint_as_float(meta->nv_max_bits);
    const double vsa = (double)__uint_as_float(meta->vsum_abs_max_bits);
    // rigorous bound of |d2~ - d^2| (see the header): operand rounding, sub-normal operands, the nv
    // triple, fp32 accumulation inside the tensor core (2^-18 of the absolute sum: >= 32 x its observed
    // error) and the rounding of the key / threshold arithmetic
    const double main_term = 1.05 * (1.0 / 1024.0) * (1.0 + 1.0 / 4096.0) * qn * 2.0 * vn;
    const double sub_term = 1.001 * ldexp(1.0, -25) * (ldexp((double)sa, -sb) + ldexp(2.0 * vsa, -sq));
    const double nv_term = ldexp(nvm, -30) + ldexp(1.0, t - sb - 24);
    // (2^-18 of the absolute sum covers K <= 128 with a factor 32 to spare; a longer contraction accumulates in more steps)
    const double acc_term = ldexp(2.0 * qn * vn + nvm, -18) * fmax(1.0, (double)pitch / 128.0);
    const double key_term = ldexp(fabs(nq) + nvm + 2.0 * qn * vn, -21);
    const double ctr_term = ldexp((qn + vn) * (qn + vn), -22);   // fp32 rounding of the centring subtractions
    const double E = main_term + sub_term + nv_term + acc_term + key_term + ctr_term;
    const float scale = scalbnf(1.f, sq + sb);
    qmeta[r] = make_float4((float)nq, scale, scalbnf(1.f, -(sq + sb)), (float)(2.0 * E * 1.0001));
    flags[r] = bad;
  }
}

// ---------------------------------------------------------------------------------------------
// tcgen05 / TMA primitives
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void sl_tma_load(void* dst, const CUtensorMap* map, int c0, int c1, uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(
          smem_u32(dst)),
      "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}
// ask L2 for a box of the tensor (no shared-memory destination, no completion): the later TMA load finds it there
__device__ __forceinline__ void sl_tma_prefetch(const CUtensorMap* map, int c0, int c1) {
  asm volatile("cp.async.bulk.prefetch.tensor.2d.L2.global.tile [%0, {%1, %2}];" ::"l"(map), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void sl_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void sl_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void sl_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void sl_mma(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(d_tmem),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc)
      : "memory");
}
// One lane of a CONVERGED warp (deterministic: the same lane for the same mask every time, so the MMAs and the commits
// that track them come from one thread).  An issuer that runs as `if (lane == 0)` makes the compiler wrap every tcgen05.mma
// in an ELECT / BRA.U.ANY loop with the operands moved by R2UR (~120 cycles of issue per MMA, measured: that - not the
// tensor pipe - bounded jobs of N = 128); from a converged warp the UTCHMMAs issue back to back out of uniform registers.
__device__ __forceinline__ bool sl_elect() {
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
// tcgen05.mma from the elected lane of a converged warp, the election inside the asm block (no C++ branch: the operands stay
// in uniform registers and consecutive UTCHMMAs are three instructions apart)
__device__ __forceinline__ void sl_mma_e(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n"
      ".reg .pred p, e;\n"
      "elect.sync _|e, 0xffffffff;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "@e tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(d_tmem),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc)
      : "memory");
}
__device__ __forceinline__ void sl_mma_ts_e(uint32_t d_tmem, uint32_t a_tmem, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n"
      ".reg .pred p, e;\n"
      "elect.sync _|e, 0xffffffff;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "@e tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n"
      "}\n" ::"r"(d_tmem),
      "r"(a_tmem), "l"(bdesc), "r"(idesc), "r"(acc)
      : "memory");
}
// N (1 ... 4) consecutive k steps of one 64-column chunk as ONE asm block from the elected lane of a converged warp: the
// operands are moved to uniform registers once per block, the + 2 k (descriptors, 32-byte steps >> 4) / + 8 k (TMEM columns)
// happen inside, and the UTCHMMAs follow each other directly.  acc: does the FIRST one accumulate?
template <int N>
__device__ __forceinline__ void sl_mma_block(uint32_t d_tmem, uint64_t a, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  if constexpr (N == 1) {
    asm volatile(
      "{\n"
      ".reg .pred p, t, e;\n"
      "elect.sync _|e, 0xffffffff;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "setp.eq.b32 t, 0, 0;\n"
      "@e tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(d_tmem),
      "l"(a), "l"(bdesc), "r"(idesc), "r"(acc)
      : "memory");
  }
  else if constexpr (N == 2) {
    asm volatile(
      "{\n"
      ".reg .pred p, t, e;\n"
      ".reg .b64 a1, b1;\n"
      "elect.sync _|e, 0xffffffff;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "setp.eq.b32 t, 0, 0;\n"
      "add.u64 a1, %1, 2;\n"
      "add.u64 b1, %2, 2;\n"
      "@e tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n"
      "@e tcgen05.mma.cta_group::1.kind::f16 [%0], a1, b1, %3, t;\n"
      "}\n" ::"r"(d_tmem),
      "l"(a), "l"(bdesc), "r"(idesc), "r"(acc)
      : "memory");
  }
  else if constexpr (N == 3) {
    asm volatile(
      "{\n"
      ".reg .pred p, t, e;\n"
      ".reg .b64 a1, b1, a2, b2;\n"
      "elect.sync _|e, 0xffffffff;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "setp.eq.b32 t, 0, 0;\n"
      "add.u64 a1, %1, 2;\n"
      "add.u64 b1, %2, 2;\n"
      "add.u64 a2, %1, 4;\n"
      "add.u64 b2, %2, 4;\n"
      "@e tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n"
      "@e tcgen05.mma.cta_group::1.kind::f16 [%0], a1, b1, %3, t;\n"
      "@e tcgen05.mma.cta_group::1.kind::f16 [%0], a2, b2, %3, t;\n"
      "}\n" ::"r"(d_tmem),
      "l"(a), "l"(bdesc), "r"(idesc), "r"(acc)
      : "memory");
  }
  else if constexpr (N == 4) {
    asm volatile(
      "{\n"
      ".reg .pred p, t, e;\n"
      ".reg .b64 a1, b1, a2, b2, a3, b3;\n"
      "elect.sync _|e, 0xffffffff;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "setp.eq.b32 t, 0, 0;\n"
      "add.u64 a1, %1, 2;\n"
      "add.u64 b1, %2, 2;\n"
      "add.u64 a2, %1, 4;\n"
      "add.u64 b2, %2, 4;\n"
      "add.u64 a3, %1, 6;\n"
      "add.u64 b3, %2, 6;\n"
      "@e tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n"
      "@e tcgen05.mma.cta_group::1.kind::f16 [%0], a1, b1, %3, t;\n"
      "@e tcgen05.mma.cta_group::1.kind::f16 [%0], a2, b2, %3, t;\n"
      "@e tcgen05.mma.cta_group::1.kind::f16 [%0], a3, b3, %3, t;\n"
      "}\n" ::"r"(d_tmem),
      "l"(a), "l"(bdesc), "r"(idesc), "r"(acc)
      : "memory");
  }
}
template <int N>
__device__ __forceinline__ void sl_mma_ts_block(uint32_t d_tmem, uint32_t a, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  if constexpr (N == 1) {
    asm volatile(
      "{\n"
      ".reg .pred p, t, e;\n"
      "elect.sync _|e, 0xffffffff;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "setp.eq.b32 t, 0, 0;\n"
      "@e tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n"
      "}\n" ::"r"(d_tmem),
      "r"(a), "l"(bdesc), "r"(idesc), "r"(acc)
      : "memory");
  }
  else if constexpr (N == 2) {
    asm volatile(
      "{\n"
      ".reg .pred p, t, e;\n"
      ".reg .b32 a1;\n"
      ".reg .b64 b1;\n"
      "elect.sync _|e, 0xffffffff;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "setp.eq.b32 t, 0, 0;\n"
      "add.u32 a1, %1, 8;\n"
      "add.u64 b1, %2, 2;\n"
      "@e tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n"
      "@e tcgen05.mma.cta_group::1.kind::f16 [%0], [a1], b1, %3, t;\n"
      "}\n" ::"r"(d_tmem),
      "r"(a), "l"(bdesc), "r"(idesc), "r"(acc)
      : "memory");
  }
  else if constexpr (N == 3) {
    asm volatile(
      "{\n"
      ".reg .pred p, t, e;\n"
      ".reg .b32 a1, a2;\n"
      ".reg .b64 b1, b2;\n"
      "elect.sync _|e, 0xffffffff;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "setp.eq.b32 t, 0, 0;\n"
      "add.u32 a1, %1, 8;\n"
      "add.u64 b1, %2, 2;\n"
      "add.u32 a2, %1, 16;\n"
      "add.u64 b2, %2, 4;\n"
      "@e tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n"
      "@e tcgen05.mma.cta_group::1.kind::f16 [%0], [a1], b1, %3, t;\n"
      "@e tcgen05.mma.cta_group::1.kind::f16 [%0], [a2], b2, %3, t;\n"
      "}\n" ::"r"(d_tmem),
      "r"(a), "l"(bdesc), "r"(idesc), "r"(acc)
      : "memory");
  }
  else if constexpr (N == 4) {
    asm volatile(
      "{\n"
      ".reg .pred p, t, e;\n"
      ".reg .b32 a1, a2, a3;\n"
      ".reg .b64 b1, b2, b3;\n"
      "elect.sync _|e, 0xffffffff;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "setp.eq.b32 t, 0, 0;\n"
      "add.u32 a1, %1, 8;\n"
      "add.u64 b1, %2, 2;\n"
      "add.u32 a2, %1, 16;\n"
      "add.u64 b2, %2, 4;\n"
      "add.u32 a3, %1, 24;\n"
      "add.u64 b3, %2, 6;\n"
      "@e tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n"
      "@e tcgen05.mma.cta_group::1.kind::f16 [%0], [a1], b1, %3, t;\n"
      "@e tcgen05.mma.cta_group::1.kind::f16 [%0], [a2], b2, %3, t;\n"
      "@e tcgen05.mma.cta_group::1.kind::f16 [%0], [a3], b3, %3, t;\n"
      "}\n" ::"r"(d_tmem),
      "r"(a), "l"(bdesc), "r"(idesc), "r"(acc)
      : "memory");
  }
}
// A operand in TMEM (lane = row, a 32-bit column = two consecutive K elements), B through its shared-memory descriptor
__device__ __forceinline__ void sl_mma_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n"
      "}\n" ::"r"(d_tmem),
      "r"(a_tmem), "l"(bdesc), "r"(idesc), "r"(acc)
      : "memory");
}
// 32 consecutive 32-bit columns of this thread's TMEM lane
__device__ __forceinline__ void sl_st32(uint32_t taddr, const uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,"
      "%31,%32};" ::"r"(taddr),
      "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]), "r"(r[10]),
      "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]), "r"(r[18]), "r"(r[19]), "r"(r[20]),
      "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]), "r"(r[28]), "r"(r[29]), "r"(r[30]),
      "r"(r[31])
      : "memory");
}
// the same load delivered to the same shared-memory offsets (data and mbarrier) of every CTA in `mask`
__device__ __forceinline__ void sl_tma_load_mc(void* dst, const CUtensorMap* map, int c0, int c1, uint64_t* bar, uint16_t mask) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1, {%3, %4}], "
      "[%2], %5;" ::"r"(smem_u32(dst)),
      "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "h"(mask)
      : "memory");
}
__device__ __forceinline__ void sl_commit_mc(uint64_t* bar, uint16_t mask) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
                   smem_u32(bar)),
               "h"(mask)
               : "memory");
}
__device__ __forceinline__ void sl_cluster_sync() {
  asm volatile("barrier.cluster.arrive.release.aligned;\nbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ int sl_cluster_rank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return (int)r;
}
// K-major, SWIZZLE_128B operand tile: rows of 128 B, 8-row groups 1024 B apart
__device__ __forceinline__ uint64_t sl_desc(const void* smem_ptr) {
  const uint32_t addr = smem_u32(smem_ptr);
  uint64_t d = 0;
  d |= (uint64_t)((addr >> 4) & 0x3fff);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}
// the upper word of sl_desc (it does not depend on the address)
__host__ __device__ constexpr uint32_t sl_desc_hi() { return (uint32_t)(1024 >> 4) | (1u << 14) | (2u << 29); }
__device__ __forceinline__ void sl_ld32(uint32_t taddr, float (&v)[64], int off) {
  uint32_t r[32];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,"
      "%30,%31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr));
#pragma unroll
  for (int i = 0; i < 32; ++i) v[off + i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ float fmin3(float a, float b, float c) {
  float d;
  asm("min.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c));
  return d;
}
// mbarrier wait with a watchdog (a protocol bug traps instead of hanging the GPU).  mode 0: try_wait with a
// suspend-time hint (the thread sleeps in hardware); mode 1: plain try_wait loop; mode 2: short hint
__device__ __forceinline__ void sl_wait(uint64_t* bar, uint32_t parity, int mode) {
  uint32_t done = 0;
  unsigned long long t0 = 0;
  for (uint32_t spin = 0;; ++spin) {
    if (mode == 1) {
      asm volatile(
          "{\n"
          ".reg .pred p;\n"
          "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
          "selp.u32 %0, 1, 0, p;\n"
          "}\n"
          : "=r"(done)
          : "r"(smem_u32(bar)), "r"(parity)
          : "memory");
    } else {
      asm volatile(
          "{\n"
          ".reg .pred p;\n"
          "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n"
          "selp.u32 %0, 1, 0, p;\n"
          "}\n"
          : "=r"(done)
          : "r"(smem_u32(bar)), "r"(parity), "r"(mode == 0 ? 1000000u : 200u)
          : "memory");
    }
    if (done) return;
    if ((spin & 0x3fff) == 0x3fff) {
      unsigned long long now;
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
      if (t0 == 0) t0 = now;
      else if (now - t0 > 4000000000ull) {
        printf("vfr: select_tc mbarrier wait timed out (block %d thread %d)\n", blockIdx.x, threadIdx.x);
        __trap();
      }
    }
  }
}

// non-blocking probe of an mbarrier phase
__device__ __forceinline__ bool sl_test(uint64_t* bar, uint32_t parity) {
  uint32_t done;
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n"
      "selp.u32 %0, 1, 0, p;\n"
      "}\n"
      : "=r"(done)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return done != 0;
}

// a smem stage is free again once the MMAs issued so far have read it - in every CTA that received it
template <int CL>
__device__ __forceinline__ void sl_release(uint64_t* bar) {
  if (CL == 1) sl_commit(bar);
  else sl_commit_mc(bar, (uint16_t)3);
}

// threshold of the accumulator domain, rounded up:  (tau2 + band2 - nq) * scale
__device__ __forceinline__ float sl_threshold(float tau2, float band2, float nq, float scale) {
  return __fmul_ru(__fsub_ru(__fadd_ru(tau2, band2), nq), scale);
}

// Band-preserving compaction of thread-private candidate lists (cf. compact_lists in vfr_topk.cuh, here for
// SL_CAP keys).  For each lane whose `need` is set the warp finds the k-th smallest key of that lane's list
// (most-significant-bit-first radix select on the d2~ word), makes it the lane's tau and keeps every key
// <= min(tau, tau_shared) + band2.  tau_part receives the ceil(k / P)-th smallest, P = lists per query: every list
// holds >= k/P keys under its own tau_part, so the LARGEST tau_part of the P lists bounds the k-th smallest d2~ of
// the whole bank - far tighter than any single list's k-th, which only knows 1/P of the clips.
constexpr int SL_SLOTS = SL_CAP / 32;   // keys per lane

// 32 x 32 bit transpose in registers: on return bit s of w[b] = bit b of the old w[s]
__device__ __forceinline__ void sl_transpose32(unsigned (&w)[SL_SLOTS]) {
  static_assert(SL_SLOTS == 32, "one key per lane and bit of the slot mask");
  unsigned m = 0x0000ffffu;
#pragma unroll
  for (int j = 16; j != 0; j >>= 1, m ^= (m << j)) {
#pragma unroll
    for (int k = 0; k < 32; ++k) {
      if ((k & j) == 0) {
        const unsigned t = ((w[k] >> j) ^ w[k + j]) & m;
        w[k + j] ^= t;
        w[k] ^= t << j;
      }
    }
  }
}

// The ka-th and the kb-th smallest word (1-based, ka, kb <= number of valid slots) of a warp's 32 x 32 words in ONE
// most-significant-bit-first radix descent over the TRANSPOSED words (tw[b] = bit b of this lane's 32 slots): per bit
// two popcounts and one hardware warp reduction carrying both counts (each <= 1024 < 2^16).
__device__ __forceinline__ void sl_radix_kth2(const unsigned (&tw)[SL_SLOTS], unsigned valid, int ka, int kb, unsigned& va,
                                              unsigned& vb) {
  unsigned ca = valid, cb = valid;
  va = 0u;
  vb = 0u;
#pragma unroll
  for (int bit = 31; bit >= 0; --bit) {
    const unsigned ones = tw[bit];
    const unsigned za = ca & ~ones, zb = cb & ~ones;
    const unsigned tot = __reduce_add_sync(0xffffffffu, (unsigned)__popc(za) | ((unsigned)__popc(zb) << 16));
    const int na = (int)(tot & 0xffffu), nb = (int)(tot >> 16);
    if (ka <= na) ca = za; else { ka -= na; ca &= ones; va |= 1u << bit; }
    if (kb <= nb) cb = zb; else { kb -= nb; cb &= ones; vb |= 1u << bit; }
  }
}

__device__ __noinline__ void sl_compact(unsigned long long* list, int& cnt, float& tau_own, float& tau_part, float tau_shared,
                                        float band2, int k, int k_part, bool need, int lane, long long* dbg = nullptr) {
  unsigned mask = __ballot_sync(0xffffffffu, need);
  while (mask) {
    const int src = __ffs(mask) - 1;
    mask &= mask - 1;
    unsigned long long* lp =
        reinterpret_cast<unsigned long long*>(__shfl_sync(0xffffffffu, reinterpret_cast<unsigned long long>(list), src));
    const int n = __shfl_sync(0xffffffffu, cnt, src);
    const float ts = __shfl_sync(0xffffffffu, tau_shared, src);
    const float b2 = __shfl_sync(0xffffffffu, band2, src);
    if (n <= k) continue;
    __syncwarp();
    const long long c0 = dbg ? clock64() : 0;
    unsigned kth_hi, kpart_hi;
    {
      // the d2~ words (high half of every key; every list owns 32 x 32 slots, slots >= n are masked), transposed
      unsigned w[SL_SLOTS];
      unsigned valid = 0u;
      const unsigned* hp = reinterpret_cast<const unsigned*>(lp) + 1;
#pragma unroll
      for (int s = 0; s < SL_SLOTS; ++s) w[s] = hp[2 * ((s << 5) | lane)];
#pragma unroll
      for (int s = 0; s < SL_SLOTS; ++s) {
        if (((s << 5) | lane) < n) valid |= 1u << s;
        else w[s] = 0xffffffffu;
      }
      long long c1 = 0;
      if (dbg) c1 = clock64();
      sl_transpose32(w);
      sl_radix_kth2(w, valid, k, min(k_part, k), kth_hi, kpart_hi);
      if (dbg && lane == 0) {
        unsigned long long* d = reinterpret_cast<unsigned long long*>(dbg + 9 * 256 * 4);
        atomicAdd(d + 3, (unsigned long long)(c1 - c0));
        atomicAdd(d + 4, (unsigned long long)(clock64() - c1));
      }
    }
    const float kth = __uint_as_float(kth_hi);
    const float kpart = __uint_as_float(kpart_hi);
    const unsigned keep_bits = __float_as_uint(__fadd_ru(fminf(kth, ts), b2));
    __syncwarp();
    const long long c2 = dbg ? clock64() : 0;
    // second read of the (cache-hot) keys: volatile asm keeps the 32 loads back to back, ahead of the in-place writes
    unsigned long long key[SL_SLOTS];
#pragma unroll
    for (int s = 0; s < SL_SLOTS; ++s)
      asm volatile("ld.global.u64 %0, [%1];" : "=l"(key[s]) : "l"(__cvta_generic_to_global(lp + ((s << 5) | lane))) : "memory");
    int base = 0;
#pragma unroll
    for (int s = 0; s < SL_SLOTS; ++s) {
      const int idx = (s << 5) | lane;
      const bool keep = idx < n && (unsigned)(key[s] >> 32) <= keep_bits;
      const unsigned bal = __ballot_sync(0xffffffffu, keep);
      if (keep) lp[base + __popc(bal & ((1u << lane) - 1u))] = key[s];
      base += __popc(bal);
    }
    __syncwarp();
    if (dbg && lane == 0) {
      unsigned long long* d = reinterpret_cast<unsigned long long*>(dbg + 9 * 256 * 4);
      atomicAdd(d + 5, (unsigned long long)(clock64() - c2));
    }
    if (lane == src) {
      cnt = base;
      tau_own = kth;
      tau_part = kpart;
    }
  }
}

// ---------------------------------------------------------------------------------------------
// stage 1: the filter kernel
// ---------------------------------------------------------------------------------------------
template <int R>
struct SlCfg {
  static constexpr int STAGES = (R == 1) ? 6 : 5;
  static constexpr uint32_t SMEM = R * 2 * SL_A_CHUNK + STAGES * SL_B_CHUNK + 1024 /*align*/ + 256 /*barriers*/;
};
// K-streaming variant (D + 3 > 128: the query tiles no longer fit in shared memory next to the ring): a stage holds one
// 64-column chunk of BOTH query tiles and of the bank tile; the two accumulators integrate over all chunks
constexpr int SL_BIG_STAGES = 3;
constexpr int SL_BIG_STAGE = 2 * SL_A_CHUNK + SL_B_CHUNK;      // 64 KB
constexpr uint32_t SL_BIG_SMEM = SL_BIG_STAGES * SL_BIG_STAGE + 1024 + 256;
// NB = 3 (query tiles in TMEM): shared memory holds only the ring of bank tiles
constexpr int SL_ATM_STAGES = 6;
constexpr uint32_t SL_ATM_SMEM = SL_ATM_STAGES * SL_B_CHUNK + 1024 + 256;
constexpr uint32_t SL_ATM_ACOL = 384;     // TMEM columns [384, 512): the two query tiles (64 columns = 128 fp16 each)

// per-thread state of the epilogue: one query row, one candidate list
struct SlRow {
  unsigned long long* list;
  int cnt;
  int trig;          // list length that triggers the next compaction of this list
  float tau_use;     // min(k-th smallest d2~ of this list at its last compaction, the query's shared tau): what the filter uses
  float thr;         // the same in the accumulator domain, band included
  float nq, inv_scale, scale, band2;   // the query's qmeta
  int64_t q;
  int64_t clip0;     // first clip of the current tile
};

__device__ __forceinline__ void sl_append(SlRow& st, float x, int64_t clip, const SlParams& p) {
  if (clip < p.n_clips) {
    const float key = fmaxf(__fmaf_rn(x, st.inv_scale, st.nq), 0.f);
    st.list[st.cnt++] = ((unsigned long long)__float_as_uint(key) << 32) | (unsigned)clip;
  }
}

// Cold path of sl_process: some of the 64 columns pass the filter.  Branch-free pass mask (FSETP + predicated
// LOP3, four independent chains); the usual case is ONE passing column, whose value is the minimum the hot path
// already holds, so no dynamic register indexing is needed.  Several passing columns (start of the scan) go
// through a local-memory copy.
__device__ __forceinline__ void sl_cold(const float (&v)[64], float ma, float mb, float m, int coff, SlRow& st,
                                        const SlParams& p) {
  // (the mask of a 32-column half is only built if that half's minimum passes: the mask is most of a pass, and a
  // pass is paid for with the hold time of the TMEM buffer)
  unsigned lo = 0u, hi = 0u;
  if (ma <= st.thr) {
    unsigned m0 = 0u, m1 = 0u, m2 = 0u, m3 = 0u;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      m0 |= (v[j] <= st.thr) ? (1u << j) : 0u;
      m1 |= (v[8 + j] <= st.thr) ? (1u << (8 + j)) : 0u;
      m2 |= (v[16 + j] <= st.thr) ? (1u << (16 + j)) : 0u;
      m3 |= (v[24 + j] <= st.thr) ? (1u << (24 + j)) : 0u;
    }
    lo = (m0 | m1) | (m2 | m3);
  }
  if (mb <= st.thr) {
    unsigned m0 = 0u, m1 = 0u, m2 = 0u, m3 = 0u;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      m0 |= (v[32 + j] <= st.thr) ? (1u << j) : 0u;
      m1 |= (v[40 + j] <= st.thr) ? (1u << (8 + j)) : 0u;
      m2 |= (v[48 + j] <= st.thr) ? (1u << (16 + j)) : 0u;
      m3 |= (v[56 + j] <= st.thr) ? (1u << (24 + j)) : 0u;
    }
    hi = (m0 | m1) | (m2 | m3);
  }
  const int64_t c0 = st.clip0 + coff;
  if (__popc(lo) + __popc(hi) == 1) {
    const int idx = lo ? (__ffs(lo) - 1) : (31 + __ffs(hi));
    sl_append(st, m, c0 + idx, p);
  } else {
    float vl[64];
#pragma unroll
    for (int j = 0; j < 64; ++j) vl[j] = v[j];
    while (lo) {
      const int idx = __ffs(lo) - 1;
      lo &= lo - 1;
      sl_append(st, vl[idx], c0 + idx, p);
    }
    while (hi) {
      const int idx = __ffs(hi) - 1;
      hi &= hi - 1;
      sl_append(st, vl[32 + idx], c0 + 32 + idx, p);
    }
  }
}

// 64 accumulator columns of one query row: FMNMX3 tree and one compare
__device__ __forceinline__ void sl_process(const float (&v)[64], int coff, SlRow& st, const SlParams& p) {
  float g[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const float a = fmin3(v[8 * i], v[8 * i + 1], v[8 * i + 2]);
    const float b = fmin3(v[8 * i + 3], v[8 * i + 4], v[8 * i + 5]);
    g[i] = fmin3(a, b, fminf(v[8 * i + 6], v[8 * i + 7]));
  }
  const float ma = fmin3(g[0], g[1], fminf(g[2], g[3]));
  const float mb = fmin3(g[4], g[5], fminf(g[6], g[7]));
  const float m = fminf(ma, mb);
  if (m <= st.thr) sl_cold(v, ma, mb, m, coff, st, p);
}

// Sample pass (MODE 1): the SL_J smallest 64-column minima of the row, kept sorted in registers
constexpr int SL_J = 32;
__device__ __forceinline__ void sl_sample(const float (&v)[64], float (&a)[SL_J]) {
  float g[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const float x = fmin3(v[8 * i], v[8 * i + 1], v[8 * i + 2]);
    const float y = fmin3(v[8 * i + 3], v[8 * i + 4], v[8 * i + 5]);
    g[i] = fmin3(x, y, fminf(v[8 * i + 6], v[8 * i + 7]));
  }
  float m = fmin3(fmin3(g[0], g[1], g[2]), fmin3(g[3], g[4], g[5]), fminf(g[6], g[7]));
#pragma unroll
  for (int i = 0; i < SL_J; ++i) {
    const float lo = fminf(a[i], m);
    m = fmaxf(a[i], m);
    a[i] = lo;
  }
}

// CL = 2: CTA pairs (thread-block cluster) of the same bank split share every bank tile: each CTA fetches one of
// the two 32 KB boxes and TMA multicasts it into both shared memories, halving the L2 reads and the TMA
// requests per SM; a stage is recycled when the MMA warps of BOTH CTAs have released it.
// BIG: the K-streaming variant for D + 3 > 128 (R = 2, CL = 1): every 64-column chunk of the two query tiles travels
// through the ring together with the bank tile's chunk; a tile's two accumulators are committed after the last chunk
// (no ping-pong - the epilogue's ~700-cycle read of a buffer is small against the >= 3 x 4 MMAs x 2 of a tile).
// Accumulator schemes (NB, VFR_SEL_NB; all return the same bits).  The default is NB = 8: the two 256-column accumulators
// of NB = 2, each with its OWN issuer warp that commits a job right after its MMAs.  With one issuer the commit of job j is
// deferred behind the probes of job j + 1 (a barrier wait right behind a tcgen05.commit stalls the thread until the MMAs
// ahead have drained - and the one issuer has the other accumulator's job to issue meanwhile), so the epilogue learns of
// a finished job 250 - 600 cycles late; an issuer of its own has nothing else to do until its accumulator comes back, the
// stall costs it nothing.  Measured in the bench step (37 888 queries x 1 M videos, same box, back to back): scan 49.5 ->
// 43.3 ms, step 76.1 -> 70.7 ms.  The history of the other schemes:  A buffer's
// round trip - MMAs issued -> commit seen by the epilogue (~550 cycles) -> its four warps have read it (~750 median, 1 200
// p90) -> the issuer sees the release (~450) - is ~3 000 cycles, i.e. 1 500 per job with two buffers against 896 of MMA
// (tools/timeline_sel.py).  Smaller jobs on more buffers would spread the same latencies over more work in flight:
//   NB = 4  four accumulators of 128 columns, a job = one query tile x one HALF of a bank tile (MMA N = 128);
//   NB = 5  the same with ONE MMA ISSUER PER QUERY TILE (warps 1 and 2, 352 threads);
//   NB = 3  the query tiles kept in TMEM (tcgen05.st once per CTA, MMA with the A operand from TMEM: no re-read of the
//           128 x 16 A slice from shared memory per MMA, a sixth ring stage), three accumulators of 128 columns in rotation.
// Measured on B200, 37 888 queries x 6 M clips, whole vfr_sel_topk, same box: NB = 2 49.7 ms, NB = 5 50.9, NB = 3 61.6,
// NB = 4 62.7.  What round 2 found behind the first NB = 4 result (70 ms, "seven N = 128 MMAs take as long as seven N = 256
// ones"): not the tensor pipe - in isolation N = 128 runs at its 64-cycle floor, dependent accumulation included
// (tools/ubench_tc.cu dep) - but the ISSUING THREAD.  Issued one by one from `if (lane == 0)` code, every tcgen05.mma is
// wrapped by the compiler in an ELECT / BRA.U.ANY loop with its operands moved by R2UR: ~120 cycles of issue per MMA
// whatever N.  The issuer now runs as a converged warp and issues each chunk's MMAs from one asm block (sl_mma_block:
// 7 MMAs in ~640 cycles, 1 100 before).  That moved NB = 4 from 77.8 to 62.7 ms and NB = 3 from 72 to 61.6, and left NB = 2
// where it was: with N = 256 the chain (896 cycles of MMA + the two hand-offs + the read) binds, not the issue.  With half
// jobs the issuer's fixed cost per job (~280 cycles of probes / commit / fence + the issue) is paid twice as often - one
// issuer is bound by it (NB = 3 / 4), two issuers (NB = 5) reach the default's time but then wait for bank tiles (every
// stage is released by both) - so N = 256 on two buffers stays the default.  Also measured and dropped: all eight
// epilogue warps reading every buffer, 128 columns each (hold 740 -> 515 cycles, but every warp then pays the fixed costs
// of every job and the lists double: 64.6 ms).
template <int R, int CL, int MODE, bool BIG = false, int NB = 2>
__global__ void __launch_bounds__((NB == 5 || NB == 8) ? SL_THREADS2 : SL_THREADS, 1)
sl_filter_kernel(const __grid_constant__ CUtensorMap tm_a, const __grid_constant__ CUtensorMap tm_b, const SlParams p) {
  static_assert(!BIG || (R == 2 && CL == 1), "the K-streaming variant serves two query tiles per CTA, no cluster");
  static_assert(NB == 2 || ((NB == 4 || NB == 3 || NB == 5 || NB == 8) && !BIG && CL == 1), "three / four accumulators: plain kernel only");
  static_assert((NB != 3 && NB != 5 && NB != 8) || R == 2, "query tiles in TMEM / two issuers: two query tiles per CTA");
  constexpr bool ATM = NB == 3;                       // A operand (the query tiles) in TMEM, three accumulators of 128 columns
  constexpr bool ISS2 = NB == 5 || NB == 8;           // ONE MMA ISSUER PER QUERY TILE (warps 1 and 2); NB = 5: four accumulators
  // NB = 8: the two 256-column accumulators of NB = 2, each with its OWN issuer, which commits a job right after its MMAs:
  // the stall of a barrier wait behind a commit (until the MMAs ahead have drained) costs an issuer nothing when its next
  // job needs the same accumulator anyway, and the epilogue learns of a finished job without the delay of the deferred commit
  constexpr bool OWN = NB == 8;
  constexpr int NBUF = OWN ? 2 : (ISS2 ? 4 : NB);
  constexpr int EW0 = ISS2 ? 3 : 2;                   // first epilogue warp
  constexpr bool HALF = NB >= 3 && !OWN;
  constexpr int JOB_N = HALF ? SL_N / 2 : SL_N;       // accumulator columns of a job
  constexpr int STAGES = BIG ? SL_BIG_STAGES : (ATM ? SL_ATM_STAGES : SlCfg<R>::STAGES);
  constexpr int STAGE_BYTES = BIG ? SL_BIG_STAGE : SL_B_CHUNK;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* smem_a = smem;                                   // [R][2 chunks]   (BIG: unused, the chunks live in the stages)
  uint8_t* smem_b = (BIG || ATM) ? smem : smem + R * 2 * SL_A_CHUNK; // [STAGES]
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_b + STAGES * STAGE_BYTES);
  uint64_t* full = bars;                        // [STAGES]
  uint64_t* empty = bars + STAGES;              // [STAGES]
  uint64_t* a_full = bars + 2 * STAGES;         // [1]
  uint64_t* tmem_full = a_full + 1;             // [NB]
  uint64_t* tmem_empty = tmem_full + 4;         // [NB]
  uint32_t* tmem_base_slot = reinterpret_cast<uint32_t*>(tmem_empty + 4);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int qgroup = blockIdx.x % p.n_qgroups;
  const int split = blockIdx.x / p.n_qgroups;
  const int tile_begin = p.tile_lo + split * p.tiles_per_split;
  const int tile_end = min(tile_begin + p.tiles_per_split, p.n_tiles);
  const int n_my_tiles = max(tile_end - tile_begin, 0);
  const int b_chunks = (p.ksteps > 4) ? 2 : 1;
  // The CTAs of a bank split would walk the same bank tiles in step - every tile asked for by all of them at the same
  // moment.  p.rot_step > 0 rotates the scan of CTA group (blockIdx.x % 8) by that many tiles (visit t reads tile
  // tile_of(t)); the order in which a list sees the clips does not change what stage 2 returns.
  const int rot = (BIG || MODE == 1 || p.rot_step <= 0 || n_my_tiles < 64) ? 0 : (int)(((blockIdx.x % 8) * (unsigned)p.rot_step) % (unsigned)n_my_tiles);
  auto tile_of = [&](int t) { const int u = t + rot; return u < n_my_tiles ? u : u - n_my_tiles; };

  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], ISS2 ? 2 : CL); }
    mbar_init(a_full, ATM ? 2 * SL_SET_WARPS : 1);      // (ATM: every epilogue warp stores its 32 query rows into TMEM)
    for (int b = 0; b < NBUF; ++b) { mbar_init(&tmem_full[b], 1); mbar_init(&tmem_empty[b], SL_SET_WARPS); }
    fence_barrier_init();
  }
  const int crank = (CL == 2) ? sl_cluster_rank() : 0;
  if (CL == 2) sl_cluster_sync();          // the peer's barriers exist before anything is multicast to them
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_base_slot)),
                 "r"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  sl_fence_before();
  __syncthreads();
  sl_fence_after();
  // (through a shuffle: the compiler then knows the value is warp-uniform and keeps the accumulator addresses of the MMA
  //  issuer in uniform registers - from a plain shared-memory load every tcgen05.mma is wrapped in an ELECT / R2UR /
  //  BRA.U.ANY loop, ~120 cycles of issue per MMA, which bounds jobs of N = 128)
  const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_base_slot, 0);

  if (warp == 0) {
    // ================= TMA producer =================
    if (BIG) {
      if (lane == 0 && n_my_tiles > 0) {
        int it = 0;
        for (int t = 0; t < n_my_tiles; ++t) {
          const int row = (tile_begin + t) * p.tile_stride * SL_N;
          for (int c = 0; c < p.n_chunks; ++c, ++it) {
            const int s = it % STAGES;
            sl_wait(&empty[s], ((it / STAGES) & 1) ^ 1, p.wait_mode);
            uint8_t* st = smem_b + s * STAGE_BYTES;
            mbar_expect_tx(&full[s], STAGE_BYTES);
            for (int r = 0; r < R; ++r) sl_tma_load(st + r * SL_A_CHUNK, &tm_a, c * 64, (qgroup * R + r) * SL_M, &full[s]);
            sl_tma_load(st + R * SL_A_CHUNK, &tm_b, c * 64, row, &full[s]);
          }
        }
      }
    } else
    if (lane == 0 && n_my_tiles > 0) {
      if constexpr (!ATM) {
        mbar_expect_tx(a_full, (uint32_t)(R * b_chunks) * SL_A_CHUNK);
        for (int r = 0; r < R; ++r)
          for (int c = 0; c < b_chunks; ++c)
            sl_tma_load(smem_a + (r * 2 + c) * SL_A_CHUNK, &tm_a, c * 64, (qgroup * R + r) * SL_M, a_full);
      }
      int it = 0;
      for (int t = 0; t < n_my_tiles; ++t) {
        const int row = (tile_begin + tile_of(t)) * p.tile_stride * SL_N;
        // The CTAs of a bank split walk the same tiles nearly in step, so whoever comes first pays the DRAM latency of a tile
        // and the ring (2.5 tiles) is not deep enough to hide it: the issuer waited ~900 cycles per tile for the bank tile
        // (tools/timeline_sel.py).  One CTA in SL_PF_SHARE asks L2 for the tile p.prefetch tiles ahead.
        if (p.prefetch > 0 && t + p.prefetch < n_my_tiles && ((tile_begin + t) % SL_PF_SHARE) == (int)(blockIdx.x % SL_PF_SHARE)) {
          const int prow = (tile_begin + tile_of(t + p.prefetch)) * p.tile_stride * SL_N;
          for (int c = 0; c < b_chunks; ++c) sl_tma_prefetch(&tm_b, c * 64, prow);
        }
        for (int c = 0; c < b_chunks; ++c, ++it) {
          const int s = it % STAGES;
          sl_wait(&empty[s], ((it / STAGES) & 1) ^ 1, p.wait_mode);
          if (p.dbg && p.prod_dbg && blockIdx.x == 0 && t >= n_my_tiles - 128) p.dbg[(8 * 256 + (t - (n_my_tiles - 128))) * 4 + c] = clock64();
          mbar_expect_tx(&full[s], SL_B_CHUNK);
          if (CL == 1) sl_tma_load(smem_b + s * SL_B_CHUNK, &tm_b, c * 64, row, &full[s]);
          else if ((it & 1) == crank) sl_tma_load_mc(smem_b + s * SL_B_CHUNK, &tm_b, c * 64, row, &full[s], (uint16_t)3);
        }
      }
    }
  } else if (warp == 1 || (ISS2 && warp == 2)) {
    // ================= MMA issuer =================
    if (BIG) {
      if (lane == 0 && n_my_tiles > 0) {
        const uint32_t idesc = (1u << 4) | ((uint32_t)(SL_N >> 3) << 17) | ((uint32_t)(SL_M >> 4) << 24);
        int it = 0;
        for (int t = 0; t < n_my_tiles; ++t) {
          for (int r = 0; r < R; ++r) sl_wait(&tmem_empty[r], (t & 1) ^ 1, p.wait_mode);
          sl_fence_after();
          for (int c = 0; c < p.n_chunks; ++c, ++it) {
            const int s = it % STAGES;
            sl_wait(&full[s], (it / STAGES) & 1, p.wait_mode);
            sl_fence_after();
            uint8_t* st = smem_b + s * STAGE_BYTES;
            const int ks = min(4, p.ksteps - 4 * c);
            const uint64_t bdesc = sl_desc(st + R * SL_A_CHUNK);
            for (int r = 0; r < R; ++r) {
              const uint64_t adesc = sl_desc(st + r * SL_A_CHUNK);
              const uint32_t d_tmem = tmem_base + (uint32_t)r * SL_N;
              for (int k = 0; k < ks; ++k)
                sl_mma(d_tmem, adesc + (uint64_t)(k * 2), bdesc + (uint64_t)(k * 2), idesc, (c | k) ? 1u : 0u);
            }
            sl_commit(&empty[s]);                  // the stage is free once these MMAs have read it
          }
          for (int r = 0; r < R; ++r) sl_commit(&tmem_full[r]);
        }
      }
    } else
    if (n_my_tiles > 0) {
      // (the whole warp runs this loop converged; the tcgen05 instructions come from its elected lane - see sl_elect)
      // kind::f16, fp16 x fp16 -> fp32, K-major A and B, N = 256 (128 with four accumulators), M = 128
      const uint32_t idesc = (1u << 4) | ((uint32_t)(JOB_N >> 3) << 17) | ((uint32_t)(SL_M >> 4) << 24);
      sl_wait(a_full, 0, p.wait_mode);
      sl_fence_after();
      // A barrier wait issued right after a tcgen05.commit stalls the thread ~250 cycles (until the MMAs ahead
      // of the commit have drained), which is a third of a job.  So the commit of job j is issued AFTER the
      // waits of job j+1 - unless one of them really has to block, then the commit goes first.
      int it = 0;
      int s0 = 0, s1 = 0, ph0 = 0, ph1 = 0;
      int pend_buf = -1, pend_s0 = 0, pend_s1 = 0;
      bool pend_release = false;
      constexpr int JPT = OWN ? 1 : (ISS2 ? 2 : (HALF ? 2 : 1) * R);     // jobs per bank tile (of this issuer)
      const int n_jobs = n_my_tiles * JPT;
      for (int job = 0; job < n_jobs; ++job) {
        // two accumulators: buffer = query tile (R = 2) or tile parity (R = 1).  Four: a tile's jobs go (r0, h0), (r1, h0),
        // (r0, h1), (r1, h1) - the sets get their work evenly spaced - on buffer 2 r + h; R = 1: tile parity s, 2 s + h.
        int r, h, buf, use;
        bool first, last;
        if (OWN) {
          r = warp - 1; h = 0; buf = r; use = job; first = true; last = true;
        } else if (ISS2) {
          // this issuer serves query tile r = warp - 1 alone: the two halves of every bank tile on its own two accumulators
          r = warp - 1; h = job & 1; buf = 2 * r + h; use = job >> 1; first = h == 0; last = h == 1;
        } else if (ATM) {
          // three accumulators in rotation; jobs of a tile as below, so each epilogue set gets every other job
          const int j4 = job & 3; r = j4 & 1; h = j4 >> 1; buf = job % 3; use = job / 3; first = j4 == 0; last = j4 == 3;
        } else if (HALF) {
          if (R == 2) { const int j4 = job & 3; r = j4 & 1; h = j4 >> 1; buf = 2 * r + h; use = job >> 2; first = j4 == 0; last = j4 == 3; }
          else { const int tt = job >> 1; r = 0; h = job & 1; buf = 2 * (tt & 1) + h; use = tt >> 1; first = h == 0; last = h == 1; }
        } else {
          r = (R == 2) ? (job & 1) : 0; h = 0; buf = job & 1; use = job >> 1; first = r == 0; last = r == R - 1;
        }
        const long long tm0 = p.dbg ? clock64() : 0;
        long long tm_full = tm0, tm_commit = tm0;   // (development aid) when the pending commit was out / the bank tile of this job was there
        if (first) {
          s0 = it % STAGES; ph0 = (it / STAGES) & 1;
          s1 = (it + 1) % STAGES; ph1 = ((it + 1) / STAGES) & 1;
          it += b_chunks;
        }
        const uint32_t pe = (use & 1) ^ 1;
        bool ready = sl_test(&tmem_empty[buf], pe);
        if (first) {
          ready = ready && sl_test(&full[s0], ph0);
          if (b_chunks == 2) ready = ready && sl_test(&full[s1], ph1);
        }
        ready = __all_sync(0xffffffffu, ready);       // (one answer for the warp: the lanes probe at slightly different times)
        if (!ready || pend_buf < 0) {
          if (pend_buf >= 0) {
            if (sl_elect()) {
              sl_commit(&tmem_full[pend_buf]);
              if (pend_release) { sl_release<CL>(&empty[pend_s0]); if (b_chunks == 2) sl_release<CL>(&empty[pend_s1]); }
            }
            pend_buf = -1;
          }
          if (p.dbg) tm_commit = clock64();
          if (first) {
            sl_wait(&full[s0], ph0, p.wait_mode);
            if (b_chunks == 2) sl_wait(&full[s1], ph1, p.wait_mode);
          }
          if (p.dbg) tm_full = clock64();
          sl_wait(&tmem_empty[buf], pe, p.wait_mode);
        }
        sl_fence_after();
        if (pend_buf >= 0 && sl_elect()) {
          sl_commit(&tmem_full[pend_buf]);
          if (pend_release) { sl_release<CL>(&empty[pend_s0]); if (b_chunks == 2) sl_release<CL>(&empty[pend_s1]); }
        }
        const long long tm1 = p.dbg ? clock64() : 0;
        const uint32_t d_tmem = tmem_base + (uint32_t)buf * JOB_N;
        // straight-line issue: compile-time k and accumulate flag, and every operand made KNOWN warp-uniform by a shuffle
        // (the job's bookkeeping lives in vector registers; without it each UTCHMMA is preceded by ~7 R2UR and two VOTEU) -
        // what sits between two UTCHMMAs is issue time of the one thread that feeds the tensor pipe
        const uint32_t d_u = __shfl_sync(0xffffffffu, d_tmem, 0);
        const uint32_t b0_lo = __shfl_sync(0xffffffffu, (uint32_t)sl_desc(smem_b + s0 * SL_B_CHUNK + h * (SL_B_CHUNK / 2)), 0);
        const uint32_t b1_lo = __shfl_sync(0xffffffffu, (uint32_t)sl_desc(smem_b + s1 * SL_B_CHUNK + h * (SL_B_CHUNK / 2)), 0);
        const uint32_t a_lo = ATM ? __shfl_sync(0xffffffffu, tmem_base + SL_ATM_ACOL + (uint32_t)(r * 64), 0)
                                  : __shfl_sync(0xffffffffu, (uint32_t)sl_desc(smem_a + r * 2 * SL_A_CHUNK), 0);
        constexpr uint64_t DESC_HI = (uint64_t)(sl_desc_hi()) << 32;
#pragma unroll
        for (int c = 0; c < 2; ++c) {
          if (c < b_chunks) {
            const int ks = (c == 0) ? min(p.ksteps, 4) : (p.ksteps - 4);
            const uint64_t bdesc = DESC_HI | (c ? b1_lo : b0_lo);
            const uint32_t acc = c ? 1u : 0u;
            if constexpr (ATM) {
              // 16 fp16 of K = 8 TMEM columns; a 64-column chunk of the packed row = 32 TMEM columns
              const uint32_t a_t = a_lo + (uint32_t)(c * 32);
              if (ks >= 4) sl_mma_ts_block<4>(d_u, a_t, bdesc, idesc, acc);
              else if (ks == 3) sl_mma_ts_block<3>(d_u, a_t, bdesc, idesc, acc);
              else if (ks == 2) sl_mma_ts_block<2>(d_u, a_t, bdesc, idesc, acc);
              else if (ks == 1) sl_mma_ts_block<1>(d_u, a_t, bdesc, idesc, acc);
            } else {
              const uint64_t adesc = DESC_HI | (a_lo + (uint32_t)(c * (SL_A_CHUNK >> 4)));
              if (ks >= 4) sl_mma_block<4>(d_u, adesc, bdesc, idesc, acc);
              else if (ks == 3) sl_mma_block<3>(d_u, adesc, bdesc, idesc, acc);
              else if (ks == 2) sl_mma_block<2>(d_u, adesc, bdesc, idesc, acc);
              else if (ks == 1) sl_mma_block<1>(d_u, adesc, bdesc, idesc, acc);
            }
          }
        }
        pend_buf = buf;
        pend_s0 = s0;
        pend_s1 = s1;
        pend_release = last;
        if constexpr (OWN) {
          if (sl_elect()) {
            sl_commit(&tmem_full[pend_buf]);
            sl_release<CL>(&empty[pend_s0]);
            if (b_chunks == 2) sl_release<CL>(&empty[pend_s1]);
          }
          pend_buf = -1;
        }
        const int job0 = n_jobs - 256;
        if (p.dbg && blockIdx.x == 0 && job >= job0 && lane == 0 && warp == 1) {
          long long* d = p.dbg + (0 * 256 + (job - job0)) * 4;
          d[0] = tm0; d[1] = tm1; d[2] = clock64(); d[3] = buf | ((tm_commit - tm0) << 8) | ((tm_full - tm0) << 32);
        }
      }
      if (pend_buf >= 0 && sl_elect()) {
        sl_commit(&tmem_full[pend_buf]);
        if (pend_release) { sl_release<CL>(&empty[pend_s0]); if (b_chunks == 2) sl_release<CL>(&empty[pend_s1]); }
      }
      __syncwarp();
    }
  } else {
    // ================= epilogue: two sets of four warps, set s owns TMEM buffer s =================
    const int ew = warp - EW0;
    const int quarter = warp & 3;                 // TMEM lane quarter this warp may touch
    const int set = ew >> 2;
    const uint32_t lane_addr = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(set * SL_N);
    // R = 2: set s serves query tile s on every bank tile; R = 1: set s serves the tiles of parity s
    const int qtile = (R == 2) ? (qgroup * 2 + set) : qgroup;
    const int part = (R == 2) ? split : (split * 2 + set);
    const int t_first = (R == 2) ? 0 : set;
    const int t_step = (R == 2) ? 1 : 2;
    if constexpr (ATM) {
      // this thread's packed query row (128 fp16 = 64 words) -> its TMEM lane, columns of query tile `set`
      const uint32_t a_addr = tmem_base + ((uint32_t)(quarter * 32) << 16) + SL_ATM_ACOL + (uint32_t)(set * 64);
      const uint4* src = reinterpret_cast<const uint4*>(p.qpack + ((int64_t)qtile * SL_M + quarter * 32 + lane) * (SL_ROW / 2));
#pragma unroll 1
      for (int half = 0; half < 2; ++half) {
        uint32_t w[32];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const uint4 x = __ldg(src + half * 8 + i);
          w[4 * i] = x.x; w[4 * i + 1] = x.y; w[4 * i + 2] = x.z; w[4 * i + 3] = x.w;
        }
        sl_st32(a_addr + (uint32_t)(half * 32), w);
      }
      asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
      sl_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(a_full);
    }
    if constexpr (MODE == 1) {
      // ---- sample pass: no lists, only the sample_j-th smallest of the 64-column minima of a strided sample of
      //      the bank.  It becomes the query's STARTING threshold (stage 2 checks that the bank really holds k
      //      keys under it, see sl_refine_kernel), so the scan proper skips the phase in which everything passes ----
      const int64_t q = (int64_t)qtile * SL_M + quarter * 32 + lane;
      float a[SL_J];
#pragma unroll
      for (int i = 0; i < SL_J; ++i) a[i] = CUDART_INF_F;
      int visit = 0;
      for (int t = t_first; t < n_my_tiles; t += t_step, ++visit) {
        if constexpr (HALF) {
#pragma unroll 1
          for (int h = 0; h < 2; ++h) {
            int buf = 2 * set + h, par = visit & 1;
            uint32_t la = lane_addr + (uint32_t)h * JOB_N;
            if constexpr (ATM) {
              const int j = 4 * visit + 2 * h + set;        // the issuer's job counter
              buf = j % 3; par = (j / 3) & 1;
              la = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(buf * JOB_N);
            }
            sl_wait(&tmem_full[buf], par, p.wait_mode);
            sl_fence_after();
            float va[64], vb[64];
            sl_ld32(la, va, 0);
            sl_ld32(la + 32, va, 32);
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            sl_ld32(la + 64, vb, 0);
            sl_ld32(la + 96, vb, 32);
            sl_sample(va, a);
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            sl_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&tmem_empty[buf]);
            sl_sample(vb, a);
          }
          continue;
        }
        sl_wait(&tmem_full[set], visit & 1, p.wait_mode);
        sl_fence_after();
        float va[64], vb[64];
        sl_ld32(lane_addr, va, 0);
        sl_ld32(lane_addr + 32, va, 32);
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        sl_ld32(lane_addr + 64, vb, 0);
        sl_ld32(lane_addr + 96, vb, 32);
        sl_sample(va, a);
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        sl_ld32(lane_addr + 128, va, 0);
        sl_ld32(lane_addr + 160, va, 32);
        sl_sample(vb, a);
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        sl_ld32(lane_addr + 192, vb, 0);
        sl_ld32(lane_addr + 224, vb, 32);
        sl_sample(va, a);
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        sl_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&tmem_empty[set]);
        sl_sample(vb, a);
      }
      float x = CUDART_INF_F;
#pragma unroll
      for (int i = 0; i < SL_J; ++i)
        if (i == p.sample_j - 1) x = a[i];
      if (q < p.n_queries) {
        const float4 qm = __ldg(p.qmeta + q);
        if (x < CUDART_INF_F) tau_publish(p.tau_g + q, fmaxf(__fmaf_ru(x, qm.z, qm.x), 0.f));
        if (p.samp) {
          // for a bank sharded over several GPUs: the sampled minima as upper bounds of exact squared distances
          // (d2~ + E), so that values of shards with different operand scales compare
          float* out = p.samp + ((int64_t)q * p.n_parts + part) * SL_J;
#pragma unroll
          for (int i = 0; i < SL_J; ++i)
            out[i] = (a[i] < CUDART_INF_F) ? __fadd_ru(fmaxf(__fmaf_ru(a[i], qm.z, qm.x), 0.f), 0.5f * qm.w) : CUDART_INF_F;
        }
      }
    } else {
    SlRow st;
    st.q = (int64_t)qtile * SL_M + quarter * 32 + lane;
    const bool valid = st.q < p.n_queries;
    st.list = p.cand + ((int64_t)st.q * p.n_parts + part) * SL_CAP;
    st.cnt = p.resume ? p.cand_cnt[st.q * p.n_parts + part] : 0;
    {
      const float4 qm = __ldg(p.qmeta + st.q);
      st.nq = qm.x; st.scale = qm.y; st.inv_scale = qm.z; st.band2 = qm.w;
    }
    st.trig = p.cap_trig;
    st.tau_use = valid ? CUDART_INF_F : -1.f;
    st.thr = valid ? CUDART_INF_F : -CUDART_INF_F;

    int visit = 0;
    for (int t = t_first; t < n_my_tiles; t += t_step, ++visit) {
      st.clip0 = (int64_t)(tile_begin + tile_of(t)) * SL_N;
      if ((visit & 3) == 0 && (p.n_parts > 1 || visit == 0)) {
        // the other lists of this query (other bank splits / tile parities) may have tightened the threshold
        const float tg = tau_fetch(p.tau_g + st.q);
        if (tg < st.tau_use) {
          st.tau_use = tg;
          st.thr = sl_threshold(tg, st.band2, st.nq, st.scale);
        }
      }
      if constexpr (HALF) {
#pragma unroll 1
        for (int h = 0; h < 2; ++h) {
          int buf = 2 * set + h, par = visit & 1;
          uint32_t la = lane_addr + (uint32_t)h * JOB_N;
          if constexpr (ATM) {
            const int j = 4 * visit + 2 * h + set;          // the issuer's job counter
            buf = j % 3; par = (j / 3) & 1;
            la = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(buf * JOB_N);
          }
          const long long te0 = p.dbg ? clock64() : 0;
          sl_wait(&tmem_full[buf], par, p.wait_mode);
          sl_fence_after();
          const long long te1 = p.dbg ? clock64() : 0;
          float va[64], vb[64];
          sl_ld32(la, va, 0);
          sl_ld32(la + 32, va, 32);
          asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
          sl_ld32(la + 64, vb, 0);             // in flight while the first 64 columns are processed
          sl_ld32(la + 96, vb, 32);
          sl_process(va, h * JOB_N, st, p);
          asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
          // the whole accumulator has been read: hand it back before the last 64 columns (in registers) are looked at
          sl_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&tmem_empty[buf]);
          const long long te2 = p.dbg ? clock64() : 0;
          sl_process(vb, h * JOB_N + 64, st, p);
          const int visit0 = 2 * ((n_my_tiles - t_first + t_step - 1) / t_step) - 128;
          if (p.dbg && blockIdx.x == 0 && 2 * visit + h >= visit0 && lane == 0) {
            long long* d = p.dbg + ((1 + ew) * 256 + (2 * visit + h - visit0)) * 4;
            d[0] = te0; d[1] = te1; d[2] = te2; d[3] = clock64();
          }
        }
      } else {
      const long long te0 = p.dbg ? clock64() : 0;
      sl_wait(&tmem_full[set], visit & 1, p.wait_mode);
      sl_fence_after();
      const long long te1 = p.dbg ? clock64() : 0;
      float va[64], vb[64];
      sl_ld32(lane_addr, va, 0);
      sl_ld32(lane_addr + 32, va, 32);
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      sl_ld32(lane_addr + 64, vb, 0);          // in flight while the first 64 columns are processed
      sl_ld32(lane_addr + 96, vb, 32);
      sl_process(va, 0, st, p);
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      sl_ld32(lane_addr + 128, va, 0);
      sl_ld32(lane_addr + 160, va, 32);
      sl_process(vb, 64, st, p);
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      sl_ld32(lane_addr + 192, vb, 0);
      sl_ld32(lane_addr + 224, vb, 32);
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      // the whole accumulator has been read: hand the TMEM buffer back to the MMA warp BEFORE the last two
      // 64-column groups (both in registers) are looked at - their hits then cost no hold time
      sl_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tmem_empty[set]);
      const long long te2 = p.dbg ? clock64() : 0;
      sl_process(va, 128, st, p);
      sl_process(vb, 192, st, p);
      const int visit0 = (n_my_tiles - t_first + t_step - 1) / t_step - 128;
      if (p.dbg && blockIdx.x == 0 && visit >= visit0 && lane == 0 && !(p.prod_dbg && ew == 7)) {
        long long* d = p.dbg + ((1 + ew) * 256 + (visit - visit0)) * 4;
        d[0] = te0; d[1] = te1; d[2] = te2; d[3] = clock64();
      }
      }

      // Every compaction refreshes the list's threshold (the k-th smallest key seen so far); a stale threshold costs
      // hits - each one in the first two 64-column groups lengthens the hold of the TMEM buffer by ~450 cycles - a
      // compaction ~10 k cycles.  The first one comes after cap_trig keys (the sampled starting threshold is loose),
      // the following ones cap_step keys above what the last one kept (default: when the list is full).
      const bool need = st.cnt > st.trig;
      if (__any_sync(0xffffffffu, need)) {
        const long long tc0 = p.dbg ? clock64() : 0;
        const int n_need = __popc(__ballot_sync(0xffffffffu, need));
        if (lane == 0) { atomicAdd(p.stats, 1ull); atomicAdd(p.stats + 1, (unsigned long long)n_need); }
        // (copies: taking the address of a member would push the whole per-thread state into local memory)
        int cnt = st.cnt;
        float tau_own = CUDART_INF_F;
        float tau_part = CUDART_INF_F;
        const int k_part = (p.k + p.n_parts - 1) / p.n_parts;
        sl_compact(st.list, cnt, tau_own, tau_part, st.tau_use, st.band2, p.k, k_part, need, lane, p.dbg);
        const bool over = cnt > SL_CAP_HI;
        if (__any_sync(0xffffffffu, over)) {
          // more than CAP_HI keys inside the band (mass duplicates): keep the list bounded and flag the query
          if (over) p.flags[st.q] = 2;
          float dummy = CUDART_INF_F;
          sl_compact(st.list, cnt, tau_own, dummy, st.tau_use, 0.f, p.k, p.k, over, lane);
          if (over) cnt = min(cnt, SL_CAP_HI);
        }
        if (need) st.trig = min(SL_CAP_HI, cnt + p.cap_step);
        st.cnt = cnt;
        if (tau_own < CUDART_INF_F) {
          float best = tau_own;
          if (p.n_parts > 1 && tau_part < CUDART_INF_F) {
            // publish this list's share and bound the query's k-th smallest by the largest share of all its lists
            volatile float* tp = p.tau_part + st.q * p.n_parts;
            tp[part] = tau_part;
            float worst = tau_part;
            for (int j = 0; j < p.n_parts; ++j) worst = fmaxf(worst, tp[j]);
            best = fminf(best, worst);
          }
          tau_publish(p.tau_g + st.q, best);
          tau_publish(p.tau_cert + st.q, best);
          if (best < st.tau_use) {
            st.tau_use = best;
            st.thr = sl_threshold(best, st.band2, st.nq, st.scale);
          }
        }
        if (p.dbg && lane == 0) {   // all CTAs: {clocks spent compacting, warp events, lists compacted}
          unsigned long long* d = reinterpret_cast<unsigned long long*>(p.dbg + 9 * 256 * 4);
          atomicAdd(d, (unsigned long long)(clock64() - tc0));
          atomicAdd(d + 1, 1ull);
          atomicAdd(d + 2, (unsigned long long)n_need);
        }
      }
    }
    // a list that ends with more than k keys still tightens the query's threshold for stage 2
    if (__any_sync(0xffffffffu, st.cnt > p.k)) {
      int cnt = st.cnt;
      float tau_own = CUDART_INF_F;
      float dummy = CUDART_INF_F;
      sl_compact(st.list, cnt, tau_own, dummy, st.tau_use, st.band2, p.k, p.k, cnt > p.k, lane);
      st.cnt = cnt;
      if (tau_own < CUDART_INF_F) {
        tau_publish(p.tau_g + st.q, tau_own);
        tau_publish(p.tau_cert + st.q, tau_own);
      }
    }
    p.cand_cnt[st.q * p.n_parts + part] = st.cnt;
    }   // MODE == 0
  }

  sl_fence_before();
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512));
  }
  if (CL == 2) sl_cluster_sync();          // the peer may still signal this CTA's barriers until it is done too
}

// ---------------------------------------------------------------------------------------------
// stage 2: exact refine - all moments of the candidate videos, arithmetic of vfr_score.cu
// ---------------------------------------------------------------------------------------------
constexpr int RF_THREADS = 256;
constexpr int RF_MAXC = 1024;      // candidate clips (and videos) per query
constexpr int RF_KEYS = 2048;      // sort buffer: [0, 128) running best, [128, 2048) the next batch
constexpr int RF_BATCH = RF_KEYS - VFR_TOPK_MAX;
constexpr int RF_DIST = 1024;      // clip distances per chunk of videos
constexpr int RF_FAST = 512;       // candidates per query up to which the counting (sort-free) path is taken
constexpr int RF_FINAL = 768;      // surviving moments up to which the final top-k is selected by counting

struct RfParams {
  const float* bank;           // fp32 [C, dim] ...
  const __nv_bfloat16* bank_b16; // ... or (bf16 embedding path, BASELINE configs[2]) the bank stored as bf16 [C, dim]
  const float* queries;        // fp32 [Q, dim]
  const int32_t* vid_off;      // [V+1]
  const int64_t* mom_off;      // [V+1]
  int64_t n_videos;
  int64_t n_clips;
  int n_max;
  int dim;
  const float4* qmeta;
  const unsigned long long* cand;
  const int32_t* cand_cnt;
  int n_parts;
  const unsigned* tau_g;
  const unsigned* tau_cert;
  int32_t* flags;
  int k;
  int64_t id_base;
  int64_t q0;                  // first query of this launch (the grid covers a query range)
  float* out_scores;
  int64_t* out_ids;
  // blocked output (sharded search, vfr_sel_refine_blocks): the queries are cut into slices of `per` (one per rank
  // of the exchange); slice j is ONE contiguous record {ids int64 [per, k] | scores fp32 [per, k] | flags int32 [per]}
  // at out_blocks + j * blk_bytes, so that a single all-to-all moves every rank's slice.  per = 0: flat [Q, k] arrays.
  int64_t per;
  int64_t blk_bytes;
  unsigned char* out_blocks;
  int fast_max;      // candidates per query up to which the counting path is taken (VFR_RF_FAST overrides; 0 = never)
  int final_max;
  unsigned long long* dbg;   // optional (VFR_RF_DBG = device address): clocks per phase, summed over all queries
};

// exact fp32 distance of one (query, clip) pair: the arithmetic of score_kernel / score_own_kernel in
// vfr_score.cu (direct-difference form, k strictly sequential).  The bank row is fetched in batches of
// 32 floats (8 independent 128-bit loads in flight) because the rows of the candidates are scattered
// over the whole bank: the loop is bound by DRAM latency, not by the 100 dependent FFMAs.
// the same arithmetic on a bank row stored in bf16 (exactly representable in fp32: the result equals the fp32 engine's on
// the rounded embeddings, bit for bit); 16 values per 256-bit of loads
__device__ __forceinline__ float rf_distance_b16(const __nv_bfloat16* __restrict__ vr, const float* qr, float sq, int dim) {
  float acc = 0.f, sv = 0.f;
  int k0 = 0;
  if ((reinterpret_cast<uintptr_t>(vr) & 15u) == 0) {
    for (; k0 + 32 <= dim; k0 += 32) {
      uint4 x[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) x[j] = __ldg(reinterpret_cast<const uint4*>(vr + k0) + j);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const unsigned w[4] = {x[j].x, x[j].y, x[j].z, x[j].w};
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const float lo = __uint_as_float(w[i] << 16), hi = __uint_as_float(w[i] & 0xffff0000u);
          float d = __fsub_rn(lo, qr[k0 + 8 * j + 2 * i]);
          acc = __fmaf_rn(d, d, acc);
          sv = __fadd_rn(sv, lo);
          d = __fsub_rn(hi, qr[k0 + 8 * j + 2 * i + 1]);
          acc = __fmaf_rn(d, d, acc);
          sv = __fadd_rn(sv, hi);
        }
      }
    }
  }
  for (int k = k0; k < dim; ++k) {
    const float x = __bfloat162float(vr[k]);
    const float d = __fsub_rn(x, qr[k]);
    acc = __fmaf_rn(d, d, acc);
    sv = __fadd_rn(sv, x);
  }
  const float corr = __fmaf_rn(2.f * VFR_PAIRWISE_EPS, __fsub_rn(sv, sq), (float)dim * VFR_PAIRWISE_EPS * VFR_PAIRWISE_EPS);
  return __fsqrt_rn(fmaxf(__fadd_rn(acc, corr), 0.f));
}

__device__ __forceinline__ float rf_distance(const float* __restrict__ vr, const float* qr, float sq, int dim) {
  float acc = 0.f, sv = 0.f;
  int k0 = 0;
  if ((reinterpret_cast<uintptr_t>(vr) & 15u) == 0) {
    for (; k0 + 32 <= dim; k0 += 32) {
      float4 x[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) x[j] = __ldg(reinterpret_cast<const float4*>(vr + k0) + j);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float xs[4] = {x[j].x, x[j].y, x[j].z, x[j].w};
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const float d = __fsub_rn(xs[i], qr[k0 + 4 * j + i]);
          acc = __fmaf_rn(d, d, acc);
          sv = __fadd_rn(sv, xs[i]);
        }
      }
    }
  }
  for (int k = k0; k < dim; ++k) {
    const float x = __ldg(vr + k);
    const float d = __fsub_rn(x, qr[k]);
    acc = __fmaf_rn(d, d, acc);
    sv = __fadd_rn(sv, x);
  }
  const float corr = __fmaf_rn(2.f * VFR_PAIRWISE_EPS, __fsub_rn(sv, sq), (float)dim * VFR_PAIRWISE_EPS * VFR_PAIRWISE_EPS);
  return __fsqrt_rn(fmaxf(__fadd_rn(acc, corr), 0.f));
}

// block-wide bitonic sort of keys[0, n), n a power of two
template <typename T>
__device__ __forceinline__ void rf_block_sort(T* keys, int n) {
  for (int size = 2; size <= n; size <<= 1) {
    for (int stride = size >> 1; stride > 0; stride >>= 1) {
      __syncthreads();
      for (int i = threadIdx.x; i < n / 2; i += RF_THREADS) {
        const int lo = ((i & ~(stride - 1)) << 1) | (i & (stride - 1));
        const int hi = lo | stride;
        const bool asc = !(lo & size) || size == n;
        const T a = keys[lo], b = keys[hi];
        if ((a > b) == asc) { keys[lo] = b; keys[hi] = a; }
      }
    }
  }
  __syncthreads();
}
__device__ __forceinline__ int rf_pow2(int x, int lo) {
  int n = lo;
  while (n < x) n <<= 1;
  return n;
}
// merge the batch keys[128, 128 + count) into the running best keys[0, k): sort, trim, new threshold
__device__ __forceinline__ void rf_merge_batch(unsigned long long* keys, int count, int k, int* s_cnt, float* s_tau) {
  const int n = rf_pow2(VFR_TOPK_MAX + count, 256);
  for (int i = VFR_TOPK_MAX + count + threadIdx.x; i < n; i += RF_THREADS) keys[i] = ~0ull;
  rf_block_sort(keys, n);
  for (int i = k + threadIdx.x; i < VFR_TOPK_MAX; i += RF_THREADS) keys[i] = ~0ull;
  if (threadIdx.x == 0) {
    *s_cnt = 0;
    const unsigned long long kth = keys[k - 1];
    if (kth != ~0ull) *s_tau = fminf(*s_tau, __uint_as_float((unsigned)(kth >> 32)));
  }
  __syncthreads();
}

__global__ void __launch_bounds__(RF_THREADS) sl_refine_kernel(const RfParams p) {
  __shared__ unsigned long long keys[RF_KEYS];
  __shared__ int vids[RF_MAXC];
  __shared__ int uvid[RF_MAXC];
  __shared__ float dist[RF_DIST];
  __shared__ float qrow[SL_MAXROW];
  __shared__ int u_c0[RF_MAXC];               // first clip of the unique candidate videos ...
  __shared__ unsigned char u_n[RF_MAXC];      // ... and their clip counts (no vid_off round trips in the scoring loops)
  __shared__ int warp_tot[RF_THREADS / 32];
  __shared__ int s_n, s_cnt;
  __shared__ float s_sq, s_tau;
  const int64_t q = p.q0 + blockIdx.x;
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  long long tq = p.dbg ? clock64() : 0;
  auto stamp = [&](int phase) {
    if (p.dbg && tid == 0) { const long long now = clock64(); atomicAdd(p.dbg + phase, (unsigned long long)(now - tq)); tq = now; }
  };

  // ---- 0. the query row and its sequential row sum ----
  for (int i = tid; i < p.dim; i += RF_THREADS) qrow[i] = p.queries[q * p.dim + i];
  if (tid == 0) { s_n = 0; s_cnt = 0; s_tau = CUDART_INF_F; }
  for (int i = tid; i < RF_MAXC; i += RF_THREADS) vids[i] = INT_MAX;
  for (int i = tid; i < VFR_TOPK_MAX; i += RF_THREADS) keys[i] = ~0ull;
  __syncthreads();
  if (tid == RF_THREADS - 32) {     // (a 100-step dependent chain: on a warp the candidate gather below leaves idle)
    float s = 0.f;
    for (int k = 0; k < p.dim; ++k) s = __fadd_rn(s, qrow[k]);
    s_sq = s;
  }

  // ---- 1. candidate clips.  The k-th smallest d2~ over the UNION of the query's lists (each list only
  //         knows its own) gives the final band; every key inside it -> video index ----
  const float4 qm = p.qmeta[q];
  const float tau_pub = tau_fetch(p.tau_g + q);
  const unsigned pre_bits = __float_as_uint(__fadd_ru(tau_pub, qm.w));
  float tau_fin = tau_pub;
  int total = 0;
  for (int part = 0; part < p.n_parts; ++part) total += min(p.cand_cnt[q * p.n_parts + part], SL_CAP);
  const bool one_batch = total <= RF_BATCH;
  // FAST PATH (the usual case: ~k + band candidates): every ordering step below is a rank-by-counting pass over shared
  // memory - O(n^2 / 256) compares per thread, two block barriers - instead of a 256-thread bitonic sort with one barrier
  // per stage (36 - 55 stages each for the key sort, the video sort and the final top-k: most of this kernel's time)
  const bool fast = total <= p.fast_max;
  __shared__ unsigned long long s_kth;
  if (fast) {
    if (tid == 0) s_kth = ~0ull;
    for (int part = 0; part < p.n_parts; ++part) {
      const int64_t li = q * p.n_parts + part;
      const int n = min(p.cand_cnt[li], SL_CAP);
      for (int i = tid; i < n; i += RF_THREADS) {
        const unsigned long long key = p.cand[li * SL_CAP + i];
        if ((unsigned)(key >> 32) <= pre_bits) keys[VFR_TOPK_MAX + atomicAdd(&s_cnt, 1)] = key;
      }
    }
    __syncthreads();
    const int g = s_cnt;
    for (int i = tid; i < g; i += RF_THREADS) {                 // the k-th smallest key (keys are unique: clip ids are)
      const unsigned long long key = keys[VFR_TOPK_MAX + i];
      int r = 0;
      for (int j = 0; j < g; ++j) r += (keys[VFR_TOPK_MAX + j] < key) ? 1 : 0;
      if (r == p.k - 1) s_kth = key;
    }
    __syncthreads();
  } else
  {
    int filled = 0;
    for (int part = 0; part < p.n_parts; ++part) {
      const int64_t li = q * p.n_parts + part;
      const int n = min(p.cand_cnt[li], SL_CAP);
      if (filled + n > RF_BATCH) {
        __syncthreads();
        rf_merge_batch(keys, s_cnt, p.k, &s_cnt, &s_tau);
        filled = 0;
      }
      for (int i = tid; i < n; i += RF_THREADS) {
        const unsigned long long key = p.cand[li * SL_CAP + i];
        if ((unsigned)(key >> 32) <= pre_bits) keys[VFR_TOPK_MAX + atomicAdd(&s_cnt, 1)] = key;
      }
      filled += n;
    }
    __syncthreads();
  }
  const int gathered = s_cnt;                       // (one_batch: every key of the band is in keys[128, 128 + gathered))
  const int n_sorted = rf_pow2(VFR_TOPK_MAX + gathered, 256);
  {
    if (!fast) {
      for (int i = VFR_TOPK_MAX + gathered + tid; i < n_sorted; i += RF_THREADS) keys[i] = ~0ull;
      rf_block_sort(keys, n_sorted);
    }
    const unsigned long long kth_key = fast ? s_kth : keys[p.k - 1];
    if (kth_key != ~0ull) tau_fin = fminf(tau_pub, __uint_as_float((unsigned)(kth_key >> 32)));
    // The filter ran under a SAMPLED threshold that no certified one undercut: it only holds if the bank really
    // has k keys under it (all of them are in the lists then).  Otherwise the query is flagged for the exact engine.
    if (tid == 0 && tau_pub < tau_fetch(p.tau_cert + q) &&
        (kth_key == ~0ull || (unsigned)(kth_key >> 32) > __float_as_uint(tau_pub)) && p.flags[q] == 0)
      p.flags[q] = 4;
  }
  stamp(0);
  const unsigned keep_bits = __float_as_uint(__fadd_ru(tau_fin, qm.w));
  // no moment scoring above this can be among the k best: the k closest clips are moments themselves
  const float s_max = (tau_fin < CUDART_INF_F) ? __fmul_ru(__fsqrt_ru(__fadd_ru(tau_fin, qm.w)), 1.00001f) : CUDART_INF_F;
  // every video of the bank has n_max clips <=> n_videos * n_max == n_clips: then clip -> video is a division
  const bool uniform = (int64_t)p.n_videos * p.n_max == p.n_clips;
  if (fast) {
    for (int i = tid; i < gathered; i += RF_THREADS) {
      const unsigned long long key = keys[VFR_TOPK_MAX + i];
      int v = INT_MAX;
      if ((unsigned)(key >> 32) <= keep_bits) {
        const int clip = (int)(unsigned)key;
        if (uniform) v = clip / p.n_max;
        else {
          int lo = 0, hi = (int)p.n_videos;
          while (hi - lo > 1) {
            const int mid = (lo + hi) >> 1;
            if (__ldg(p.vid_off + mid) <= clip) lo = mid; else hi = mid;
          }
          v = lo;
        }
      }
      vids[i] = v;
    }
  } else if (one_batch) {
    // the sorted keys ARE the band: a prefix of keys[]
    for (int i = tid; i < VFR_TOPK_MAX + gathered; i += RF_THREADS) {
      const unsigned long long key = keys[i];
      if (key != ~0ull && (unsigned)(key >> 32) <= keep_bits) {
        const int clip = (int)(unsigned)key;
        int lo = 0, hi = (int)p.n_videos;            // largest v with vid_off[v] <= clip
        while (hi - lo > 1) {
          const int mid = (lo + hi) >> 1;
          if (__ldg(p.vid_off + mid) <= clip) lo = mid; else hi = mid;
        }
        const int pos = atomicAdd(&s_n, 1);
        if (pos < RF_MAXC) vids[pos] = lo;
      }
    }
  } else {
    for (int part = 0; part < p.n_parts; ++part) {
      const int64_t li = q * p.n_parts + part;
      const int n = min(p.cand_cnt[li], SL_CAP);
      for (int i = tid; i < n; i += RF_THREADS) {
        const unsigned long long key = p.cand[li * SL_CAP + i];
        if ((unsigned)(key >> 32) <= keep_bits) {
          const int clip = (int)(unsigned)key;
          int lo = 0, hi = (int)p.n_videos;
          while (hi - lo > 1) {
            const int mid = (lo + hi) >> 1;
            if (__ldg(p.vid_off + mid) <= clip) lo = mid; else hi = mid;
          }
          const int pos = atomicAdd(&s_n, 1);
          if (pos < RF_MAXC) vids[pos] = lo;
        }
      }
    }
  }
  __syncthreads();
  for (int i = tid; i < VFR_TOPK_MAX; i += RF_THREADS) keys[i] = ~0ull;
  if (tid == 0) { s_cnt = 0; s_tau = CUDART_INF_F; }
  __syncthreads();
  int n_unique = 0;
  if (fast) {
    // unique videos in ascending order by counting: a video's place = the number of DISTINCT smaller videos
    __shared__ int s_nu;
    __shared__ unsigned char s_first[RF_FAST];
    if (tid == 0) s_nu = 0;
    __syncthreads();
    for (int i = tid; i < gathered; i += RF_THREADS) {          // first occurrence of its video?
      const int v = vids[i];
      bool first = v != INT_MAX;
      for (int j = 0; j < i && first; ++j) first = vids[j] != v;
      s_first[i] = first ? 1 : 0;
    }
    __syncthreads();
    for (int i = tid; i < gathered; i += RF_THREADS) {
      if (!s_first[i]) continue;
      const int v = vids[i];
      int pos = 0;
      for (int j = 0; j < gathered; ++j) pos += (s_first[j] && vids[j] < v) ? 1 : 0;
      uvid[pos] = v;
      atomicAdd(&s_nu, 1);
    }
    __syncthreads();
    n_unique = s_nu;
  } else {
  if (s_n > RF_MAXC && tid == 0) p.flags[q] = 3;   // more candidates than the refine stage holds
  rf_block_sort(vids, rf_pow2(min(s_n, RF_MAXC), 32));

  // ---- 2. unique videos, ascending ----
  constexpr int PER = RF_MAXC / RF_THREADS;
  int flagsum = 0;
  bool isnew[PER];
#pragma unroll
  for (int j = 0; j < PER; ++j) {
    const int i = tid * PER + j;
    const int v = vids[i];
    isnew[j] = v != INT_MAX && (i == 0 || vids[i - 1] != v);
    flagsum += isnew[j] ? 1 : 0;
  }
  int incl = flagsum;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int y = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += y;
  }
  if (lane == 31) warp_tot[wid] = incl;
  __syncthreads();
  int base = incl - flagsum;
#pragma unroll
  for (int w = 0; w < RF_THREADS / 32; ++w) {
    if (w < wid) base += warp_tot[w];
    n_unique += warp_tot[w];
  }
#pragma unroll
  for (int j = 0; j < PER; ++j)
    if (isnew[j]) uvid[base++] = vids[tid * PER + j];
  __syncthreads();
  }   // !fast

  for (int i = tid; i < min(n_unique, RF_MAXC); i += RF_THREADS) {
    const int v = uvid[i];
    if (uniform) { u_c0[i] = v * p.n_max; u_n[i] = (unsigned char)p.n_max; }
    else {
      const int c0 = __ldg(p.vid_off + v);
      u_c0[i] = c0;
      u_n[i] = (unsigned char)(__ldg(p.vid_off + v + 1) - c0);
    }
  }
  __syncthreads();
  stamp(1);
  // ---- 3. chunks of videos: exact clip distances, exact moment means, streaming top-k ----
  const int mom_max = num_moments(p.n_max);
  const int vch = max(1, min(RF_DIST / p.n_max, RF_BATCH / mom_max));
  const float sq = s_sq;
  bool merged = false;
  for (int v0 = 0; v0 < n_unique; v0 += vch) {
    const int nv = min(vch, n_unique - v0);
    if (s_cnt + nv * mom_max > RF_BATCH) { rf_merge_batch(keys, s_cnt, p.k, &s_cnt, &s_tau); merged = true; }
    for (int idx = tid; idx < nv * p.n_max; idx += RF_THREADS) {
      const int vi = idx / p.n_max, c = idx - vi * p.n_max;
      const int c0 = u_c0[v0 + vi];
      const int n = u_n[v0 + vi];
      if (c < n)
        dist[idx] = p.bank_b16 ? rf_distance_b16(p.bank_b16 + (int64_t)(c0 + c) * p.dim, qrow, sq, p.dim)
                               : rf_distance(p.bank + (int64_t)(c0 + c) * p.dim, qrow, sq, p.dim);
    }
    __syncthreads();
    stamp(2);
    const float tau = fminf(s_tau, s_max);
    for (int idx = tid; idx < nv * p.n_max; idx += RF_THREADS) {
      const int vi = idx / p.n_max, s = idx - vi * p.n_max;
      const int n = u_n[v0 + vi];
      if (s < n) {
        float run = 0.f;
        for (int e = s; e < n; ++e) {
          run = __fadd_rn(run, dist[vi * p.n_max + e]);
          const float score = __fdiv_rn(run, (float)(e - s + 1));
          if (score <= tau) {
            const int pos = atomicAdd(&s_cnt, 1);
            keys[VFR_TOPK_MAX + pos] =
                ((unsigned long long)__float_as_uint(score) << 32) | (unsigned)(((v0 + vi) << 10) | moment_index(n, s, e));
          }
        }
      }
    }
    __syncthreads();
    stamp(3);
  }
  if (!merged && s_cnt <= p.final_max) {
    // the k best of the <= RF_FINAL surviving moments by counting (keys are unique: (score, video, moment)); slot r of
    // keys[0, 128) - initialised to "empty" above and untouched since - receives the key of rank r
    const int m = s_cnt;
    for (int i = tid; i < m; i += RF_THREADS) {
      const unsigned long long key = keys[VFR_TOPK_MAX + i];
      int r = 0;
      for (int j = 0; j < m; ++j) r += (keys[VFR_TOPK_MAX + j] < key) ? 1 : 0;
      if (r < p.k) keys[r] = key;
    }
    __syncthreads();
  } else {
    rf_merge_batch(keys, s_cnt, p.k, &s_cnt, &s_tau);
  }

  stamp(4);
  if (p.dbg && tid == 0) { atomicAdd(p.dbg + 6, (unsigned long long)n_unique); atomicAdd(p.dbg + 7, (unsigned long long)s_cnt); }
  // ---- 4. output: ascending (score, global moment id) ----
  float* out_s;
  int64_t* out_i;
  if (p.per == 0) {
    out_s = p.out_scores + q * p.k;
    out_i = p.out_ids + q * p.k;
  } else {
    unsigned char* blk = p.out_blocks + (q / p.per) * p.blk_bytes;
    const int64_t r = q % p.per;
    out_i = reinterpret_cast<int64_t*>(blk) + r * p.k;
    out_s = reinterpret_cast<float*>(blk + p.per * p.k * 8) + r * p.k;
    // (flags 3 / 4 of this query were written by thread 0 of this CTA above, 1 / 2 by earlier kernels)
    if (tid == 0) reinterpret_cast<int32_t*>(blk + p.per * p.k * 12)[r] = p.flags[q];
  }
  for (int i = tid; i < p.k; i += RF_THREADS) {
    const unsigned long long key = keys[i];
    const bool ok = key != ~0ull;
    float score = CUDART_INF_F;
    int64_t id = -1;
    if (ok) {
      const unsigned low = (unsigned)key;
      score = __uint_as_float((unsigned)(key >> 32));
      id = p.id_base + __ldg(p.mom_off + uvid[low >> 10]) + (int64_t)(low & 1023u);
    }
    out_s[i] = score;
    out_i[i] = id;
  }
}

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------
typedef CUresult (*SlEncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                               const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                               CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static SlEncodeFn sl_get_encode() {
  static SlEncodeFn fn = nullptr;
  if (!fn) {
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<SlEncodeFn>(ptr);
  }
  return fn;
}

static int sl_make_map(CUtensorMap* map, const void* base, uint64_t rows, uint32_t box_rows, int pitch) {
  SlEncodeFn enc = sl_get_encode();
  VFR_REQUIRE(enc, VFR_ERR_CUDA, "cuTensorMapEncodeTiled entry point not available");
  cuuint64_t gdim[2] = {(cuuint64_t)pitch, (cuuint64_t)rows};
  cuuint64_t gstride[1] = {(cuuint64_t)pitch * 2};
  cuuint32_t box[2] = {64, box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<void*>(base), gdim, gstride, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  VFR_REQUIRE(r == CUDA_SUCCESS, VFR_ERR_CUDA, "cuTensorMapEncodeTiled failed with CUresult %d", (int)r);
  return VFR_OK;
}

static int64_t sl_tiles(int64_t n_clips) { return (n_clips + SL_N - 1) / SL_N; }
// fp16 columns per packed row: 128 while D + 3 fits, then whole 64-column chunks
static int sl_pitch(int dim) { return (dim + 3 <= SL_ROW) ? SL_ROW : (dim + 3 + 63) / 64 * 64; }
static bool sl_dim_ok(int dim) { return dim >= 1 && dim + 3 <= SL_MAXROW; }
static int64_t sl_qrows(int64_t n_queries) { return (n_queries + SL_QPAD - 1) / SL_QPAD * SL_QPAD; }

// query tiles per CTA: two whenever that still fills the machine (halves the L2 -> SM bank traffic)
static int sl_rows(int64_t n_queries) {
  const char* env = getenv("VFR_SEL_R");
  if (env && (env[0] == '1' || env[0] == '2')) return env[0] - '0';
  return (n_queries > SL_M) ? 2 : 1;
}

// bank splits per query group (one CTA per SM): the fewest that keep >= 85 % of the SMs busy
static int sl_split(int64_t n_qgroups, int64_t n_tiles, int n_split) {
  if (n_split > 0) return (int)(n_split < n_tiles ? n_split : n_tiles);
  int sms = 148, dev = 0;
  if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  int best = 1;
  double best_util = 0.0;
  for (int ns = 1; ns <= 148 && ns <= n_tiles; ++ns) {
    const int64_t ctas = ns * n_qgroups;
    const int64_t waves = (ctas + sms - 1) / sms;
    const double util = (double)ctas / (double)(waves * sms);
    if (util >= 0.85) return ns;
    if (util > best_util) { best_util = util; best = ns; }
  }
  return best;
}

struct SlPlan {
  int R, CL, n_qgroups, ns, tiles_per_split, n_parts;   // n_qgroups: CTAs along the queries (even when CL = 2)
  int64_t n_tiles, qrows;
  int pitch;      // fp16 columns per packed row
  bool big;       // K-streaming kernel (pitch > 128)
  int nb;         // accumulator scheme of the plain R = 2 kernel (sl_filter_kernel's NB): 2, or an opt-in variant (VFR_SEL_NB)
  int lists;      // candidate lists per (query, bank split): 2 with R = 1 (one per tile parity)
};
// accumulator / epilogue scheme requested for the plain R = 2 kernel
static int sl_nb_env() {
  static const int nb = [] { const char* e = getenv("VFR_SEL_NB"); const int v = e ? atoi(e) : 8; return ((v >= 2 && v <= 5) || v == 8) ? v : 8; }();
  return nb;
}

static SlPlan sl_plan(int64_t n_queries, int64_t n_clips, int n_split, int dim = 100) {
  SlPlan pl;
  pl.pitch = sl_pitch(dim);
  pl.big = pl.pitch > SL_ROW;
  pl.R = pl.big ? 2 : sl_rows(n_queries);
  pl.n_tiles = sl_tiles(n_clips);
  pl.qrows = sl_qrows(n_queries);
  const int64_t qtiles = (n_queries + SL_M - 1) / SL_M;
  pl.n_qgroups = (int)((qtiles + pl.R - 1) / pl.R);
  {
    const char* env = getenv("VFR_SEL_CL");
    // measured on B200 (18 944 queries x 6 M clips): the filter is bound by the MMA -> epilogue -> MMA hand-off
    // chain, not by L2 / TMA, so the CTA-pair multicast is off by default (VFR_SEL_CL=2 turns it on)
    pl.CL = (env && env[0] == '2' && pl.n_qgroups >= 2 && !pl.big) ? 2 : 1;
    if (pl.CL == 2) pl.n_qgroups = (pl.n_qgroups + 1) / 2 * 2;     // an odd group gets an idle partner CTA
  }
  const int ns_req = sl_split(pl.n_qgroups, pl.n_tiles, n_split);
  pl.tiles_per_split = (int)((pl.n_tiles + ns_req - 1) / ns_req);
  pl.ns = (int)((pl.n_tiles + pl.tiles_per_split - 1) / pl.tiles_per_split);
  pl.nb = (pl.R == 2 && pl.CL == 1 && !pl.big) ? sl_nb_env() : 2;
  pl.lists = pl.R == 1 ? 2 : 1;
  pl.n_parts = pl.ns * pl.lists;
  return pl;
}

}  // namespace vfr

using namespace vfr;

extern "C" size_t vfr_sel_bank_bytes(int64_t n_clips, int dim) {
  if (n_clips <= 0 || !sl_dim_ok(dim)) return 0;
  return (size_t)sl_tiles(n_clips) * SL_N * sl_pitch(dim) * 2 + sizeof(SlBankMeta);
}

extern "C" int vfr_sel_bank_pack(const float* bank, int64_t n_clips, int dim, void* packed, vfr_stream_t stream) {
  VFR_REQUIRE(bank && packed, VFR_ERR_INVALID, "vfr_sel_bank_pack: null pointer");
  VFR_REQUIRE(n_clips > 0 && n_clips < (int64_t(1) << 31) - SL_N, VFR_ERR_UNSUPPORTED,
              "vfr_sel_bank_pack: n_clips=%lld out of range", (long long)n_clips);
  VFR_REQUIRE(sl_dim_ok(dim), VFR_ERR_UNSUPPORTED, "vfr_sel_bank_pack: dim=%d must be <= %d", dim, SL_MAXROW - 3);
  const int64_t rows = sl_tiles(n_clips) * SL_N;
  const int pitch = sl_pitch(dim);
  __half* out = reinterpret_cast<__half*>(packed);
  SlBankMeta* meta = reinterpret_cast<SlBankMeta*>(out + rows * pitch);
  cudaStream_t st = (cudaStream_t)stream;
  VFR_CUDA(cudaMemsetAsync(meta, 0, sizeof(SlBankMeta), st));
  const int64_t sum_blocks = n_clips < 148 * 8 ? n_clips : 148 * 8;
  sl_bank_colsum_kernel<<<(unsigned)sum_blocks, 256, 0, st>>>(bank, n_clips, dim, meta);
  int rc0 = check_launch("sl_bank_colsum_kernel");
  if (rc0) return rc0;
  sl_bank_center_kernel<<<1, 256, 0, st>>>(n_clips, dim, meta);
  rc0 = check_launch("sl_bank_center_kernel");
  if (rc0) return rc0;
  const int64_t stat_blocks = (n_clips + 7) / 8 < 148 * 16 ? (n_clips + 7) / 8 : 148 * 16;
  sl_bank_stats_kernel<<<(unsigned)stat_blocks, 256, 0, st>>>(bank, n_clips, dim, meta);
  int rc = check_launch("sl_bank_stats_kernel");
  if (rc) return rc;
  sl_bank_scales_kernel<<<1, 1, 0, st>>>(meta);
  rc = check_launch("sl_bank_scales_kernel");
  if (rc) return rc;
  sl_bank_pack_kernel<<<(unsigned)((rows + 7) / 8), 256, 0, st>>>(bank, n_clips, rows, dim, pitch, meta, out);
  return check_launch("sl_bank_pack_kernel");
}

extern "C" size_t vfr_sel_query_bytes(int64_t n_queries, int dim) {
  if (n_queries <= 0 || !sl_dim_ok(dim)) return 0;
  const size_t rows = (size_t)sl_qrows(n_queries);
  return rows * sl_pitch(dim) * 2 + rows * sizeof(float4) + rows * sizeof(int32_t);
}

extern "C" int vfr_sel_query_pack(const float* queries, int64_t n_queries, int dim, const void* bank_packed,
                                  int64_t n_clips, void* packed, vfr_stream_t stream) {
  VFR_REQUIRE(queries && bank_packed && packed, VFR_ERR_INVALID, "vfr_sel_query_pack: null pointer");
  VFR_REQUIRE(n_queries > 0 && n_clips > 0 && sl_dim_ok(dim), VFR_ERR_UNSUPPORTED, "vfr_sel_query_pack: bad shape");
  const int64_t rows = sl_qrows(n_queries);
  const int pitch = sl_pitch(dim);
  const SlBankMeta* meta =
      reinterpret_cast<const SlBankMeta*>(reinterpret_cast<const __half*>(bank_packed) + sl_tiles(n_clips) * SL_N * pitch);
  __half* out = reinterpret_cast<__half*>(packed);
  float4* qmeta = reinterpret_cast<float4*>(out + rows * pitch);
  int32_t* flags = reinterpret_cast<int32_t*>(qmeta + rows);
  sl_query_pack_kernel<<<(unsigned)((rows + 7) / 8), 256, 0, (cudaStream_t)stream>>>(queries, n_queries, rows, dim, pitch, meta,
                                                                                     out, qmeta, flags);
  return check_launch("sl_query_pack_kernel");
}

extern "C" size_t vfr_sel_topk_bytes(int64_t n_queries, int64_t n_clips, int n_split) {
  if (n_queries <= 0 || n_clips <= 0) return 0;
  // the plan (query tiles per CTA, bank splits) depends on the batch size: size for the worst batch <= n_queries
  size_t worst = 0;
  const int64_t qtiles = (n_queries + SL_M - 1) / SL_M;
  for (int64_t qt = 1; qt <= qtiles; ++qt) {
    // (the K-streaming kernel always serves two query tiles per CTA: lists per query <= the plan at D <= 125)
    const SlPlan pl = sl_plan(qt == qtiles ? n_queries : qt * SL_M, n_clips, n_split);
    const size_t qpad = (size_t)pl.qrows;
    const size_t need = qpad * pl.n_parts * SL_CAP * sizeof(unsigned long long) + qpad * pl.n_parts * sizeof(int32_t) +
                        2 * qpad * sizeof(unsigned) + qpad * pl.n_parts * sizeof(float) + qpad * pl.n_parts * SL_J * sizeof(float) +
                        8 * sizeof(unsigned long long);
    if (need > worst) worst = need;
  }
  return worst;
}

namespace vfr {

// everything the two stages share: the launch plan, the workspace carve-up, the per-query metadata
static int sl_setup(SlPlan& pl, SlParams& p, void* query_packed, int64_t n_queries, int64_t n_clips, int dim, int k,
                    void* workspace, int n_split) {
  VFR_REQUIRE(n_clips > 0 && n_queries > 0, VFR_ERR_INVALID, "vfr_sel: empty bank or batch");
  VFR_REQUIRE(n_clips < (int64_t(1) << 31) - SL_N, VFR_ERR_UNSUPPORTED, "vfr_sel: bank shard too large");
  VFR_REQUIRE(sl_dim_ok(dim), VFR_ERR_UNSUPPORTED, "vfr_sel: dim=%d must be <= %d", dim, SL_MAXROW - 3);
  VFR_REQUIRE(k >= 1 && k <= VFR_TOPK_MAX, VFR_ERR_UNSUPPORTED, "k=%d not in [1,%d]", k, VFR_TOPK_MAX);
  pl = sl_plan(n_queries, n_clips, n_split, dim);
  __half* qp = reinterpret_cast<__half*>(query_packed);
  p = SlParams{};
  p.qmeta = reinterpret_cast<const float4*>(qp + pl.qrows * pl.pitch);
  p.qpack = reinterpret_cast<const uint32_t*>(query_packed);
  p.flags = reinterpret_cast<int32_t*>(const_cast<float4*>(p.qmeta) + pl.qrows);
  p.n_clips = n_clips;
  p.n_queries = n_queries;
  p.n_qgroups = pl.n_qgroups;
  p.n_tiles = (int)pl.n_tiles;
  p.tiles_per_split = pl.tiles_per_split;
  p.ksteps = (dim + 3 + 15) / 16;
  p.n_chunks = pl.pitch / 64;
  p.k = k;
  p.n_parts = pl.n_parts;
  { const char* wm = getenv("VFR_SEL_WAIT"); p.wait_mode = wm ? atoi(wm) : 0; }
  { const char* pf = getenv("VFR_SEL_PF"); p.prefetch = pf ? std::max(0, atoi(pf)) : 0; }
  { const char* pd = getenv("VFR_SEL_PDBG"); p.prod_dbg = pd ? atoi(pd) : 0; }
  { const char* ro = getenv("VFR_SEL_ROT"); p.rot_step = ro ? std::max(0, atoi(ro)) : 0; }
  { const char* dg = getenv("VFR_SEL_DBG"); p.dbg = dg ? reinterpret_cast<long long*>(strtoull(dg, nullptr, 0)) : nullptr; }
  const size_t qpad = (size_t)pl.qrows;
  p.cand = reinterpret_cast<unsigned long long*>(workspace);
  p.cand_cnt = reinterpret_cast<int32_t*>(p.cand + qpad * (size_t)pl.n_parts * SL_CAP);
  p.tau_g = reinterpret_cast<unsigned*>(p.cand_cnt + qpad * (size_t)pl.n_parts);
  p.tau_cert = p.tau_g + qpad;
  p.tau_part = reinterpret_cast<float*>(p.tau_cert + qpad);
  p.samp = nullptr;   // (the export buffer lives behind tau_part: sl_samp_buffer)
  p.stats = reinterpret_cast<unsigned long long*>(p.tau_part + qpad * (size_t)pl.n_parts * (1 + SL_J));
  p.tile_stride = 1;
  // measured (37 888 queries, k = 100; whole 6 M-clip bank / one of 8 shards): first compaction after 3k keys, later
  // ones only when the list is full: 52.3 / 9.9 ms; every k keys: 52.0 / 12.1 ms - a compaction stalls the CTA's
  // pipeline for ~10 k cycles whatever the bank size, and a shard's scan is short
  p.cap_trig = std::min(SL_CAP_HI, std::max(64, 3 * k));
  p.cap_step = SL_CAP_HI;
  { const char* e = getenv("VFR_SEL_TRIG"); if (e) p.cap_trig = std::min(SL_CAP_HI, std::max(k + 1, atoi(e))); }
  { const char* e = getenv("VFR_SEL_STEP"); if (e) p.cap_step = std::min(SL_CAP_HI, std::max(1, atoi(e))); }
  return VFR_OK;
}

// Sample pass planning.  Every list draws `tiles` bank tiles (a strided sample of the whole bank, n_eff = 256 * tiles
// clips) and takes the j-th smallest of its 64-column minima as the starting threshold.  The threshold is wrong only
// if fewer than k of the bank's clips lie under it, i.e. if at least j of the bank's k - 1 closest clips fell into
// the sample: P <= P(Poisson(k * n_eff / n_clips) >= j) for a bank in no particular order.  j is the smallest rank
// that makes this < 1e-10 per list; a wrong threshold is detected by stage 2 (flag 4) and costs a rerun, never a
// wrong result.  tiles = 0: no sample (bank too small to gain from it).
struct SlSample { int tiles, stride, j; };
static SlSample sl_sample_plan(const SlPlan& pl, int64_t n_clips, int k, bool need_rank = true) {
  SlSample sp{0, 1, 0};
  { const char* e = getenv("VFR_SEL_SAMPLE"); if (e && e[0] == '0') return sp; }
  const int lists = pl.ns * pl.lists;                         // lists per query, each samples on its own
  const int64_t per_list = pl.n_tiles / lists;
  int tiles = 64;                                              // more sample for larger k: 8 k tiles, 64 ... 512
  while (tiles < 512 && tiles < 8 * k) tiles <<= 1;
  { const char* e = getenv("VFR_SEL_SAMPLE_TILES"); if (e) tiles = std::max(8, atoi(e)); }
  while (tiles >= 8) {
    const int64_t stride = (pl.n_tiles - 1) / ((int64_t)tiles * lists);   // never reaches the (padded) last tile
    // (the sample costs tiles / per_list of a scan: <= 1/32; the tighter the starting threshold, the fewer keys pass)
    if ((per_list >= 32 * tiles || tiles == 8) && per_list >= 128 && stride >= 2) {
      if (!need_rank) { sp.tiles = tiles; sp.stride = (int)stride; sp.j = 0; return sp; }
      const double x = (double)k * 256.0 * tiles / (double)((pl.n_tiles - 1) * SL_N);
      double term = exp(-x), cdf = 0.0;                      // P(Poisson(x) >= j) = 1 - sum_{i<j} e^-x x^i / i!
      for (int j = 1; j <= SL_J; ++j) {
        cdf += term;
        term *= x / j;
        if (1.0 - cdf < 1e-10) { sp.tiles = tiles; sp.stride = (int)stride; sp.j = j; return sp; }
      }
    }
    tiles >>= 1;
  }
  return sp;
}

template <int MODE>
static int sl_launch_filter(const SlPlan& pl, const SlParams& p, const CUtensorMap& ma, const CUtensorMap& mb, cudaStream_t st) {
  const int CL = (MODE == 1) ? 1 : pl.CL;
  const unsigned grid = (unsigned)(pl.n_qgroups * pl.ns);
  void (*kern)(CUtensorMap, CUtensorMap, SlParams) = nullptr;
  uint32_t smem_bytes = 0;
  unsigned threads = SL_THREADS;
  // accumulators per CTA: two of 256 columns (default); VFR_SEL_NB=4: four of 128 (measured slower, see the kernel)
  const int nb = pl.nb;
  if (pl.big) { kern = sl_filter_kernel<2, 1, MODE, true>; smem_bytes = SL_BIG_SMEM; }
  else if (pl.R == 2 && CL == 2) { kern = sl_filter_kernel<2, 2, 0>; smem_bytes = SlCfg<2>::SMEM; }
  else if (pl.R == 2 && nb == 3) { kern = sl_filter_kernel<2, 1, MODE, false, 3>; smem_bytes = SL_ATM_SMEM; }
  else if (pl.R == 2 && nb == 8) { kern = sl_filter_kernel<2, 1, MODE, false, 8>; smem_bytes = SlCfg<2>::SMEM; threads = SL_THREADS2; }
  else if (pl.R == 2 && nb == 5) { kern = sl_filter_kernel<2, 1, MODE, false, 5>; smem_bytes = SlCfg<2>::SMEM; threads = SL_THREADS2; }
  else if (pl.R == 2) { kern = (nb == 4) ? sl_filter_kernel<2, 1, MODE, false, 4> : sl_filter_kernel<2, 1, MODE>; smem_bytes = SlCfg<2>::SMEM; }
  else if (CL == 2) { kern = sl_filter_kernel<1, 2, 0>; smem_bytes = SlCfg<1>::SMEM; }
  else { kern = (nb == 4) ? sl_filter_kernel<1, 1, MODE, false, 4> : sl_filter_kernel<1, 1, MODE>; smem_bytes = SlCfg<1>::SMEM; }
  VFR_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes));
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(grid, 1, 1);
  cfg.blockDim = dim3(threads, 1, 1);
  cfg.dynamicSmemBytes = smem_bytes;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = (unsigned)CL;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  VFR_CUDA(cudaLaunchKernelEx(&cfg, kern, ma, mb, p));
  return check_launch(MODE == 1 ? "sl_filter_kernel (sample pass)" : "sl_filter_kernel");
}

static float* sl_samp_buffer(const SlPlan& pl, const SlParams& p) { return p.tau_part + (size_t)pl.qrows * pl.n_parts; }

static int sl_launch_sample(const SlPlan& pl, const SlParams& p, const SlSample& sp, float* samp, const CUtensorMap& ma,
                            const CUtensorMap& mb, cudaStream_t st) {
  SlParams ps = p;
  ps.tile_lo = 0;
  ps.tiles_per_split = sp.tiles * pl.lists;    // R = 1: the two sets of a CTA take alternate sample tiles
  ps.n_tiles = ps.tiles_per_split * pl.ns;
  ps.tile_stride = sp.stride;
  ps.sample_j = sp.j;
  ps.samp = samp;
  return sl_launch_filter<1>(pl, ps, ma, mb, st);
}

// stage 1 over the bank tiles [tile_lo, tile_hi)
static int sl_run_filter(const SlPlan& pl, SlParams p, const void* bank_packed, void* query_packed, int64_t tile_lo,
                         int64_t tile_hi, int resume, cudaStream_t st) {
  if (tile_hi < 0 || tile_hi > pl.n_tiles) tile_hi = pl.n_tiles;
  VFR_REQUIRE(tile_lo >= 0 && tile_lo <= tile_hi, VFR_ERR_INVALID, "vfr_sel_filter: bad tile range");
  CUtensorMap ma, mb;
  int rc = sl_make_map(&ma, query_packed, (uint64_t)pl.qrows, SL_M, pl.pitch);
  if (rc) return rc;
  rc = sl_make_map(&mb, bank_packed, (uint64_t)(pl.n_tiles * SL_N), SL_N, pl.pitch);
  if (rc) return rc;
  if (!resume) {
    const size_t qpad = (size_t)pl.qrows;
    rc = launch_fill_u32(p.tau_g, 0x7f800000u, qpad * (2 + (size_t)pl.n_parts), st);   // tau_g, tau_cert, tau_part: +inf
    if (rc) return rc;
    VFR_CUDA(cudaMemsetAsync(p.stats, 0, 8 * sizeof(unsigned long long), st));
    if (tile_lo == tile_hi) VFR_CUDA(cudaMemsetAsync(p.cand_cnt, 0, qpad * (size_t)pl.n_parts * sizeof(int32_t), st));
    // starting thresholds from a strided sample of the WHOLE bank (whatever slice this call scans)
    const SlSample sp = sl_sample_plan(pl, p.n_clips, p.k);
    if (sp.tiles > 0) {
      rc = sl_launch_sample(pl, p, sp, nullptr, ma, mb, st);
      if (rc) return rc;
    }
  }
  if (tile_lo == tile_hi) return VFR_OK;
  p.tile_lo = (int)tile_lo;
  p.n_tiles = (int)tile_hi;
  p.tiles_per_split = (int)((tile_hi - tile_lo + pl.ns - 1) / pl.ns);
  p.resume = resume;
  return sl_launch_filter<0>(pl, p, ma, mb, st);
}

static int sl_run_refine(const SlParams& p, const float* bank, const int32_t* vid_off, const int64_t* mom_off,
                         int64_t n_videos, int n_max, int dim, const float* queries, int64_t n_queries, int k,
                         int64_t id_base, float* out_scores, int64_t* out_ids, cudaStream_t st, int64_t per = 0,
                         void* out_blocks = nullptr, bool bank_is_b16 = false, int64_t q_begin = 0, int64_t q_count = -1) {
  VFR_REQUIRE(bank && vid_off && mom_off && queries && ((out_scores && out_ids) || (per > 0 && out_blocks)), VFR_ERR_INVALID,
              "vfr_sel_refine: null pointer");
  VFR_REQUIRE(per == 0 || (per * (3 * (int64_t)k + 1)) % 2 == 0, VFR_ERR_INVALID, "vfr_sel_refine_blocks: per * (3k + 1) must be even");
  VFR_REQUIRE(n_videos > 0 && n_videos < (int64_t(1) << 31) - 1, VFR_ERR_UNSUPPORTED, "vfr_sel_refine: n_videos");
  VFR_REQUIRE(n_max >= 1 && n_max <= VFR_MAX_SEG, VFR_ERR_UNSUPPORTED, "vfr_sel_refine: n_max=%d", n_max);
  RfParams r{};
  r.bank = bank_is_b16 ? nullptr : bank;
  r.bank_b16 = bank_is_b16 ? reinterpret_cast<const __nv_bfloat16*>(bank) : nullptr;
  r.queries = queries;
  r.vid_off = vid_off;
  r.mom_off = mom_off;
  r.n_videos = n_videos;
  r.n_clips = p.n_clips;
  r.n_max = n_max;
  r.dim = dim;
  r.qmeta = p.qmeta;
  r.cand = p.cand;
  r.cand_cnt = p.cand_cnt;
  r.n_parts = p.n_parts;
  r.tau_g = p.tau_g;
  r.tau_cert = p.tau_cert;
  r.flags = p.flags;
  r.k = k;
  r.id_base = id_base;
  r.out_scores = out_scores;
  r.out_ids = out_ids;
  r.per = per;
  r.blk_bytes = per * ((int64_t)k * 12 + 4);
  r.out_blocks = reinterpret_cast<unsigned char*>(out_blocks);
  r.fast_max = RF_FAST;
  r.final_max = RF_FINAL;
  { const char* e = getenv("VFR_RF_DBG"); r.dbg = e ? reinterpret_cast<unsigned long long*>(strtoull(e, nullptr, 0)) : nullptr; }
  { const char* e = getenv("VFR_RF_FAST"); if (e) { r.fast_max = std::min(RF_FAST, std::max(0, atoi(e))); if (r.fast_max == 0) r.final_max = -1; } }
  if (q_count < 0) q_count = n_queries - q_begin;
  VFR_REQUIRE(q_begin >= 0 && q_count >= 0 && q_begin + q_count <= n_queries, VFR_ERR_INVALID, "vfr_sel_refine: bad query range");
  if (q_count == 0) return VFR_OK;
  r.q0 = q_begin;
  sl_refine_kernel<<<(unsigned)q_count, RF_THREADS, 0, st>>>(r);
  return check_launch("sl_refine_kernel");
}

// bound[q] = tau_cert[q] + E_q (certified thresholds only - a sampled one is a guess until stage 2) : an upper bound of the EXACT k-th smallest squared clip distance of this shard
__global__ void sl_bound_get_kernel(const unsigned* __restrict__ tau_cert, const float4* __restrict__ qmeta, int64_t n,
                                    float* __restrict__ bound) {
  const int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (q < n) bound[q] = __fadd_ru(__uint_as_float(tau_cert[q]), 0.5f * qmeta[q].w);
}
// a bound that holds for the k-th smallest of a LARGER bank (all shards) tightens this shard's threshold:
// a clip can only matter if its exact d^2 <= bound, i.e. its d2~ <= bound + E = (bound - E) + 2E
__global__ void sl_bound_put_kernel(unsigned* __restrict__ tau_g, unsigned* __restrict__ tau_cert,
                                    const float4* __restrict__ qmeta, int64_t n, const float* __restrict__ bound) {
  const int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (q < n) {
    const float t = fmaxf(__fsub_ru(bound[q], 0.5f * qmeta[q].w), 0.f);
    if (t < __uint_as_float(tau_g[q])) tau_g[q] = __float_as_uint(t);
    if (t < __uint_as_float(tau_cert[q])) tau_cert[q] = __float_as_uint(t);
  }
}

// count[q] = number of retained keys with d2~ <= bound[q] - E_q, i.e. of clips whose EXACT squared distance is
// certainly <= bound[q].  One warp per query.
__global__ void sl_count_under_kernel(const unsigned long long* __restrict__ cand, const int32_t* __restrict__ cand_cnt,
                                      int n_parts, const float4* __restrict__ qmeta, int64_t n,
                                      const float* __restrict__ bound, int32_t* __restrict__ count) {
  const int lane = threadIdx.x & 31;
  const int64_t q = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (q >= n) return;
  const float t = __fsub_rd(bound[q], 0.5f * qmeta[q].w);
  int c = 0;
  if (t >= 0.f) {
    const unsigned tb = __float_as_uint(t);
    for (int part = 0; part < n_parts; ++part) {
      const int64_t li = q * n_parts + part;
      const int m = min(cand_cnt[li], SL_CAP);
      for (int i = lane; i < m; i += 32) c += ((unsigned)(cand[li * SL_CAP + i] >> 32) <= tb) ? 1 : 0;
    }
  }
  c = warp_sum_int(c);
  if (lane == 0) count[q] = c;
}

// out[0] = sum of the list lengths, out[1] = longest list, out[2] = flagged queries, out[3] = lists,
// out[4] = warp-level compaction events, out[5] = lists compacted  (out zeroed by the caller)
__global__ void sl_stats_kernel(const int32_t* __restrict__ cand_cnt, int n_parts, const int32_t* __restrict__ flags, int64_t n,
                                const unsigned long long* __restrict__ stats, unsigned long long* __restrict__ out) {
  const int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  unsigned long long tot = 0, mx = 0, fl = 0;
  if (q < n) {
    for (int part = 0; part < n_parts; ++part) {
      const unsigned long long c = (unsigned long long)min(cand_cnt[q * n_parts + part], SL_CAP);
      tot += c;
      mx = c > mx ? c : mx;
    }
    fl = flags[q] != 0 ? 1 : 0;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    tot += __shfl_xor_sync(0xffffffffu, tot, o);
    fl += __shfl_xor_sync(0xffffffffu, fl, o);
    const unsigned long long m2 = __shfl_xor_sync(0xffffffffu, mx, o);
    mx = m2 > mx ? m2 : mx;
  }
  if ((threadIdx.x & 31) == 0) {
    atomicAdd(out, tot);
    atomicMax(out + 1, mx);
    atomicAdd(out + 2, fl);
  }
  if (q == 0) { out[3] = (unsigned long long)n * n_parts; out[4] = stats[0]; out[5] = stats[1]; }
}

// ---- the shard-level glue of the pooled-sample protocol, one kernel each (was a chain of torch ops) ----------------
// levels[l][q] = the ranks[l]-th smallest (1-based) of the n_src * width pooled sample values of query q
// (pooled fp32 [n_src, Q, width], +inf padded).  One warp per query; rank by counting (ties by position).
constexpr int PL_MAX = 1024;
__global__ void __launch_bounds__(256) sl_pool_levels_kernel(const float* __restrict__ pooled, int n_src, int64_t n_queries,
                                                             int width, int4 ranks, int n_levels, int sorted_runs,
                                                             float* __restrict__ levels) {
  __shared__ float vals[8][PL_MAX];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int64_t q = (int64_t)blockIdx.x * 8 + w;
  if (q >= n_queries) return;
  const int n = n_src * width;
  for (int i = lane; i < n; i += 32) {
    const int src = i / width, j = i - src * width;
    vals[w][i] = pooled[((int64_t)src * n_queries + q) * width + j];
  }
  __syncwarp();
  const int rk[4] = {ranks.x, ranks.y, ranks.z, ranks.w};
  // every run of SL_J pooled values is ascending (the sample pass exports them sorted), so a value at position p of its run
  // has p smaller-or-equal predecessors: only the first max(ranks) positions of a run can hold one of the wanted ranks,
  // and only they need to be counted (n^2 -> (P max_rank)^2 compares: 0.54 -> 0.1 ms for 37 888 queries x 8 shards)
  const int consider = sorted_runs ? min(SL_J, max(max(rk[0], rk[1]), max(rk[2], rk[3]))) : SL_J;
  for (int i = lane; i < n; i += 32) {
    if ((i % SL_J) >= consider) continue;
    const float v = vals[w][i];
    int before = 0;
    if (sorted_runs) {
      // rank = own position + binary searches in the other runs (ties by position: runs in front count their equal
      // values, runs behind do not) - 8 x 5 probes instead of 8 x 32 compares per value
      const int own = i / SL_J;
      before = i - own * SL_J;
      for (int r = 0, j0 = 0; j0 < n; ++r, j0 += SL_J) {
        if (r == own) continue;
        const int len = min(consider, n - j0);
        int lo = 0, hi = len;                    // first position whose value is > v (r < own) or >= v (r > own)
        while (lo < hi) {
          const int mid = (lo + hi) >> 1;
          const float u = vals[w][j0 + mid];
          const bool left = (r < own) ? (u <= v) : (u < v);
          lo = left ? mid + 1 : lo;
          hi = left ? hi : mid;
        }
        before += lo;
      }
    } else {
      for (int j = 0; j < n; ++j) {
        const float u = vals[w][j];
        before += (u < v || (u == v && j < i)) ? 1 : 0;
      }
    }
#pragma unroll
    for (int l = 0; l < 4; ++l)
      if (l < n_levels && before == rk[l] - 1) levels[(int64_t)l * n_queries + q] = v;
  }
}

// The same for <= 8 ascending runs of SL_J values per query (what the sample passes of <= 8 shards export): an 8-lane
// group merges the runs head by head - lane s walks run s, a 3-step shuffle minimum on (value, run) picks the next
// value in the order the counting kernel ranks by (ties by position) - and stops at the largest wanted rank.  Four
// queries per warp; ~200 instructions per query instead of ~3 400 (0.28 -> 0.03 ms for 37 888 queries x 8 shards).
__global__ void __launch_bounds__(256) sl_pool_levels_merge_kernel(const float* __restrict__ pooled, int n_src, int64_t n_queries,
                                                                   int width, int4 ranks, int n_levels, float* __restrict__ levels) {
  __shared__ float heads[32][8][SL_J + 1];
  const int lane = threadIdx.x & 31, sub = lane & 7, grp = (threadIdx.x >> 3);      // grp: query slot of the block (0..31)
  const int64_t q = (int64_t)blockIdx.x * 32 + grp;
  const bool qv = q < n_queries;
  const int runs_per_src = width / SL_J, n_runs = n_src * runs_per_src;
  const bool mine = qv && sub < n_runs;
  if (mine) {
    const int src = sub / runs_per_src, j0 = (sub - src * runs_per_src) * SL_J;
    const float4* run = reinterpret_cast<const float4*>(pooled + ((int64_t)src * n_queries + q) * width + j0);
#pragma unroll
    for (int i = 0; i < SL_J / 4; ++i) {
      const float4 v = run[i];
      heads[grp][sub][4 * i] = v.x; heads[grp][sub][4 * i + 1] = v.y; heads[grp][sub][4 * i + 2] = v.z; heads[grp][sub][4 * i + 3] = v.w;
    }
  }
  __syncwarp();
  const int rk[4] = {ranks.x, ranks.y, ranks.z, ranks.w};
  int max_rank = 0;
  for (int l = 0; l < n_levels; ++l) max_rank = max(max_rank, rk[l]);
  int pos = 0;
  float head = mine ? heads[grp][sub][0] : CUDART_INF_F;
  for (int r = 1; r <= max_rank; ++r) {
    float v = head;
    int who = sub;
#pragma unroll
    for (int off = 4; off >= 1; off >>= 1) {
      const float ov = __shfl_xor_sync(0xffffffffu, v, off);
      const int ow = __shfl_xor_sync(0xffffffffu, who, off);
      if (ov < v || (ov == v && ow < who)) { v = ov; who = ow; }
    }
    if (sub == 0 && qv) {
#pragma unroll
      for (int l = 0; l < 4; ++l)
        if (l < n_levels && r == rk[l]) levels[(int64_t)l * n_queries + q] = v;
    }
    if (sub == who) {
      ++pos;
      head = (mine && pos < SL_J) ? heads[grp][sub][pos] : CUDART_INF_F;
    }
  }
}

// count[l][q] = number of retained keys with d2~ <= levels[l][q] - E_q (clips CERTAINLY within that bound)
__global__ void sl_count_levels_kernel(const unsigned long long* __restrict__ cand, const int32_t* __restrict__ cand_cnt,
                                       int n_parts, const float4* __restrict__ qmeta, int64_t n, const float* __restrict__ levels,
                                       int n_levels, int32_t* __restrict__ count) {
  const int lane = threadIdx.x & 31;
  const int64_t q = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (q >= n) return;
  const float e = 0.5f * qmeta[q].w;
  unsigned tb[4];
  bool ok[4];
#pragma unroll
  for (int l = 0; l < 4; ++l) {
    const float t = (l < n_levels) ? __fsub_rd(levels[(int64_t)l * n + q], e) : -1.f;
    ok[l] = t >= 0.f;
    tb[l] = __float_as_uint(fmaxf(t, 0.f));
  }
  int c[4] = {0, 0, 0, 0};
  for (int part = 0; part < n_parts; ++part) {
    const int64_t li = q * n_parts + part;
    const int m = min(cand_cnt[li], SL_CAP);
    for (int i = lane; i < m; i += 32) {
      const unsigned key = (unsigned)(cand[li * SL_CAP + i] >> 32);
#pragma unroll
      for (int l = 0; l < 4; ++l) c[l] += (ok[l] && key <= tb[l]) ? 1 : 0;
    }
  }
#pragma unroll
  for (int l = 0; l < 4; ++l) {
    const int tot = warp_sum_int(c[l]);
    if (lane == 0 && l < n_levels) count[(int64_t)l * n + q] = tot;
  }
}

// after the all-reduce of the counts: the tightest level that still holds k clips over ALL shards is a certified bound
// of the global k-th distance (levels are ordered loosest first, so ok is monotone); a query whose loosest level fails
// was promised k clips the bank does not have -> flag 4.  The bound is put into tau_g / tau_cert like sl_bound_put.
__global__ void sl_pick_put_kernel(const float* __restrict__ levels, const int32_t* __restrict__ count, int n_levels, int k,
                                   unsigned* __restrict__ tau_g, unsigned* __restrict__ tau_cert, const float4* __restrict__ qmeta,
                                   int32_t* __restrict__ flags, int64_t n) {
  const int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= n) return;
  int n_ok = 0;
  for (int l = 0; l < n_levels; ++l) n_ok += (count[(int64_t)l * n + q] >= k) ? 1 : 0;
  if (count[q] < k && flags[q] == 0) flags[q] = 4;
  const float bound = levels[(int64_t)(max(n_ok, 1) - 1) * n + q];
  const float t = fmaxf(__fsub_ru(bound, 0.5f * qmeta[q].w), 0.f);
  if (t < __uint_as_float(tau_g[q])) tau_g[q] = __float_as_uint(t);
  if (t < __uint_as_float(tau_cert[q])) tau_cert[q] = __float_as_uint(t);
}

}  // namespace vfr

extern "C" int vfr_sel_pool_levels(const float* pooled, int n_src, int64_t n_queries, int width, const int32_t* ranks,
                                   int n_levels, int sorted_runs, float* levels, vfr_stream_t stream) {
  VFR_REQUIRE(pooled && ranks && levels, VFR_ERR_INVALID, "vfr_sel_pool_levels: null pointer");
  VFR_REQUIRE(n_src >= 1 && width >= 1 && n_src * (int64_t)width <= PL_MAX && n_levels >= 1 && n_levels <= 4 && n_queries > 0,
              VFR_ERR_UNSUPPORTED, "vfr_sel_pool_levels: %d x %d pooled values, %d levels", n_src, width, n_levels);
  int4 rk = make_int4(0, 0, 0, 0);
  int* rp = &rk.x;
  for (int l = 0; l < n_levels; ++l) {
    VFR_REQUIRE(ranks[l] >= 1 && ranks[l] <= n_src * width, VFR_ERR_INVALID, "vfr_sel_pool_levels: rank %d", ranks[l]);
    rp[l] = ranks[l];
  }
  // (width is a multiple of SL_J = 32 when the values come from vfr_sel_sample: runs of 32 ascending values)
  if (sorted_runs && width % SL_J == 0 && n_src * (width / SL_J) <= 8 && (reinterpret_cast<uintptr_t>(pooled) & 15) == 0) {
    sl_pool_levels_merge_kernel<<<(unsigned)((n_queries + 31) / 32), 256, 0, (cudaStream_t)stream>>>(pooled, n_src, n_queries, width,
                                                                                                   rk, n_levels, levels);
    return check_launch("sl_pool_levels_merge_kernel");
  }
  sl_pool_levels_kernel<<<(unsigned)((n_queries + 7) / 8), 256, 0, (cudaStream_t)stream>>>(pooled, n_src, n_queries, width, rk,
                                                                                        n_levels, sorted_runs && width % SL_J == 0, levels);
  return check_launch("sl_pool_levels_kernel");
}

extern "C" int vfr_sel_count_levels(void* query_packed, int64_t n_queries, int64_t n_clips, int dim, int k, void* workspace,
                                    int n_split, const float* levels, int n_levels, int32_t* count, vfr_stream_t stream) {
  VFR_REQUIRE(query_packed && workspace && levels && count && n_levels >= 1 && n_levels <= 4, VFR_ERR_INVALID,
              "vfr_sel_count_levels: bad argument");
  SlPlan pl;
  SlParams p;
  int rc = sl_setup(pl, p, query_packed, n_queries, n_clips, dim, k, workspace, n_split);
  if (rc) return rc;
  sl_count_levels_kernel<<<(unsigned)((n_queries + 7) / 8), 256, 0, (cudaStream_t)stream>>>(
      p.cand, p.cand_cnt, p.n_parts, p.qmeta, n_queries, levels, n_levels, count);
  return check_launch("sl_count_levels_kernel");
}

extern "C" int vfr_sel_pick_put(void* query_packed, int64_t n_queries, int64_t n_clips, int dim, int k, void* workspace,
                                int n_split, const float* levels, const int32_t* count, int n_levels, vfr_stream_t stream) {
  VFR_REQUIRE(query_packed && workspace && levels && count && n_levels >= 1 && n_levels <= 4, VFR_ERR_INVALID,
              "vfr_sel_pick_put: bad argument");
  SlPlan pl;
  SlParams p;
  int rc = sl_setup(pl, p, query_packed, n_queries, n_clips, dim, k, workspace, n_split);
  if (rc) return rc;
  sl_pick_put_kernel<<<(unsigned)((n_queries + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
      levels, count, n_levels, k, p.tau_g, p.tau_cert, p.qmeta, p.flags, n_queries);
  return check_launch("sl_pick_put_kernel");
}

extern "C" size_t vfr_topk_block_bytes(int64_t per, int k) { return (per > 0 && k > 0) ? (size_t)per * ((size_t)k * 12 + 4) : 0; }

extern "C" int vfr_sel_refine_blocks(const float* bank, const int32_t* vid_off, const int64_t* mom_off, int64_t n_videos,
                                     int64_t n_clips, int n_max, int dim, void* query_packed, const float* queries,
                                     int64_t n_queries, int k, int64_t id_base, int64_t per, void* out_blocks, void* workspace,
                                     int n_split, vfr_stream_t stream) {
  VFR_REQUIRE(query_packed && workspace && out_blocks && per > 0, VFR_ERR_INVALID, "vfr_sel_refine_blocks: bad argument");
  SlPlan pl;
  SlParams p;
  int rc = sl_setup(pl, p, query_packed, n_queries, n_clips, dim, k, workspace, n_split);
  if (rc) return rc;
  return sl_run_refine(p, bank, vid_off, mom_off, n_videos, n_max, dim, queries, n_queries, k, id_base, nullptr, nullptr,
                       (cudaStream_t)stream, per, out_blocks);
}

extern "C" int vfr_sel_stats(void* query_packed, int64_t n_queries, int64_t n_clips, int dim, int k, void* workspace,
                             int n_split, int64_t* out, vfr_stream_t stream) {
  VFR_REQUIRE(query_packed && workspace && out, VFR_ERR_INVALID, "vfr_sel_stats: null pointer");
  SlPlan pl;
  SlParams p;
  int rc = sl_setup(pl, p, query_packed, n_queries, n_clips, dim, k, workspace, n_split);
  if (rc) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  VFR_CUDA(cudaMemsetAsync(out, 0, 8 * sizeof(int64_t), st));
  sl_stats_kernel<<<(unsigned)((n_queries + 255) / 256), 256, 0, st>>>(p.cand_cnt, p.n_parts, p.flags, n_queries, p.stats,
                                                                      reinterpret_cast<unsigned long long*>(out));
  return check_launch("sl_stats_kernel");
}

extern "C" int64_t vfr_sel_tiles(int64_t n_clips) { return n_clips > 0 ? sl_tiles(n_clips) : 0; }

extern "C" int vfr_sel_sample_rank(int k, int64_t n_sampled, int64_t n_total) {
  if (k < 1 || n_sampled <= 0 || n_total <= 0) return 0;
  const double x = (double)k * (double)n_sampled / (double)n_total;
  double term = exp(-x), cdf = 0.0;
  for (int j = 1; j <= SL_J; ++j) {
    cdf += term;
    term *= x / j;
    if (1.0 - cdf < 1e-10) return j;
  }
  return 0;
}

extern "C" int vfr_sel_sample_lists(int64_t n_queries, int64_t n_clips, int n_split, int dim) {
  if (n_queries <= 0 || n_clips <= 0 || !sl_dim_ok(dim)) return 0;
  return sl_plan(n_queries, n_clips, n_split, dim).n_parts;
}

extern "C" int64_t vfr_sel_sample_clips(int64_t n_queries, int64_t n_clips, int k, int n_split, int dim) {
  if (n_queries <= 0 || n_clips <= 0 || k < 1 || !sl_dim_ok(dim)) return 0;
  const SlPlan pl = sl_plan(n_queries, n_clips, n_split, dim);
  return (int64_t)sl_sample_plan(pl, n_clips, k, /*need_rank=*/false).tiles * pl.n_parts * SL_N;
}

extern "C" int vfr_sel_sample(const void* bank_packed, int64_t n_clips, int dim, void* query_packed, int64_t n_queries,
                              int k, void* workspace, int n_split, float* out, int64_t* n_sampled, vfr_stream_t stream) {
  VFR_REQUIRE(bank_packed && query_packed && workspace && out && n_sampled, VFR_ERR_INVALID, "vfr_sel_sample: null pointer");
  SlPlan pl;
  SlParams p;
  int rc = sl_setup(pl, p, query_packed, n_queries, n_clips, dim, k, workspace, n_split);
  if (rc) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  const size_t qpad = (size_t)pl.qrows;
  rc = launch_fill_u32(p.tau_g, 0x7f800000u, qpad * (2 + (size_t)pl.n_parts), st);   // fresh thresholds and lists
  if (rc) return rc;
  VFR_CUDA(cudaMemsetAsync(p.cand_cnt, 0, qpad * (size_t)pl.n_parts * sizeof(int32_t), st));
  VFR_CUDA(cudaMemsetAsync(p.stats, 0, 8 * sizeof(unsigned long long), st));
  const SlSample sp = sl_sample_plan(pl, n_clips, k, /*need_rank=*/false);
  *n_sampled = (int64_t)sp.tiles * pl.n_parts * SL_N;
  if (sp.tiles == 0) return VFR_OK;
  CUtensorMap ma, mb;
  rc = sl_make_map(&ma, query_packed, (uint64_t)pl.qrows, SL_M, pl.pitch);
  if (rc) return rc;
  rc = sl_make_map(&mb, bank_packed, (uint64_t)(pl.n_tiles * SL_N), SL_N, pl.pitch);
  if (rc) return rc;
  float* samp = sl_samp_buffer(pl, p);
  rc = sl_launch_sample(pl, p, sp, samp, ma, mb, st);
  if (rc) return rc;
  VFR_CUDA(cudaMemcpyAsync(out, samp, (size_t)n_queries * pl.n_parts * SL_J * sizeof(float), cudaMemcpyDeviceToDevice, st));
  return VFR_OK;
}

extern "C" int vfr_sel_count_under(void* query_packed, int64_t n_queries, int64_t n_clips, int dim, int k, void* workspace,
                                   int n_split, const float* bound, int32_t* count, vfr_stream_t stream) {
  VFR_REQUIRE(query_packed && workspace && bound && count, VFR_ERR_INVALID, "vfr_sel_count_under: null pointer");
  SlPlan pl;
  SlParams p;
  int rc = sl_setup(pl, p, query_packed, n_queries, n_clips, dim, k, workspace, n_split);
  if (rc) return rc;
  sl_count_under_kernel<<<(unsigned)((n_queries + 7) / 8), 256, 0, (cudaStream_t)stream>>>(p.cand, p.cand_cnt, p.n_parts, p.qmeta,
                                                                                        n_queries, bound, count);
  return check_launch("sl_count_under_kernel");
}

extern "C" int vfr_sel_filter(const void* bank_packed, int64_t n_clips, int dim, void* query_packed, int64_t n_queries,
                              int k, void* workspace, int n_split, int64_t tile_lo, int64_t tile_hi, int resume,
                              vfr_stream_t stream) {
  VFR_REQUIRE(bank_packed && query_packed && workspace, VFR_ERR_INVALID, "vfr_sel_filter: null pointer");
  SlPlan pl;
  SlParams p;
  int rc = sl_setup(pl, p, query_packed, n_queries, n_clips, dim, k, workspace, n_split);
  if (rc) return rc;
  return sl_run_filter(pl, p, bank_packed, query_packed, tile_lo, tile_hi, resume, (cudaStream_t)stream);
}

extern "C" int vfr_sel_bound_get(void* query_packed, int64_t n_queries, int64_t n_clips, int dim, int k, void* workspace,
                                 int n_split, float* bound, vfr_stream_t stream) {
  VFR_REQUIRE(query_packed && workspace && bound, VFR_ERR_INVALID, "vfr_sel_bound_get: null pointer");
  SlPlan pl;
  SlParams p;
  int rc = sl_setup(pl, p, query_packed, n_queries, n_clips, dim, k, workspace, n_split);
  if (rc) return rc;
  sl_bound_get_kernel<<<(unsigned)((n_queries + 255) / 256), 256, 0, (cudaStream_t)stream>>>(p.tau_cert, p.qmeta, n_queries, bound);
  return check_launch("sl_bound_get_kernel");
}

extern "C" int vfr_sel_bound_put(void* query_packed, int64_t n_queries, int64_t n_clips, int dim, int k, void* workspace,
                                 int n_split, const float* bound, vfr_stream_t stream) {
  VFR_REQUIRE(query_packed && workspace && bound, VFR_ERR_INVALID, "vfr_sel_bound_put: null pointer");
  SlPlan pl;
  SlParams p;
  int rc = sl_setup(pl, p, query_packed, n_queries, n_clips, dim, k, workspace, n_split);
  if (rc) return rc;
  sl_bound_put_kernel<<<(unsigned)((n_queries + 255) / 256), 256, 0, (cudaStream_t)stream>>>(p.tau_g, p.tau_cert, p.qmeta, n_queries, bound);
  return check_launch("sl_bound_put_kernel");
}

extern "C" int vfr_sel_refine(const float* bank, const int32_t* vid_off, const int64_t* mom_off, int64_t n_videos,
                              int64_t n_clips, int n_max, int dim, void* query_packed, const float* queries,
                              int64_t n_queries, int k, int64_t id_base, float* out_scores, int64_t* out_ids,
                              void* workspace, int n_split, vfr_stream_t stream) {
  VFR_REQUIRE(query_packed && workspace, VFR_ERR_INVALID, "vfr_sel_refine: null pointer");
  SlPlan pl;
  SlParams p;
  int rc = sl_setup(pl, p, query_packed, n_queries, n_clips, dim, k, workspace, n_split);
  if (rc) return rc;
  return sl_run_refine(p, bank, vid_off, mom_off, n_videos, n_max, dim, queries, n_queries, k, id_base, out_scores,
                       out_ids, (cudaStream_t)stream);
}

// stage 2 for the queries [q_begin, q_begin + q_count) only (rows q of out_scores / out_ids): lets a caller copy finished
// rows to the host while the rest is still being re-scored (vfr_search_host)
extern "C" int vfr_sel_refine_range(const float* bank, const int32_t* vid_off, const int64_t* mom_off, int64_t n_videos,
                                    int64_t n_clips, int n_max, int dim, void* query_packed, const float* queries,
                                    int64_t n_queries, int k, int64_t id_base, float* out_scores, int64_t* out_ids,
                                    void* workspace, int n_split, int64_t q_begin, int64_t q_count, vfr_stream_t stream) {
  VFR_REQUIRE(query_packed && workspace, VFR_ERR_INVALID, "vfr_sel_refine_range: null pointer");
  SlPlan pl;
  SlParams p;
  int rc = sl_setup(pl, p, query_packed, n_queries, n_clips, dim, k, workspace, n_split);
  if (rc) return rc;
  return sl_run_refine(p, bank, vid_off, mom_off, n_videos, n_max, dim, queries, n_queries, k, id_base, out_scores,
                       out_ids, (cudaStream_t)stream, 0, nullptr, false, q_begin, q_count);
}

extern "C" int vfr_sel_topk(const void* bank_packed, const float* bank, const int32_t* vid_off, const int64_t* mom_off,
                            int64_t n_videos, int64_t n_clips, int n_max, int dim, void* query_packed,
                            const float* queries, int64_t n_queries, int k, int64_t id_base, float* out_scores,
                            int64_t* out_ids, void* workspace, int n_split, vfr_stream_t stream) {
  VFR_REQUIRE(bank_packed && bank && vid_off && mom_off && query_packed && queries && out_scores && out_ids && workspace,
              VFR_ERR_INVALID, "vfr_sel_topk: null pointer");
  SlPlan pl;
  SlParams p;
  int rc = sl_setup(pl, p, query_packed, n_queries, n_clips, dim, k, workspace, n_split);
  if (rc) return rc;
  rc = sl_run_filter(pl, p, bank_packed, query_packed, 0, -1, 0, (cudaStream_t)stream);
  if (rc) return rc;
  return sl_run_refine(p, bank, vid_off, mom_off, n_videos, n_max, dim, queries, n_queries, k, id_base, out_scores,
                       out_ids, (cudaStream_t)stream);
}

// bf16 embedding path: the bank's fp32 rows are replaced by a bf16 [C, dim] array (half the HBM of stage 2); the packed
// fp16 operand must have been built from the same (bf16-representable) values.  Scores = the exact engine's on those values.
extern "C" int vfr_sel_topk_b16(const void* bank_packed, const void* bank_b16, const int32_t* vid_off, const int64_t* mom_off,
                                int64_t n_videos, int64_t n_clips, int n_max, int dim, void* query_packed,
                                const float* queries, int64_t n_queries, int k, int64_t id_base, float* out_scores,
                                int64_t* out_ids, void* workspace, int n_split, vfr_stream_t stream) {
  VFR_REQUIRE(bank_packed && bank_b16 && vid_off && mom_off && query_packed && queries && out_scores && out_ids && workspace,
              VFR_ERR_INVALID, "vfr_sel_topk_b16: null pointer");
  SlPlan pl;
  SlParams p;
  int rc = sl_setup(pl, p, query_packed, n_queries, n_clips, dim, k, workspace, n_split);
  if (rc) return rc;
  rc = sl_run_filter(pl, p, bank_packed, query_packed, 0, -1, 0, (cudaStream_t)stream);
  if (rc) return rc;
  return sl_run_refine(p, reinterpret_cast<const float*>(bank_b16), vid_off, mom_off, n_videos, n_max, dim, queries, n_queries,
                       k, id_base, out_scores, out_ids, (cudaStream_t)stream, 0, nullptr, true);
}

// flags of the last vfr_sel_topk on this packed query buffer: device pointer to int32 [n_queries]
// (0 = exact result; 1 = operand scales out of fp16 range, 2 = candidate list overflow, 3 = refine overflow)
extern "C" const int32_t* vfr_sel_flags(const void* query_packed, int64_t n_queries, int dim) {
  if (!query_packed || n_queries <= 0 || !sl_dim_ok(dim)) return nullptr;
  const int64_t rows = sl_qrows(n_queries);
  const __half* qp = reinterpret_cast<const __half*>(query_packed);
  return reinterpret_cast<const int32_t*>(reinterpret_cast<const float4*>(qp + rows * sl_pitch(dim)) + rows);
}
