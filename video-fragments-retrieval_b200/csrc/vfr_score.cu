// K4 (exact-fp32 path): query x clip L2 distance -> moment means -> full / count / top-k.
//
// Replaces reference model/evaluate.py:49-58,71-80 (hot loops 1+2 and the argsort),
// model/evaluate_single.py:48-53, model/main.py:148-157.
//
// Layout (B200-first): both operands are pre-packed k-major in HBM so that one pipeline stage =
// one contiguous block moved by a single cp.async.bulk (UBLKCP) onto an mbarrier:
//   bank  packed [tile][kchunk][KC][96]   then [tile][96]  row sums  (sum_k v_k)
//   query packed [qtile][kchunk][KC][128] then [qtile][128] row sums (sum_k q_k)
// A CTA owns one 128-query tile and walks a contiguous range of bank tiles; 256 threads each keep
// an 8-query x 6-clip block of squared distances in registers.  Direct-difference form: with
// delta = v - q (exact for near-duplicates, unlike the GEMM expansion),
//   sum_k (delta_k + eps)^2 = sum_k delta_k^2 + 2 eps (sum_k v_k - sum_k q_k) + D eps^2,
// so the inner loop is one FADD + one FFMA per element and the pairwise_distance eps enters as a
// per-pair correction from the pre-computed row sums.  k is strictly sequential, so every mode and
// vfr_score_own produce bit-identical values.  After the last k-chunk the distances go through
// shared memory to the moment phase, where a thread owns ONE query (so count / top-k state is
// thread-private: no atomics in the loop) and walks the tile's videos.
#include "vfr_common.cuh"
#include "vfr_topk.cuh"
#include <math_constants.h>

namespace vfr {

constexpr int TQ = VFR_TILE_Q;      // 128
constexpr int TC = VFR_TILE_C;      // 96
constexpr int KC = 20;              // k per pipeline stage
constexpr int STAGES = 3;
constexpr int NTHREADS = 256;
constexpr int DS_LD = TC + 1;       // 97: conflict-free row walks
constexpr int Q_STAGE = KC * TQ;    // floats
constexpr int V_STAGE = KC * TC;    // floats
constexpr int STAGE_FLOATS = Q_STAGE + V_STAGE;
constexpr uint32_t STAGE_BYTES = STAGE_FLOATS * 4;
constexpr int CAP_HI = CAP - VFR_MAX_SEG;  // compaction trigger: at most 32 appends between checks

enum Mode { MODE_FULL = 0, MODE_COUNT = 1, MODE_TOPK = 2 };

struct ScoreParams {
  const float* bank_packed;
  const float* query_packed;
  const float* bank_rowsum;   // [n_tiles][96]
  const float* query_rowsum;  // [q_tiles][128]
  float deps2;                // D * eps^2
  const int32_t* vid_off;
  const int64_t* mom_off;
  int64_t n_videos;
  int64_t n_queries;
  int vt;          // videos per tile
  int n_tiles;     // bank tiles
  int nkc;         // k chunks
  int tiles_per_split;
  // FULL
  float* out_full;
  int64_t m_total;
  // COUNT
  const float* tau;
  int n_tau;
  const int32_t* q_video;
  uint32_t* cnt_lt;
  uint32_t* cnt_eqb;
  // TOPK
  int k;
  unsigned long long* cand;
  int32_t* cand_cnt;
  int n_parts;
  unsigned* tau_g;            // [Qpad] shared per-query threshold
};

static inline int nkc_of(int dim) { return (dim + KC - 1) / KC; }

// ---------------------------------------------------------------------------------------------
// packing
// ---------------------------------------------------------------------------------------------
// sequential fp32 row sum: the same loop in the pack kernels and in score_own (bit-identical)
__device__ __forceinline__ float row_sum(const float* __restrict__ r, int dim) {
  float s = 0.f;
  for (int k = 0; k < dim; ++k) s = __fadd_rn(s, r[k]);
  return s;
}

// d^2 from sum delta^2 and the row sums (see the header comment)
__device__ __forceinline__ float dist_from(float s2, float sv, float sq, float deps2) {
  const float corr = __fmaf_rn(2.f * VFR_PAIRWISE_EPS, __fsub_rn(sv, sq), deps2);
  return __fsqrt_rn(fmaxf(__fadd_rn(s2, corr), 0.f));
}

__global__ void pack_bank_kernel(const float* __restrict__ bank, const int32_t* __restrict__ vid_off,
                                 int64_t n_videos, int vt, int dim, int nkc, float* __restrict__ packed,
                                 float* __restrict__ rowsum) {
  const int64_t tile = blockIdx.x;
  const int64_t v0 = tile * vt;
  const int64_t v1 = min(v0 + (int64_t)vt, n_videos);
  const int64_t c0 = vid_off[v0];
  const int ncols = (int)(vid_off[v1] - c0);
  float* dst = packed + tile * (int64_t)nkc * V_STAGE;
  const int total = nkc * KC * TC;
  for (int i = threadIdx.x; i < total; i += blockDim.x) {
    const int col = i % TC;
    const int k = i / TC;
    float val = 0.f;
    if (col < ncols && k < dim) val = __ldg(bank + (c0 + col) * dim + k);
    dst[i] = val;  // [k][col] with k = chunk*KC + kk : chunks are contiguous
  }
  if (threadIdx.x < TC) {
    const int col = threadIdx.x;
    rowsum[tile * TC + col] = (col < ncols) ? row_sum(bank + (c0 + col) * dim, dim) : 0.f;
  }
}

__global__ void pack_query_kernel(const float* __restrict__ q, int64_t n_queries, int dim, int nkc,
                                  float* __restrict__ packed, float* __restrict__ rowsum) {
  const int64_t tile = blockIdx.x;
  float* dst = packed + tile * (int64_t)nkc * Q_STAGE;
  const int total = nkc * KC * TQ;
  for (int i = threadIdx.x; i < total; i += blockDim.x) {
    const int row = i % TQ;
    const int k = i / TQ;
    const int64_t qi = tile * TQ + row;
    dst[i] = (qi < n_queries && k < dim) ? __ldg(q + qi * dim + k) : 0.f;
  }
  if (threadIdx.x < TQ) {
    const int64_t qi = tile * TQ + threadIdx.x;
    rowsum[tile * TQ + threadIdx.x] = (qi < n_queries) ? row_sum(q + qi * dim, dim) : 0.f;
  }
}

// ---------------------------------------------------------------------------------------------
// main kernel
// ---------------------------------------------------------------------------------------------
template <int MODE, int NTAU>
__global__ void __launch_bounds__(NTHREADS, 2) score_kernel(const ScoreParams p) {
  extern __shared__ __align__(128) float smem[];
  float* stage_base = smem;                                  // STAGES x (q stage | v stage)
  float* ds = smem + STAGES * STAGE_FLOATS;                  // [TQ][DS_LD]
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(ds + TQ * DS_LD + 1);  // keep 8-B alignment below
  full_bar = reinterpret_cast<uint64_t*>((reinterpret_cast<uintptr_t>(full_bar) + 7) & ~uintptr_t(7));

  const int tid = threadIdx.x;
  const int lane = tid & 31;
  const int tq = tid & 15;   // query group: queries 4*tq..+3 and 64+4*tq..+3
  const int tc = tid >> 4;   // clip group : columns 6*tc..+5
  const int qtile = blockIdx.x;
  const int split = blockIdx.y;
  const int tile_begin = split * p.tiles_per_split;
  const int tile_end = min(tile_begin + p.tiles_per_split, p.n_tiles);
  const int n_my_tiles = tile_end - tile_begin;
  if (n_my_tiles <= 0) return;
  const int total_iters = n_my_tiles * p.nkc;

  if (tid == 0) {
    for (int s = 0; s < STAGES; ++s) mbar_init(&full_bar[s], 1);
    fence_barrier_init();
  }
  __syncthreads();

  const float* qsrc = p.query_packed + (int64_t)qtile * p.nkc * Q_STAGE;
  auto issue = [&](int it) {
    const int s = it % STAGES;
    const int tile = tile_begin + it / p.nkc;
    const int kc = it % p.nkc;
    float* dst = stage_base + s * STAGE_FLOATS;
    mbar_expect_tx(&full_bar[s], STAGE_BYTES);
    bulk_g2s(dst, qsrc + (int64_t)kc * Q_STAGE, Q_STAGE * 4, &full_bar[s]);
    bulk_g2s(dst + Q_STAGE, p.bank_packed + ((int64_t)tile * p.nkc + kc) * V_STAGE, V_STAGE * 4, &full_bar[s]);
  };
  if (tid == 0) {
    for (int it = 0; it < STAGES - 1 && it < total_iters; ++it) issue(it);
  }

  // moment-phase ownership: one query per thread, two threads (halves) per query
  const int mq = tid & (TQ - 1);
  const int half = tid >> 7;
  const int64_t q_global = (int64_t)qtile * TQ + mq;
  const bool q_valid = q_global < p.n_queries;

  // per-mode thread-private state
  float tau_r[NTAU];
  unsigned lt[NTAU], eqb[NTAU];
  int my_qvid = -1;
  if (MODE == MODE_COUNT) {
#pragma unroll
    for (int t = 0; t < NTAU; ++t) {
      tau_r[t] = (q_valid && t < p.n_tau) ? p.tau[q_global * p.n_tau + t] : -CUDART_INF_F;
      lt[t] = 0;
      eqb[t] = 0;
    }
    if (q_valid) my_qvid = p.q_video[q_global];
  }
  unsigned long long* my_list = nullptr;
  int my_cnt = 0;
  float my_tau = CUDART_INF_F;
  const int part = split * 2 + half;
  if (MODE == MODE_TOPK) {
    // (invalid queries of a ragged last tile get a dummy list at slot of query n_queries-1? no:
    //  the workspace is sized for whole tiles, so every thread owns real storage)
    my_list = p.cand + ((int64_t)q_global * p.n_parts + part) * CAP;
  }

  float acc[8][6];

  for (int it = 0; it < total_iters; ++it) {
    const int s = it % STAGES;
    const int kc = it % p.nkc;
    if (tid == 0 && it + STAGES - 1 < total_iters) issue(it + STAGES - 1);
    if (kc == 0) {
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 6; ++j) acc[i][j] = 0.f;
    }
    mbar_wait(&full_bar[s], (it / STAGES) & 1);
    const float* qs = stage_base + s * STAGE_FLOATS;
    const float* vs = qs + Q_STAGE;
#pragma unroll 4
    for (int kk = 0; kk < KC; ++kk) {
      const float4 qa = *reinterpret_cast<const float4*>(qs + kk * TQ + 4 * tq);
      const float4 qb = *reinterpret_cast<const float4*>(qs + kk * TQ + 64 + 4 * tq);
      const float2 va = *reinterpret_cast<const float2*>(vs + kk * TC + 6 * tc);
      const float2 vb = *reinterpret_cast<const float2*>(vs + kk * TC + 6 * tc + 2);
      const float2 vc = *reinterpret_cast<const float2*>(vs + kk * TC + 6 * tc + 4);
      const float qv[8] = {qa.x, qa.y, qa.z, qa.w, qb.x, qb.y, qb.z, qb.w};
      const float vv[6] = {va.x, va.y, vb.x, vb.y, vc.x, vc.y};
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 6; ++j) {
          const float d = __fsub_rn(vv[j], qv[i]);
          acc[i][j] = __fmaf_rn(d, d, acc[i][j]);
        }
    }
    if (kc != p.nkc - 1) {
      __syncthreads();  // everyone is done with stage s before it is refilled (issue at it+1)
      continue;
    }

    // ---- distances of this tile -> shared ----
    {
      const int tile_e = tile_begin + it / p.nkc;
      float sv[6];
#pragma unroll
      for (int j = 0; j < 6; ++j) sv[j] = __ldg(p.bank_rowsum + (int64_t)tile_e * TC + 6 * tc + j);
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int ql = (i < 4) ? (4 * tq + i) : (64 + 4 * tq + (i - 4));
        const float sq = __ldg(p.query_rowsum + (int64_t)qtile * TQ + ql);
#pragma unroll
        for (int j = 0; j < 6; ++j) ds[ql * DS_LD + 6 * tc + j] = dist_from(acc[i][j], sv[j], sq, p.deps2);
      }
    }
    __syncthreads();  // ds complete; also releases stage s

    if (MODE == MODE_TOPK) my_tau = fminf(my_tau, tau_fetch(p.tau_g + q_global));
    // ---- moment phase: thread = (query mq, half) walks the videos of the tile ----
    const int tile = tile_begin + it / p.nkc;
    const int64_t v0 = (int64_t)tile * p.vt;
    const int nv = (int)min((int64_t)p.vt, p.n_videos - v0);
    const int c0 = p.vid_off[v0];
    const float* drow = ds + mq * DS_LD;
    for (int j = half; j < nv; j += 2) {
      const int64_t v = v0 + j;
      const int cb = p.vid_off[v] - c0;
      const int n = p.vid_off[v + 1] - p.vid_off[v];
      int64_t mbase = 0;
      if (MODE != MODE_COUNT) mbase = p.mom_off[v];
      for (int sidx = 0; sidx < n; ++sidx) {
        float run = 0.f;
        for (int e = sidx; e < n; ++e) {
          run = __fadd_rn(run, drow[cb + e]);
          const float score = __fdiv_rn(run, (float)(e - sidx + 1));
          const int m = moment_index(n, sidx, e);
          if (MODE == MODE_FULL) {
            if (q_valid) p.out_full[q_global * p.m_total + mbase + m] = score;
          } else if (MODE == MODE_COUNT) {
#pragma unroll
            for (int t = 0; t < NTAU; ++t) {
              lt[t] += (score < tau_r[t]) ? 1u : 0u;
              eqb[t] += (score == tau_r[t] && v < my_qvid) ? 1u : 0u;
            }
          } else {
            if (score <= my_tau) {
              my_list[my_cnt++] =
                  ((unsigned long long)__float_as_uint(score) << 32) | (unsigned)(mbase + m);
            }
          }
        }
        if (MODE == MODE_TOPK) {
          const float before = my_tau;
          compact_lists(my_list, my_cnt, my_tau, p.k, my_cnt > CAP_HI, lane);
          if (my_tau < before) tau_publish(p.tau_g + q_global, my_tau);
        }
      }
    }
    __syncthreads();  // ds free for the next tile
  }

  if (MODE == MODE_COUNT) {
    if (q_valid) {
#pragma unroll
      for (int t = 0; t < NTAU; ++t) {
        if (t < p.n_tau) {
          if (lt[t]) atomicAdd(p.cnt_lt + q_global * p.n_tau + t, lt[t]);
          if (eqb[t]) atomicAdd(p.cnt_eqb + q_global * p.n_tau + t, eqb[t]);
        }
      }
    }
  }
  if (MODE == MODE_TOPK) p.cand_cnt[q_global * p.n_parts + part] = my_cnt;
}

// ---------------------------------------------------------------------------------------------
// own-video scores: one thread per (query, clip) for the distances, one per query for the means
// ---------------------------------------------------------------------------------------------
__global__ void score_own_kernel(const float* __restrict__ bank, const int32_t* __restrict__ vid_off, int dim,
                                 const float* __restrict__ queries, int64_t n_queries,
                                 const int32_t* __restrict__ q_video, float* __restrict__ out, int m_stride) {
  __shared__ float d[8][VFR_MAX_SEG];
  const int ql = threadIdx.y;
  const int64_t q = (int64_t)blockIdx.x * 8 + ql;
  const int c = threadIdx.x;
  int n = 0;
  if (q < n_queries) {
    const int v = q_video[q];
    const int c0 = vid_off[v];
    n = vid_off[v + 1] - c0;
    if (c < n) {
      const float* vr = bank + (int64_t)(c0 + c) * dim;
      const float* qr = queries + q * dim;
      float acc = 0.f;
      for (int k = 0; k < dim; ++k) {
        const float diff = __fsub_rn(vr[k], qr[k]);
        acc = __fmaf_rn(diff, diff, acc);
      }
      d[ql][c] = dist_from(acc, row_sum(vr, dim), row_sum(qr, dim), (float)dim * VFR_PAIRWISE_EPS * VFR_PAIRWISE_EPS);
    }
  }
  __syncthreads();
  if (q < n_queries) {
    // lanes share the (s, e) loop: lane c handles start s = c
    const int s = c;
    if (s < n) {
      float run = 0.f;
      for (int e = s; e < n; ++e) {
        run = __fadd_rn(run, d[ql][e]);
        out[q * m_stride + moment_index(n, s, e)] = __fdiv_rn(run, (float)(e - s + 1));
      }
    }
    for (int m = num_moments(n) + c; m < m_stride; m += VFR_MAX_SEG) out[q * m_stride + m] = CUDART_INF_F;
  }
}

// ---------------------------------------------------------------------------------------------
// per-query merge of the part lists (block bitonic sort in shared memory)
// ---------------------------------------------------------------------------------------------
constexpr int MERGE_N = 2048;
constexpr int MERGE_THREADS = 256;

__device__ __forceinline__ void block_sort(unsigned long long* keys /*MERGE_N*/) {
  for (int size = 2; size <= MERGE_N; size <<= 1) {
    for (int stride = size >> 1; stride > 0; stride >>= 1) {
      __syncthreads();
      for (int i = threadIdx.x; i < MERGE_N / 2; i += MERGE_THREADS) {
        const int lo = ((i & ~(stride - 1)) << 1) | (i & (stride - 1));
        const int hi = lo | stride;
        const bool asc = !(lo & size) || size == MERGE_N;
        const unsigned long long a = keys[lo], b = keys[hi];
        if ((a > b) == asc) { keys[lo] = b; keys[hi] = a; }
      }
    }
  }
  __syncthreads();
}

// Part lists are unsorted and hold up to CAP keys each.  All keys of a query are streamed through a
// 2048-key shared-memory bitonic sort: slots [0,128) carry the running best, slots [128,2048) the
// next batch.
__global__ void __launch_bounds__(MERGE_THREADS) topk_finish_kernel(const unsigned long long* __restrict__ cand,
                                                                     const int32_t* __restrict__ cand_cnt,
                                                                     int n_parts, int k, int64_t id_base,
                                                                     float* __restrict__ out_scores,
                                                                     int64_t* __restrict__ out_ids) {
  __shared__ unsigned long long keys[MERGE_N];
  const int64_t q = blockIdx.x;
  constexpr int BATCH = MERGE_N - VFR_TOPK_MAX;
  for (int i = threadIdx.x; i < VFR_TOPK_MAX; i += MERGE_THREADS) keys[i] = ~0ull;
  int part = 0, pos = 0;          // stream position: next key is cand[(q*n_parts+part)*CAP + pos]
  while (part < n_parts) {
    __syncthreads();
    // every thread walks the same (part, pos) sequence; thread i of the batch takes the i-th key
    int filled = 0, pp = part, po = pos;
    while (pp < n_parts && filled < BATCH) {
      const int64_t li = q * n_parts + pp;
      const int cnt = min(cand_cnt[li], CAP);
      const int take = min(cnt - po, BATCH - filled);
      for (int i = threadIdx.x; i < take; i += MERGE_THREADS) keys[VFR_TOPK_MAX + filled + i] = cand[li * CAP + po + i];
      filled += take;
      po += take;
      if (po >= cnt) { ++pp; po = 0; }
    }
    for (int i = VFR_TOPK_MAX + filled + threadIdx.x; i < MERGE_N; i += MERGE_THREADS) keys[i] = ~0ull;
    part = pp;
    pos = po;
    block_sort(keys);
    for (int i = k + threadIdx.x; i < VFR_TOPK_MAX; i += MERGE_THREADS) keys[i] = ~0ull;
  }
  __syncthreads();
  for (int i = threadIdx.x; i < k; i += MERGE_THREADS) {
    const unsigned long long key = keys[i];
    const bool ok = key != ~0ull;
    out_scores[q * k + i] = ok ? __uint_as_float((unsigned)(key >> 32)) : CUDART_INF_F;
    out_ids[q * k + i] = ok ? id_base + (int64_t)(unsigned)(key & 0xffffffffu) : -1;
  }
}

// K7: merge of P sorted per-shard lists with 64-bit ids (after the all-gather)
struct SI { unsigned s; int64_t id; };
__device__ __forceinline__ bool si_gt(const SI& a, const SI& b) { return a.s > b.s || (a.s == b.s && a.id > b.id); }

// part pi of the input starts stride_s floats / stride_i int64s after part pi - 1 (flat [P, Q, k] arrays: Q k for both;
// the blocked records of a sharded search: one record size for both, see vfr_topk_merge_blocks)
__global__ void __launch_bounds__(MERGE_THREADS) topk_merge_kernel(const float* __restrict__ in_scores,
                                                                    const int64_t* __restrict__ in_ids, int n_parts,
                                                                    int64_t stride_s, int64_t stride_i, int k, int n_pad,
                                                                    float* __restrict__ out_scores,
                                                                    int64_t* __restrict__ out_ids,
                                                                    const int32_t* __restrict__ in_flags, int64_t stride_f,
                                                                    int32_t* __restrict__ out_flags,
                                                                    unsigned long long* __restrict__ n_flagged) {
  extern __shared__ __align__(16) unsigned char raw[];
  unsigned* ss = reinterpret_cast<unsigned*>(raw);                         // [n_pad]
  int64_t* ids = reinterpret_cast<int64_t*>(raw + (size_t)n_pad * 8);      // [n_pad] (8-B aligned)
  const int64_t q = blockIdx.x;
  const int total = n_parts * k;
  for (int i = threadIdx.x; i < n_pad; i += MERGE_THREADS) {
    if (i < total) {
      const int pi = i / k, j = i % k;
      const int64_t id = in_ids[(int64_t)pi * stride_i + q * k + j];
      ss[i] = id < 0 ? 0xffffffffu : __float_as_uint(in_scores[(int64_t)pi * stride_s + q * k + j]);
      ids[i] = id < 0 ? INT64_MAX : id;
    } else {
      ss[i] = 0xffffffffu;
      ids[i] = INT64_MAX;
    }
  }
  if (in_flags && threadIdx.x == 0) {
    // a query is only as good as its worst shard list: OR of the shards' filter + refine flags
    int32_t f = 0;
    for (int pi = 0; pi < n_parts; ++pi) f |= in_flags[(int64_t)pi * stride_f + q];
    out_flags[q] = f;
    if (f && n_flagged) atomicAdd(n_flagged, 1ull);
  }
  // The lists of a sharded search arrive sorted (ascending (score, id), padding last): then every key's place in the
  // merged order is its own index plus, per other list, the number of keys before it (binary search) - no sort.
  __syncthreads();
  bool unsorted = false;
  for (int i = threadIdx.x; i < total; i += MERGE_THREADS)
    if (i % k != 0 && si_gt(SI{ss[i - 1], ids[i - 1]}, SI{ss[i], ids[i]})) unsorted = true;
  if (!__syncthreads_or(unsorted)) {
    for (int i = threadIdx.x; i < k; i += MERGE_THREADS) { out_scores[q * k + i] = CUDART_INF_F; out_ids[q * k + i] = -1; }
    __syncthreads();
    for (int i = threadIdx.x; i < total; i += MERGE_THREADS) {
      const SI me{ss[i], ids[i]};
      if (me.id == INT64_MAX) continue;
      const int pi = i / k;
      int rank = i - pi * k;
      for (int b = 0; b < n_parts && rank < k; ++b) {
        if (b == pi) continue;
        int lo = 0, hi = k;                      // first index of list b whose key is not before mine
        while (lo < hi) {
          const int mid = (lo + hi) >> 1;
          if (si_gt(me, SI{ss[b * k + mid], ids[b * k + mid]})) lo = mid + 1; else hi = mid;
        }
        rank += lo;
      }
      if (rank < k) { out_scores[q * k + rank] = __uint_as_float(me.s); out_ids[q * k + rank] = me.id; }
    }
    return;
  }
  for (int size = 2; size <= n_pad; size <<= 1) {
    for (int stride = size >> 1; stride > 0; stride >>= 1) {
      __syncthreads();
      for (int i = threadIdx.x; i < n_pad / 2; i += MERGE_THREADS) {
        const int lo = ((i & ~(stride - 1)) << 1) | (i & (stride - 1));
        const int hi = lo | stride;
        const bool asc = !(lo & size) || size == n_pad;
        const SI a{ss[lo], ids[lo]}, b{ss[hi], ids[hi]};
        if (si_gt(a, b) == asc) { ss[lo] = b.s; ids[lo] = b.id; ss[hi] = a.s; ids[hi] = a.id; }
      }
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < k; i += MERGE_THREADS) {
    const bool ok = ids[i] != INT64_MAX;
    out_scores[q * k + i] = ok ? __uint_as_float(ss[i]) : CUDART_INF_F;
    out_ids[q * k + i] = ok ? ids[i] : -1;
  }
}

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------
static size_t score_smem_bytes() { return (size_t)(STAGES * STAGE_FLOATS + TQ * DS_LD + 4) * 4 + STAGES * 8 + 16; }

static int plan(int64_t n_videos, int n_max, int dim, int& vt, int& n_tiles, int& nkc) {
  VFR_REQUIRE(n_videos > 0 && n_videos < (int64_t(1) << 31), VFR_ERR_INVALID, "n_videos=%lld out of range", (long long)n_videos);
  VFR_REQUIRE(n_max >= 1 && n_max <= VFR_MAX_SEG, VFR_ERR_UNSUPPORTED, "n_max=%d not in [1,%d]", n_max, VFR_MAX_SEG);
  VFR_REQUIRE(dim >= 1 && dim <= 8192, VFR_ERR_UNSUPPORTED, "dim=%d not in [1,8192]", dim);
  vt = TC / n_max;
  const int64_t nt = (n_videos + vt - 1) / vt;
  VFR_REQUIRE(nt < (int64_t(1) << 31), VFR_ERR_UNSUPPORTED, "too many bank tiles");
  n_tiles = (int)nt;
  nkc = nkc_of(dim);
  return VFR_OK;
}

template <int MODE, int NTAU>
static int launch_score(const ScoreParams& p, int n_split, cudaStream_t st) {
  const size_t smem = score_smem_bytes();
  VFR_CUDA(cudaFuncSetAttribute(score_kernel<MODE, NTAU>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int qtiles = (int)((p.n_queries + TQ - 1) / TQ);
  dim3 grid(qtiles, n_split);
  score_kernel<MODE, NTAU><<<grid, NTHREADS, smem, st>>>(p);
  return check_launch("score_kernel");
}

// Bank splits per query tile.  The kernel runs 2 CTAs per SM (launch bounds + 103 KB smem), so the
// grid is sized to fill whole waves of 2 x SMs CTAs: a grid of 1.08 waves costs as much as 2 waves.
static int wave_split(int64_t n_queries, int n_tiles, int max_split) {
  const int64_t qtiles = (n_queries + TQ - 1) / TQ;
  int sms = 148;
  int dev = 0;
  if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const int64_t slots = 2LL * sms;
  int best = 1;
  double best_util = 0.0;
  for (int w = 1; w <= 4; ++w) {
    int64_t ns = slots * w / qtiles;
    if (ns < 1) ns = 1;
    if (ns > n_tiles) ns = n_tiles;
    if (ns > max_split) ns = max_split;
    const int64_t ctas = ns * qtiles;
    const int64_t waves = (ctas + slots - 1) / slots;
    const double util = (double)ctas / (double)(waves * slots);
    if (util > best_util + 0.02) { best_util = util; best = (int)ns; }
  }
  return best;
}

static int auto_split(int64_t n_queries, int n_tiles, int n_split) {
  if (n_split > 0) return n_split < n_tiles ? n_split : n_tiles;
  return wave_split(n_queries, n_tiles, 64);
}

__global__ void fill_u32_kernel(unsigned* dst, unsigned value, size_t n) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) dst[i] = value;
}
int launch_fill_u32(unsigned* dst, unsigned value, size_t n, cudaStream_t st) {
  fill_u32_kernel<<<(unsigned)((n + 255) / 256 < 1024 ? (n + 255) / 256 : 1024), 256, 0, st>>>(dst, value, n);
  return check_launch("fill_u32_kernel");
}

int launch_topk_finish(const unsigned long long* cand, const int32_t* cand_cnt, int n_parts, int k, int64_t id_base,
                       int64_t n_queries, float* out_scores, int64_t* out_ids, cudaStream_t st) {
  topk_finish_kernel<<<(unsigned)n_queries, MERGE_THREADS, 0, st>>>(cand, cand_cnt, n_parts, k, id_base, out_scores,
                                                                   out_ids);
  return check_launch("topk_finish_kernel");
}

}  // namespace vfr

using namespace vfr;

extern "C" size_t vfr_bank_pack_bytes(int64_t n_videos, int n_max, int dim) {
  int vt, nt, nkc;
  if (plan(n_videos, n_max, dim, vt, nt, nkc) != VFR_OK) return 0;
  return ((size_t)nt * nkc * V_STAGE + (size_t)nt * TC) * sizeof(float);
}

extern "C" int vfr_bank_pack(const float* bank, const int32_t* vid_off, int64_t n_videos, int n_max, int dim,
                             float* packed, vfr_stream_t stream) {
  VFR_REQUIRE(bank && vid_off && packed, VFR_ERR_INVALID, "vfr_bank_pack: null pointer");
  int vt, nt, nkc;
  int rc = plan(n_videos, n_max, dim, vt, nt, nkc);
  if (rc) return rc;
  pack_bank_kernel<<<nt, 256, 0, (cudaStream_t)stream>>>(bank, vid_off, n_videos, vt, dim, nkc, packed,
                                                         packed + (size_t)nt * nkc * V_STAGE);
  return check_launch("pack_bank_kernel");
}

extern "C" size_t vfr_query_pack_bytes(int64_t n_queries, int dim) {
  if (n_queries <= 0 || dim <= 0) return 0;
  const size_t qt = (size_t)((n_queries + TQ - 1) / TQ);
  return (qt * nkc_of(dim) * Q_STAGE + qt * TQ) * sizeof(float);
}

extern "C" int vfr_query_pack(const float* queries, int64_t n_queries, int dim, float* packed, vfr_stream_t stream) {
  VFR_REQUIRE(queries && packed, VFR_ERR_INVALID, "vfr_query_pack: null pointer");
  VFR_REQUIRE(n_queries > 0 && dim > 0 && dim <= 8192, VFR_ERR_INVALID, "vfr_query_pack: bad shape");
  const int64_t qt = (n_queries + TQ - 1) / TQ;
  VFR_REQUIRE(qt < (int64_t(1) << 31), VFR_ERR_UNSUPPORTED, "too many query tiles");
  pack_query_kernel<<<(int)qt, 256, 0, (cudaStream_t)stream>>>(queries, n_queries, dim, nkc_of(dim), packed,
                                                               packed + (size_t)qt * nkc_of(dim) * Q_STAGE);
  return check_launch("pack_query_kernel");
}

static int fill_common(ScoreParams& p, const float* bank_packed, const int32_t* vid_off, const int64_t* mom_off,
                       int64_t n_videos, int n_max, int dim, const float* query_packed, int64_t n_queries) {
  VFR_REQUIRE(bank_packed && vid_off && query_packed, VFR_ERR_INVALID, "score: null pointer");
  VFR_REQUIRE(n_queries > 0, VFR_ERR_INVALID, "score: n_queries=%lld", (long long)n_queries);
  VFR_REQUIRE((n_queries + TQ - 1) / TQ <= 2147483647LL, VFR_ERR_UNSUPPORTED, "too many query tiles");
  int vt, nt, nkc;
  int rc = plan(n_videos, n_max, dim, vt, nt, nkc);
  if (rc) return rc;
  p = ScoreParams{};
  p.bank_packed = bank_packed;
  p.query_packed = query_packed;
  p.vid_off = vid_off;
  p.mom_off = mom_off;
  p.n_videos = n_videos;
  p.n_queries = n_queries;
  p.vt = vt;
  p.n_tiles = nt;
  p.nkc = nkc;
  p.bank_rowsum = bank_packed + (size_t)nt * nkc * V_STAGE;
  p.query_rowsum = query_packed + (size_t)((n_queries + TQ - 1) / TQ) * nkc * Q_STAGE;
  p.deps2 = (float)dim * VFR_PAIRWISE_EPS * VFR_PAIRWISE_EPS;
  return VFR_OK;
}

extern "C" int vfr_score_full(const float* bank_packed, const int32_t* vid_off, const int64_t* mom_off,
                              int64_t n_videos, int n_max, int dim, const float* query_packed, int64_t n_queries,
                              float* out, int64_t m_total, vfr_stream_t stream) {
  ScoreParams p;
  int rc = fill_common(p, bank_packed, vid_off, mom_off, n_videos, n_max, dim, query_packed, n_queries);
  if (rc) return rc;
  VFR_REQUIRE(out && mom_off && m_total > 0, VFR_ERR_INVALID, "vfr_score_full: bad output");
  const int ns = auto_split(n_queries, p.n_tiles, 0);
  p.tiles_per_split = (p.n_tiles + ns - 1) / ns;
  p.out_full = out;
  p.m_total = m_total;
  return launch_score<MODE_FULL, 1>(p, ns, (cudaStream_t)stream);
}

extern "C" int vfr_score_count(const float* bank_packed, const int32_t* vid_off, int64_t n_videos, int n_max,
                               int dim, const float* query_packed, int64_t n_queries, const float* tau, int n_tau,
                               const int32_t* q_video, uint32_t* cnt_lt, uint32_t* cnt_eqb, int n_split,
                               vfr_stream_t stream) {
  ScoreParams p;
  int rc = fill_common(p, bank_packed, vid_off, nullptr, n_videos, n_max, dim, query_packed, n_queries);
  if (rc) return rc;
  VFR_REQUIRE(tau && q_video && cnt_lt && cnt_eqb, VFR_ERR_INVALID, "vfr_score_count: null pointer");
  VFR_REQUIRE(n_tau >= 1 && n_tau <= VFR_MAX_TAU, VFR_ERR_UNSUPPORTED, "n_tau=%d not in [1,%d]", n_tau, VFR_MAX_TAU);
  const int ns = auto_split(n_queries, p.n_tiles, n_split);
  p.tiles_per_split = (p.n_tiles + ns - 1) / ns;
  p.tau = tau;
  p.n_tau = n_tau;
  p.q_video = q_video;
  p.cnt_lt = cnt_lt;
  p.cnt_eqb = cnt_eqb;
  cudaStream_t st = (cudaStream_t)stream;
  if (n_tau <= 2) return launch_score<MODE_COUNT, 2>(p, ns, st);
  if (n_tau <= 4) return launch_score<MODE_COUNT, 4>(p, ns, st);
  return launch_score<MODE_COUNT, VFR_MAX_TAU>(p, ns, st);
}

static int topk_split(int64_t n_queries, int n_split) {
  if (n_split > 0) return n_split;
  return wave_split(n_queries, 1 << 30, 32);
}

extern "C" size_t vfr_score_topk_bytes(int64_t n_queries, int n_split) {
  if (n_queries <= 0) return 0;
  const int ns = topk_split(n_queries, n_split);
  const size_t qpad = (size_t)((n_queries + TQ - 1) / TQ) * TQ;
  const size_t parts = (size_t)ns * 2;
  return qpad * parts * CAP * sizeof(unsigned long long) + qpad * parts * sizeof(int32_t) + qpad * sizeof(unsigned);
}

extern "C" int vfr_score_topk(const float* bank_packed, const int32_t* vid_off, const int64_t* mom_off,
                              int64_t n_videos, int n_max, int dim, const float* query_packed, int64_t n_queries,
                              int k, int64_t id_base, float* out_scores, int64_t* out_ids, void* workspace,
                              int n_split, vfr_stream_t stream) {
  ScoreParams p;
  int rc = fill_common(p, bank_packed, vid_off, mom_off, n_videos, n_max, dim, query_packed, n_queries);
  if (rc) return rc;
  VFR_REQUIRE(mom_off && out_scores && out_ids && workspace, VFR_ERR_INVALID, "vfr_score_topk: null pointer");
  VFR_REQUIRE(k >= 1 && k <= VFR_TOPK_MAX, VFR_ERR_UNSUPPORTED, "k=%d not in [1,%d]", k, VFR_TOPK_MAX);
  const int ns_req = topk_split(n_queries, n_split);   // workspace was sized with this many parts
  const int ns = ns_req < p.n_tiles ? ns_req : p.n_tiles;
  p.tiles_per_split = (p.n_tiles + ns - 1) / ns;
  const int ns_eff = (p.n_tiles + p.tiles_per_split - 1) / p.tiles_per_split;  // no empty CTAs
  const size_t qpad = (size_t)((n_queries + TQ - 1) / TQ) * TQ;
  p.k = k;
  p.n_parts = ns_eff * 2;
  p.cand = reinterpret_cast<unsigned long long*>(workspace);
  p.cand_cnt = reinterpret_cast<int32_t*>(p.cand + qpad * (size_t)ns_req * 2 * CAP);
  p.tau_g = reinterpret_cast<unsigned*>(p.cand_cnt + qpad * (size_t)ns_req * 2);
  cudaStream_t st = (cudaStream_t)stream;
  rc = launch_fill_u32(p.tau_g, 0x7f800000u, qpad, st);
  if (rc) return rc;
  rc = launch_score<MODE_TOPK, 1>(p, ns_eff, st);
  if (rc) return rc;
  return launch_topk_finish(p.cand, p.cand_cnt, p.n_parts, k, id_base, n_queries, out_scores, out_ids, st);
}

extern "C" int vfr_score_own(const float* bank, const int32_t* vid_off, int dim, const float* queries,
                             int64_t n_queries, const int32_t* q_video, float* out, int m_stride,
                             vfr_stream_t stream) {
  VFR_REQUIRE(bank && vid_off && queries && q_video && out, VFR_ERR_INVALID, "vfr_score_own: null pointer");
  VFR_REQUIRE(n_queries > 0 && dim > 0 && m_stride > 0, VFR_ERR_INVALID, "vfr_score_own: bad shape");
  dim3 block(VFR_MAX_SEG, 8);
  const int64_t blocks = (n_queries + 7) / 8;
  VFR_REQUIRE(blocks < (int64_t(1) << 31), VFR_ERR_UNSUPPORTED, "too many queries");
  score_own_kernel<<<(unsigned)blocks, block, 0, (cudaStream_t)stream>>>(bank, vid_off, dim, queries, n_queries,
                                                                         q_video, out, m_stride);
  return check_launch("score_own_kernel");
}

extern "C" int vfr_topk_merge(const float* in_scores, const int64_t* in_ids, int n_parts, int64_t n_queries, int k,
                              float* out_scores, int64_t* out_ids, vfr_stream_t stream) {
  VFR_REQUIRE(in_scores && in_ids && out_scores && out_ids, VFR_ERR_INVALID, "vfr_topk_merge: null pointer");
  VFR_REQUIRE(n_parts >= 1 && k >= 1 && n_queries > 0, VFR_ERR_INVALID, "vfr_topk_merge: bad shape");
  const int total = n_parts * k;
  VFR_REQUIRE(total <= 4096, VFR_ERR_UNSUPPORTED, "vfr_topk_merge: n_parts*k=%d > 4096", total);
  int n_pad = 2;
  while (n_pad < total) n_pad <<= 1;
  const size_t smem = (size_t)n_pad * 16;
  VFR_CUDA(cudaFuncSetAttribute(topk_merge_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  topk_merge_kernel<<<(unsigned)n_queries, MERGE_THREADS, smem, (cudaStream_t)stream>>>(
      in_scores, in_ids, n_parts, n_queries * k, n_queries * k, k, n_pad, out_scores, out_ids, nullptr, 0, nullptr, nullptr);
  return check_launch("topk_merge_kernel");
}

// K7 on the records of the query-slice exchange: `blocks` holds n_parts records (one per shard, as written by
// vfr_sel_refine_blocks: {ids int64 [per, k] | scores fp32 [per, k] | flags int32 [per]}, vfr_topk_block_bytes(per, k)
// bytes each) for the SAME slice of `per` queries; the first n_rows of them are merged.
extern "C" int vfr_topk_merge_blocks(const void* blocks, int n_parts, int64_t per, int64_t n_rows, int k, float* out_scores,
                                     int64_t* out_ids, int32_t* out_flags, int64_t* n_flagged, vfr_stream_t stream) {
  VFR_REQUIRE(blocks && out_scores && out_ids && out_flags, VFR_ERR_INVALID, "vfr_topk_merge_blocks: null pointer");
  VFR_REQUIRE(n_parts >= 1 && k >= 1 && per > 0 && n_rows >= 0 && n_rows <= per, VFR_ERR_INVALID, "vfr_topk_merge_blocks: bad shape");
  const int total = n_parts * k;
  VFR_REQUIRE(total <= 4096, VFR_ERR_UNSUPPORTED, "vfr_topk_merge_blocks: n_parts*k=%d > 4096", total);
  const int64_t blk = per * ((int64_t)k * 12 + 4);
  VFR_REQUIRE(blk % 8 == 0, VFR_ERR_INVALID, "vfr_topk_merge_blocks: per * (3k + 1) must be even");
  if (n_rows == 0) return VFR_OK;
  int n_pad = 2;
  while (n_pad < total) n_pad <<= 1;
  const size_t smem = (size_t)n_pad * 16;
  const unsigned char* base = reinterpret_cast<const unsigned char*>(blocks);
  VFR_CUDA(cudaFuncSetAttribute(topk_merge_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  topk_merge_kernel<<<(unsigned)n_rows, MERGE_THREADS, smem, (cudaStream_t)stream>>>(
      reinterpret_cast<const float*>(base + per * k * 8), reinterpret_cast<const int64_t*>(base), n_parts, blk / 4, blk / 8, k,
      n_pad, out_scores, out_ids, reinterpret_cast<const int32_t*>(base + per * k * 12), blk / 4, out_flags,
      reinterpret_cast<unsigned long long*>(n_flagged));
  return check_launch("topk_merge_kernel (blocks)");
}
