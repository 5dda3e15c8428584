// K5: integer-exact temporal IoU, ground-truth labelling, ranking of one video's moments and the
// per-query rank statistics the two evaluation protocols consume.
//
// Replaces reference model/utils.py:78-82 (get_iou), model/evaluate.py:59-65 (gt rule),
// model/evaluate_single.py:52-54 (argsort + reversal), :58-73 (ranks, top-1 IoU, recall bits).
// IoU is kept as the integer pair (intersection, union); "iou > thr" is a lookup in a host-built
// table T[inter][union] evaluated in float64 exactly as NumPy does, so labels are bit-identical
// for any python-float threshold.  Final float64 means/medians stay on the host (NumPy).
#include "vfr_common.cuh"
#include <math_constants.h>

namespace vfr {

constexpr int TAB = 65;  // table side: inter, union in [0, 64]

__device__ __forceinline__ void iou_int(int s, int e, int ts, int te, int& inter, int& uni) {
  inter = max(min(te, e) + 1 - max(ts, s), 0);
  uni = max(te, e) + 1 - min(ts, s);
}

// one warp per query
__global__ void gt_select_kernel(const float* __restrict__ own, int m_stride, const int32_t* __restrict__ q_nseg,
                                 const int32_t* __restrict__ times, int n_annot, const uint8_t* __restrict__ tables,
                                 int n_thr, int64_t n_queries, uint8_t* __restrict__ gt, float* __restrict__ tau,
                                 int32_t* __restrict__ pos, int32_t* __restrict__ npos,
                                 int32_t* __restrict__ eq_before) {
  const int lane = threadIdx.x & 31;
  const int64_t q = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (q >= n_queries) return;
  const int n = q_nseg[q];
  const int M = num_moments(n);
  const int32_t* tq = times + q * n_annot * 2;
  for (int t = 0; t < n_thr; ++t) {
    const uint8_t* tab = tables + (size_t)t * TAB * TAB;
    unsigned long long best = ~0ull;  // (score bits, m)
    int count = 0;
    for (int m0 = 0; m0 < m_stride; m0 += 32) {
      const int m = m0 + lane;
      int positive = 0;
      if (m < M) {
        int s, e;
        moment_se(n, m, s, e);
        int hits = 0;
        for (int a = 0; a < n_annot; ++a) {
          const int ts = tq[2 * a], te = tq[2 * a + 1];
          if (ts < 0) continue;
          int inter, uni;
          iou_int(s, e, ts, te, inter, uni);
          hits += (inter < TAB && uni < TAB) ? tab[inter * TAB + uni] : 0;
        }
        positive = hits >= 2;
        if (positive) {
          const unsigned long long key = ((unsigned long long)__float_as_uint(own[q * m_stride + m]) << 32) | (unsigned)m;
          best = key < best ? key : best;
        }
      }
      if (m < m_stride) gt[(q * n_thr + t) * m_stride + m] = (uint8_t)positive;
      count += __popc(__ballot_sync(0xffffffffu, positive));
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const unsigned long long other = __shfl_xor_sync(0xffffffffu, best, o);
      best = other < best ? other : best;
    }
    const bool any = best != ~0ull;
    const int bpos = any ? (int)(best & 0xffffffffu) : -1;
    const float btau = any ? __uint_as_float((unsigned)(best >> 32)) : CUDART_INF_F;
    // own-video moments that tie with tau and precede pos in index order (deterministic tie rank)
    int eqb = 0;
    for (int m0 = 0; m0 < bpos; m0 += 32) {
      const int m = m0 + lane;
      eqb += __popc(__ballot_sync(0xffffffffu, m < bpos && own[q * m_stride + m] == btau));
    }
    if (lane == 0) {
      tau[q * n_thr + t] = btau;
      pos[q * n_thr + t] = bpos;
      npos[q * n_thr + t] = count;
      eq_before[q * n_thr + t] = eqb;
    }
  }
}

// ranking of one video's moments: block per query, bitonic sort of (score, m) in shared memory
constexpr int RANK_N = 1024;
__global__ void __launch_bounds__(256) rank_order_kernel(const float* __restrict__ own, int m_stride,
                                                         const int32_t* __restrict__ q_nseg, int descending,
                                                         int32_t* __restrict__ order) {
  __shared__ unsigned long long keys[RANK_N];
  const int64_t q = blockIdx.x;
  const int M = num_moments(q_nseg[q]);
  for (int i = threadIdx.x; i < RANK_N; i += blockDim.x)
    keys[i] = (i < M) ? (((unsigned long long)__float_as_uint(own[q * m_stride + i]) << 32) | (unsigned)i) : ~0ull;
  for (int size = 2; size <= RANK_N; size <<= 1) {
    for (int stride = size >> 1; stride > 0; stride >>= 1) {
      __syncthreads();
      for (int i = threadIdx.x; i < RANK_N / 2; i += blockDim.x) {
        const int lo = ((i & ~(stride - 1)) << 1) | (i & (stride - 1));
        const int hi = lo | stride;
        const bool asc = !(lo & size) || size == RANK_N;
        const unsigned long long a = keys[lo], b = keys[hi];
        if ((a > b) == asc) { keys[lo] = b; keys[hi] = a; }
      }
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < m_stride; i += blockDim.x) {
    int val = -1;
    if (i < M) val = (int)(keys[descending ? (M - 1 - i) : i] & 0xffffffffu);
    order[q * m_stride + i] = val;
  }
}

// per-query rank statistics for a given ranked list of moment indices; one warp per query
__global__ void single_metrics_kernel(const int32_t* __restrict__ order, int m_stride,
                                      const int32_t* __restrict__ q_nseg, const int32_t* __restrict__ times,
                                      int n_annot, const uint8_t* __restrict__ tables, int n_thr, int64_t n_queries,
                                      int32_t* __restrict__ ranks, int32_t* __restrict__ top1_inter,
                                      int32_t* __restrict__ top1_union, int32_t* __restrict__ first_pos) {
  const int lane = threadIdx.x & 31;
  const int64_t q = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (q >= n_queries) return;
  const int n = q_nseg[q];
  const int M = num_moments(n);
  const int32_t* tq = times + q * n_annot * 2;
  const int32_t* ord = order + q * m_stride;
  int s1, e1;
  moment_se(n, ord[0], s1, e1);
  for (int a = lane; a < n_annot; a += 32) {
    const int ts = tq[2 * a], te = tq[2 * a + 1];
    int rank = 0, inter = 0, uni = 0;
    if (ts >= 0) {
      rank = -1;  // an annotated time that is not a candidate moment: the reference raises ValueError
      if (ts <= te && te < n) {
        const int target = moment_index(n, ts, te);
        for (int i = 0; i < M; ++i)
          if (ord[i] == target) { rank = i + 1; break; }
      }
      // get_iou([predicts[0]], time[0], time[1]): the top-1 moment plays "times", the annotation (s, e)
      iou_int(ts, te, s1, e1, inter, uni);
    }
    ranks[q * n_annot + a] = rank;
    top1_inter[q * n_annot + a] = inter;
    top1_union[q * n_annot + a] = uni;
  }
  for (int t = 0; t < n_thr; ++t) {
    const uint8_t* tab = tables + (size_t)t * TAB * TAB;
    int first = m_stride;
    for (int i0 = 0; i0 < M && first == m_stride; i0 += 32) {
      const int i = i0 + lane;
      int positive = 0;
      if (i < M) {
        int s, e;
        moment_se(n, ord[i], s, e);
        int hits = 0;
        for (int a = 0; a < n_annot; ++a) {
          const int ts = tq[2 * a], te = tq[2 * a + 1];
          if (ts < 0) continue;
          int inter, uni;
          iou_int(s, e, ts, te, inter, uni);
          hits += (inter < TAB && uni < TAB) ? tab[inter * TAB + uni] : 0;
        }
        positive = hits >= 2;
      }
      const unsigned b = __ballot_sync(0xffffffffu, positive);
      if (b) first = i0 + __ffs(b) - 1;
    }
    if (lane == 0) first_pos[q * n_thr + t] = first;
  }
}

}  // namespace vfr

using namespace vfr;

extern "C" int vfr_gt_select(const float* own_scores, int m_stride, const int32_t* q_nseg, const int32_t* times,
                             int n_annot, const uint8_t* tables, int n_thr, int64_t n_queries, uint8_t* gt,
                             float* tau, int32_t* pos, int32_t* npos, int32_t* eq_before, vfr_stream_t stream) {
  VFR_REQUIRE(own_scores && q_nseg && times && tables && gt && tau && pos && npos && eq_before, VFR_ERR_INVALID,
              "vfr_gt_select: null pointer");
  VFR_REQUIRE(n_queries > 0 && m_stride > 0 && n_annot > 0 && n_thr > 0, VFR_ERR_INVALID, "vfr_gt_select: bad shape");
  const int64_t blocks = (n_queries + 7) / 8;
  VFR_REQUIRE(blocks < (int64_t(1) << 31), VFR_ERR_UNSUPPORTED, "too many queries");
  gt_select_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(own_scores, m_stride, q_nseg, times, n_annot,
                                                                       tables, n_thr, n_queries, gt, tau, pos, npos, eq_before);
  return check_launch("gt_select_kernel");
}

extern "C" int vfr_rank_order(const float* own_scores, int m_stride, const int32_t* q_nseg, int64_t n_queries,
                              int descending, int32_t* order, vfr_stream_t stream) {
  VFR_REQUIRE(own_scores && q_nseg && order, VFR_ERR_INVALID, "vfr_rank_order: null pointer");
  VFR_REQUIRE(n_queries > 0 && n_queries < (int64_t(1) << 31) && m_stride > 0 && m_stride <= RANK_N, VFR_ERR_UNSUPPORTED,
              "vfr_rank_order: m_stride=%d (max %d)", m_stride, RANK_N);
  rank_order_kernel<<<(unsigned)n_queries, 256, 0, (cudaStream_t)stream>>>(own_scores, m_stride, q_nseg, descending,
                                                                          order);
  return check_launch("rank_order_kernel");
}

extern "C" int vfr_single_metrics(const int32_t* order, int m_stride, const int32_t* q_nseg, const int32_t* times,
                                  int n_annot, const uint8_t* tables, int n_thr, int64_t n_queries, int32_t* ranks,
                                  int32_t* top1_inter, int32_t* top1_union, int32_t* first_pos, vfr_stream_t stream) {
  VFR_REQUIRE(order && q_nseg && times && tables && ranks && top1_inter && top1_union && first_pos, VFR_ERR_INVALID,
              "vfr_single_metrics: null pointer");
  VFR_REQUIRE(n_queries > 0 && m_stride > 0 && n_annot > 0 && n_thr > 0, VFR_ERR_INVALID,
              "vfr_single_metrics: bad shape");
  const int64_t blocks = (n_queries + 7) / 8;
  VFR_REQUIRE(blocks < (int64_t(1) << 31), VFR_ERR_UNSUPPORTED, "too many queries");
  single_metrics_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(
      order, m_stride, q_nseg, times, n_annot, tables, n_thr, n_queries, ranks, top1_inter, top1_union, first_pos);
  return check_launch("single_metrics_kernel");
}
