// Thread-private top-k candidate lists with warp-cooperative compaction (shared by the exact-fp32
// and the tcgen05 scoring kernels) + the per-query merge of the part lists.
#pragma once
#include "vfr_common.cuh"
#include <math_constants.h>

namespace vfr {

constexpr int CAP = VFR_TOPK_CAP;   // 256 candidate slots per (query, part) list

// ---------------------------------------------------------------------------------------------
// warp-cooperative compaction of one thread-private candidate list (top-k mode)
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void cswap(unsigned long long& a, unsigned long long& b, bool asc) {
  const bool sw = (a > b) == asc;
  const unsigned long long t = a;
  a = sw ? b : a;
  b = sw ? t : b;
}

// bitonic sort of 256 keys held 8 per lane, element index e = slot*32 + lane, ascending in e
__device__ __forceinline__ void warp_sort256(unsigned long long (&key)[8], int lane) {
#pragma unroll
  for (int size = 2; size <= 256; size <<= 1) {
#pragma unroll
    for (int stride = size >> 1; stride > 0; stride >>= 1) {
      if (stride >= 32) {
        const int ss = stride >> 5;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          if (!(i & ss)) {
            const bool asc = (size == 256) ? true : !((i << 5) & size);
            cswap(key[i], key[i | ss], asc);
          }
        }
      } else {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const unsigned long long other = __shfl_xor_sync(0xffffffffu, key[i], stride);
          const int e = (i << 5) | lane;
          const bool asc = (size == 256) ? true : !(e & size);
          const bool lower = !(lane & stride);
          const bool take_min = (lower == asc);
          const unsigned long long mn = key[i] < other ? key[i] : other;
          const unsigned long long mx = key[i] < other ? other : key[i];
          key[i] = take_min ? mn : mx;
        }
      }
    }
  }
}

// Every lane calls this (warp-uniform call site).  For each lane whose `need` is set the whole warp
// sorts that lane's list and keeps the k smallest (score, id) keys, sorted, at its front.
__device__ __forceinline__ void compact_lists(unsigned long long* list, int& cnt, float& tau, int k,
                                              bool need, int lane) {
  unsigned mask = __ballot_sync(0xffffffffu, need);
  while (mask) {
    const int src = __ffs(mask) - 1;
    mask &= mask - 1;
    unsigned long long* lp =
        reinterpret_cast<unsigned long long*>(__shfl_sync(0xffffffffu, reinterpret_cast<unsigned long long>(list), src));
    const int n = __shfl_sync(0xffffffffu, cnt, src);
    __syncwarp();
    unsigned long long key[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int idx = (i << 5) | lane;
      key[i] = idx < n ? lp[idx] : ~0ull;
    }
    warp_sort256(key, lane);
    const int keep = n < k ? n : k;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int idx = (i << 5) | lane;
      if (idx < keep) lp[idx] = key[i];
    }
    // k-th smallest (rank k-1) lives in slot (k-1)>>5 of lane (k-1)&31
    unsigned long long kth = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const unsigned long long cand = __shfl_sync(0xffffffffu, key[i], (k - 1) & 31);
      if (i == ((k - 1) >> 5)) kth = cand;
    }
    __syncwarp();
    if (lane == src) {
      cnt = keep;
      tau = (n >= k) ? __uint_as_float((unsigned)(kth >> 32)) : CUDART_INF_F;
    }
  }
}


// Per-query threshold shared by all part lists of a query (non-negative fp32 compared as uint32).
// Each list's k-th best score is an upper bound of the GLOBAL k-th best, so publishing it with an
// atomic min lets every other list filter with it: appends per query drop from
// parts * k ln(N/(parts k)) to ~k ln(N/k), which keeps the append branch rare per WARP as well.
__device__ __forceinline__ void tau_publish(unsigned* tau_g, float tau) {
  if (tau < CUDART_INF_F) atomicMin(tau_g, __float_as_uint(tau));
}
__device__ __forceinline__ float tau_fetch(const unsigned* tau_g) {
  return __uint_as_float(*reinterpret_cast<const volatile unsigned*>(tau_g));
}
int launch_fill_u32(unsigned* dst, unsigned value, size_t n, cudaStream_t st);

// merges the part lists of every query: out [Q, k] ascending by (score, id); defined in vfr_score.cu
int launch_topk_finish(const unsigned long long* cand, const int32_t* cand_cnt, int n_parts, int k, int64_t id_base,
                       int64_t n_queries, float* out_scores, int64_t* out_ids, cudaStream_t st);

}  // namespace vfr
