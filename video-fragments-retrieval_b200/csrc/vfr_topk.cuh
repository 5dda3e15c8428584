// Thread-private top-k candidate lists with warp-cooperative compaction (shared by the exact-fp32
// and the tcgen05 scoring kernels) + the per-query merge of the part lists.
#pragma once
#include "vfr_common.cuh"
#include <math_constants.h>

namespace vfr {

constexpr int CAP = VFR_TOPK_CAP;   // candidate slots per (query, part) list

// ---------------------------------------------------------------------------------------------
// warp-cooperative compaction of one thread-private candidate list (top-k mode)
//
// A list holds up to CAP unsorted 64-bit keys (score bits << 32 | moment id; scores are non-negative
// fp32, so unsigned order == score order, ties broken by id).  When it fills up, the whole warp
// SELECTS the k smallest keys (no sort): a most-significant-bit-first radix select over the bits in
// which the candidates actually differ - first on the score word, then (only for exact score ties)
// on the id word - followed by a ballot/popc stream compaction.  ~1.5 k warp instructions per 512
// keys, i.e. a handful per discarded candidate.  Final ordering is left to topk_finish_kernel.
// ---------------------------------------------------------------------------------------------
constexpr int TK_SLOTS = CAP / 32;   // keys per lane

__device__ __forceinline__ unsigned warp_and(unsigned v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v &= __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ unsigned warp_or(unsigned v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v |= __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ int warp_sum_int(int v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// kk-th smallest (1-based) value of word w among the candidates (bit s of candmask = slot s of this
// lane is a candidate; m = number of candidates in the warp).  On return candmask / m describe the
// candidates whose word equals the returned value and kk is the rank still to resolve among them.
__device__ __forceinline__ unsigned radix_kth(const unsigned (&w)[TK_SLOTS], unsigned& candmask, int& kk, int& m) {
  unsigned a = 0xffffffffu, o = 0u;
#pragma unroll
  for (int s = 0; s < TK_SLOTS; ++s)
    if ((candmask >> s) & 1u) { a &= w[s]; o |= w[s]; }
  a = warp_and(a);
  o = warp_or(o);
  const unsigned diff = a ^ o;
  if (diff == 0u) return a;                      // all candidates agree on this word
  const int top = 31 - __clz(diff);
  unsigned value = a & ~((2u << top) - 1u);      // bits above `top` are common to every candidate
  int bit = top;
  for (; bit >= 0 && m > 1; --bit) {
    unsigned ones = 0u;
#pragma unroll
    for (int s = 0; s < TK_SLOTS; ++s) ones |= ((w[s] >> bit) & 1u) << s;
    const int c0 = warp_sum_int(__popc(candmask & ~ones));
    if (kk <= c0) { m = c0; candmask &= ~ones; }
    else { kk -= c0; m -= c0; candmask &= ones; value |= 1u << bit; }
  }
  if (bit >= 0) {
    // a single candidate is left before all bits were decided: take its word
    unsigned mine = 0u;
#pragma unroll
    for (int s = 0; s < TK_SLOTS; ++s)
      if ((candmask >> s) & 1u) mine = w[s];
    const unsigned who = __ballot_sync(0xffffffffu, candmask != 0u);
    value = __shfl_sync(0xffffffffu, mine, __ffs(who) - 1);
  }
  return value;
}

// Every lane calls this (warp-uniform call site).  For each lane whose `need` is set the whole warp
// reduces that lane's list to its k smallest keys (unsorted) and tightens that lane's tau.
__device__ __forceinline__ void compact_lists(unsigned long long* list, int& cnt, float& tau, int k,
                                              bool need, int lane) {
  unsigned mask = __ballot_sync(0xffffffffu, need);
  while (mask) {
    const int src = __ffs(mask) - 1;
    mask &= mask - 1;
    unsigned long long* lp =
        reinterpret_cast<unsigned long long*>(__shfl_sync(0xffffffffu, reinterpret_cast<unsigned long long>(list), src));
    const int n = __shfl_sync(0xffffffffu, cnt, src);
    if (n <= k) continue;                        // nothing to drop (warp-uniform)
    __syncwarp();
    unsigned hi[TK_SLOTS], lo[TK_SLOTS];
    unsigned candmask = 0u;
#pragma unroll
    for (int s = 0; s < TK_SLOTS; ++s) {
      const int idx = (s << 5) | lane;
      unsigned long long key = ~0ull;
      if (idx < n) { key = lp[idx]; candmask |= 1u << s; }
      hi[s] = (unsigned)(key >> 32);
      lo[s] = (unsigned)key;
    }
    int kk = k, m = n;
    const unsigned kth_hi = radix_kth(hi, candmask, kk, m);
    const unsigned kth_lo = radix_kth(lo, candmask, kk, m);
    __syncwarp();
    int base = 0;
#pragma unroll
    for (int s = 0; s < TK_SLOTS; ++s) {
      const int idx = (s << 5) | lane;
      const bool keep = idx < n && (hi[s] < kth_hi || (hi[s] == kth_hi && lo[s] <= kth_lo));
      const unsigned bal = __ballot_sync(0xffffffffu, keep);
      if (keep) lp[base + __popc(bal & ((1u << lane) - 1u))] = ((unsigned long long)hi[s] << 32) | lo[s];
      base += __popc(bal);
    }
    __syncwarp();
    if (lane == src) {
      cnt = base;                                // == k (keys are unique)
      tau = __uint_as_float(kth_hi);
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
