// Shared device/host helpers for libvfr (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include "../../include/vfr.h"

#define VFR_PAIRWISE_EPS 1e-6f   // F.pairwise_distance eps, inside the norm (reference model/evaluate.py:53)
#define VFR_NORM_EPS 1e-5f       // x / (|x| + 1e-5)  (reference model/data.py:177, model/main.py:220)
#define VFR_MAX_SEG 32           // device cap on clips per video

namespace vfr {

void set_error(const char* fmt, ...);
int check_launch(const char* what);

#define VFR_REQUIRE(cond, code, ...)      \
  do {                                    \
    if (!(cond)) {                        \
      vfr::set_error(__VA_ARGS__);        \
      return (code);                      \
    }                                     \
  } while (0)

#define VFR_CUDA(call)                                                                  \
  do {                                                                                  \
    cudaError_t e_ = (call);                                                            \
    if (e_ != cudaSuccess) {                                                            \
      vfr::set_error("%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); \
      return VFR_ERR_CUDA;                                                              \
    }                                                                                   \
  } while (0)

__host__ __device__ __forceinline__ int num_moments(int n) { return n * (n + 1) / 2; }

// index of the inclusive clip range (s, e) in the reference's enumeration (model/utils.py:71-75):
// the n single clips first, then all s<e pairs in lexicographic order.
__host__ __device__ __forceinline__ int moment_index(int n, int s, int e) {
  return (s == e) ? s : n + s * (n - 1) - (s * (s - 1)) / 2 + (e - s - 1);
}

// inverse map m -> (s, e)
__host__ __device__ __forceinline__ void moment_se(int n, int m, int& s, int& e) {
  if (m < n) { s = m; e = m; return; }
  int r = m - n;
  s = 0;
  while (r >= n - 1 - s) { r -= n - 1 - s; ++s; }
  e = s + 1 + r;
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// ---- mbarrier + bulk async copy (UBLKCP) helpers -------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void fence_barrier_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_LOOP:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra DONE;\n"
      "bra WAIT_LOOP;\n"
      "DONE:\n"
      "}\n" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}
// 1-D bulk copy global -> shared, completion on an mbarrier (bytes % 16 == 0, 16-B aligned)
__device__ __forceinline__ void bulk_g2s(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(dst_smem)),
               "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

}  // namespace vfr
