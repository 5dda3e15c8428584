// fp32 CUDA-core GEMM  C[M,N] = A[M,K] * B[N,K]^T  with a fused epilogue functor.
// The exact-fp32 building block of K2 (visual MLP) and K3 (BiLSTM): 128x128x16 tiles, 256 threads,
// 8x8 register blocks, register-staged double buffering.  Both operands are K-contiguous
// ("NT"), which is how nn.Linear / nn.LSTM store their weights, so no transposes are needed.
#pragma once
#include "vfr_common.cuh"

namespace vfr {

constexpr int G_BM = 128, G_BN = 128, G_BK = 16, G_THREADS = 256, G_PAD = 4;

// batch entry (blockIdx.z) of a GEMM launch
struct GemmOperand {
  const float* A;  // [M, K], row stride lda
  const float* B;  // [N, K], row stride ldb
};
struct GemmBatch {   // passed by value as a kernel parameter
  GemmOperand op[2];
};

template <int VEC>
__device__ __forceinline__ void load_vec(const float* p, bool ok, float (&dst)[VEC]) {
  if (!ok) {
#pragma unroll
    for (int i = 0; i < VEC; ++i) dst[i] = 0.f;
    return;
  }
  if (VEC == 4) {
    const float4 v = __ldg(reinterpret_cast<const float4*>(p));
    dst[0] = v.x; dst[1] = v.y; dst[2] = v.z; dst[3] = v.w;
  } else if (VEC == 2) {
    const float2 v = __ldg(reinterpret_cast<const float2*>(p));
    dst[0] = v.x; dst[1] = v.y;
  } else {
    dst[0] = __ldg(p);
  }
}

// Epi: struct with  __device__ void operator()(int z, int m, int n0, const float (&v)[4]) const
//      called for 4 consecutive columns n0..n0+3 of row m (bounds: m < M checked by the caller,
//      columns must be checked by the functor against N).
template <int VEC, class Epi>
__global__ void __launch_bounds__(G_THREADS, 2)
sgemm_nt_kernel(const GemmBatch ops, int lda, int ldb, int M, int N, int K, Epi epi) {
  __shared__ __align__(16) float As[2][G_BK][G_BM + G_PAD];
  __shared__ __align__(16) float Bs[2][G_BK][G_BN + G_PAD];
  const int z = blockIdx.z;
  const float* __restrict__ A = ops.op[z].A;
  const float* __restrict__ B = ops.op[z].B;
  const int tid = threadIdx.x;
  const int m0 = blockIdx.y * G_BM, n0 = blockIdx.x * G_BN;
  constexpr int TPR = G_BK / VEC;            // threads per tile row
  constexpr int RPP = G_THREADS / TPR;       // rows per pass
  constexpr int NV = G_BM / RPP;             // vectors per thread per operand
  const int lrow = tid / TPR, lk = (tid % TPR) * VEC;
  float ra[NV][VEC], rb[NV][VEC];

  auto gload = [&](int k0) {
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int r = lrow + i * RPP;
      const int k = k0 + lk;
      load_vec<VEC>(A + (int64_t)(m0 + r) * lda + k, (m0 + r) < M && k < K, ra[i]);
      load_vec<VEC>(B + (int64_t)(n0 + r) * ldb + k, (n0 + r) < N && k < K, rb[i]);
    }
  };
  auto sstore = [&](int buf) {
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int r = lrow + i * RPP;
#pragma unroll
      for (int v = 0; v < VEC; ++v) {
        As[buf][lk + v][r] = ra[i][v];
        Bs[buf][lk + v][r] = rb[i][v];
      }
    }
  };

  const int ty = tid >> 4, tx = tid & 15;
  float acc[8][8];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;

  const int nk = (K + G_BK - 1) / G_BK;
  gload(0);
  sstore(0);
  __syncthreads();
  for (int kt = 0; kt < nk; ++kt) {
    const int buf = kt & 1;
    if (kt + 1 < nk) gload((kt + 1) * G_BK);
#pragma unroll
    for (int kk = 0; kk < G_BK; ++kk) {
      const float4 a0 = *reinterpret_cast<const float4*>(&As[buf][kk][ty * 4]);
      const float4 a1 = *reinterpret_cast<const float4*>(&As[buf][kk][64 + ty * 4]);
      const float4 b0 = *reinterpret_cast<const float4*>(&Bs[buf][kk][tx * 4]);
      const float4 b1 = *reinterpret_cast<const float4*>(&Bs[buf][kk][64 + tx * 4]);
      const float av[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
      const float bv[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = __fmaf_rn(av[i], bv[j], acc[i][j]);
    }
    if (kt + 1 < nk) {
      sstore(buf ^ 1);
      __syncthreads();
    }
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int m = m0 + ((i < 4) ? (ty * 4 + i) : (64 + ty * 4 + i - 4));
    if (m >= M) continue;
    const float v0[4] = {acc[i][0], acc[i][1], acc[i][2], acc[i][3]};
    const float v1[4] = {acc[i][4], acc[i][5], acc[i][6], acc[i][7]};
    epi(z, m, n0 + tx * 4, v0);
    epi(z, m, n0 + 64 + tx * 4, v1);
  }
}

// host launcher: picks the widest vector load the pointers / strides allow
template <class Epi>
static int launch_sgemm_nt(const GemmBatch& ops, int batch, int lda, int ldb, int M, int N, int K, Epi epi,
                           cudaStream_t st) {
  if (M <= 0 || N <= 0 || K <= 0) return VFR_OK;
  int vec = 4;
  for (int z = 0; z < batch; ++z) {
    const uintptr_t a = reinterpret_cast<uintptr_t>(ops.op[z].A), b = reinterpret_cast<uintptr_t>(ops.op[z].B);
    while (vec > 1 && ((a | b) % (vec * 4) != 0)) vec >>= 1;
  }
  while (vec > 1 && (lda % vec != 0 || ldb % vec != 0 || K % vec != 0)) vec >>= 1;
  dim3 grid((N + G_BN - 1) / G_BN, (M + G_BM - 1) / G_BM, batch);
  if (vec == 4) sgemm_nt_kernel<4, Epi><<<grid, G_THREADS, 0, st>>>(ops, lda, ldb, M, N, K, epi);
  else if (vec == 2) sgemm_nt_kernel<2, Epi><<<grid, G_THREADS, 0, st>>>(ops, lda, ldb, M, N, K, epi);
  else sgemm_nt_kernel<1, Epi><<<grid, G_THREADS, 0, st>>>(ops, lda, ldb, M, N, K, epi);
  return check_launch("sgemm_nt_kernel");
}

}  // namespace vfr
