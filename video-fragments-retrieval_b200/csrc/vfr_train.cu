// Training step: hand-written forward-with-saved-activations and BACKWARD of the two embedding
// branches, a fused multi-tensor Adam and the gradient-norm statistic - what replaces the autograd graph / cuBLAS / cuDNN
// calls behind the reference's  loss.backward(); optimizer.step()  (model/main.py:57-67, optimiser main.py:358,
// grad-norm logging utils.py:85-92).  The ranking loss itself is K6 (vfr_loss.cu).
//
// Shapes of a training step are tiny (R ~ 120 clip rows x 3 streams, B ~ 87 queries, 20 tokens, H = 1000), so the step is
// bound by launch latency and by the 20 strictly sequential recurrent steps, not by FLOP/s: everything here is fp32 on the
// CUDA cores (bit-level behaviour of the fp32 reference up to summation order), organised to keep all SMs busy on skinny
// problems - a strided 64 x 64 x 16 SGEMM with deterministic split-K (partials are summed by their consumer), both LSTM
// directions in one launch, the weight gradients of all 20 steps in ONE GEMM per matrix (K = T * B).
//
// LSTM bookkeeping: all per-step tensors are indexed by ORIGINAL time t (rows t * B + b).  The forward direction walks
// t = 0 .. T-1 (previous state: t - 1), the backward direction t = T-1 .. 0 (previous state: t + 1); the padding is fed
// through the recurrence as the reference does (models.py:65, no packing).  Gate order i, f, g, o (PyTorch).
#include "vfr_common.cuh"
#include <algorithm>

namespace vfr {

// ---------------------------------------------------------------------------------------------------
// strided SGEMM:  C[d][z][m][n] = sum_{k in split z} A[d](m, k) * B[d](k, n)  (+ bias[n])  (+ C if accumulate)
// ---------------------------------------------------------------------------------------------------
struct TgArgs {
  const float* A; const float* B; float* C;
  int M, N, K;
  int64_t sam, sak;        // A(m, k) = A[m * sam + k * sak]
  int64_t sbk, sbn;        // B(k, n) = B[k * sbk + n * sbn]
  int64_t ldc;
  int dirs;                // independent problems along blockIdx.z (the two LSTM directions)
  int64_t a_dir, b_dir, c_dir;
  int ksplit;              // split-K parts along blockIdx.z; part z is written at C + z * c_split (NOT summed here)
  int64_t c_split;
  const float* bias;       // optional [dirs][N] (stride bias_dir), added to split 0
  const float* bias2;      // optional second bias vector (the LSTM keeps b_ih and b_hh apart)
  int64_t bias_dir;
  int accumulate;          // C += instead of C =
  const float* Ad[2];      // optional per-direction operand pointers (override A + d * a_dir etc.): the two LSTM directions
  const float* Bd[2];      //   are separate allocations, one launch serves both
  float* Cd[2];
};
constexpr int TG_T = 64, TG_K = 16;

template <bool A_KC, bool B_NC>
__global__ void __launch_bounds__(256) tg_gemm_kernel(const TgArgs a) {
  // two shared-memory buffers + register prefetch: the global loads of tile i + 1 are in flight while tile i is multiplied
  // (these problems are skinny - M ~ 87 - so a CTA's k loop is a latency chain, not a bandwidth stream)
  __shared__ float As[2][TG_K][TG_T + 4];
  __shared__ float Bs[2][TG_K][TG_T + 4];
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  const int d = blockIdx.z / a.ksplit, z = blockIdx.z % a.ksplit;
  const int m0 = blockIdx.y * TG_T, n0 = blockIdx.x * TG_T;
  const int kchunk = ((a.K + a.ksplit - 1) / a.ksplit + TG_K - 1) / TG_K * TG_K;
  const int kbeg = z * kchunk, kend = min(a.K, kbeg + kchunk);
  const float* A = a.Ad[d] ? a.Ad[d] : a.A + d * a.a_dir;
  const float* B = a.Bd[d] ? a.Bd[d] : a.B + d * a.b_dir;
  float acc[4][4] = {};
  float ra[4], rb[4];
  auto fetch = [&](int k0) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int e = tid + 256 * i;
      const int m = A_KC ? e / TG_K : e % TG_T, k = A_KC ? e % TG_K : e / TG_T;
      ra[i] = (m0 + m < a.M && k0 + k < kend) ? __ldg(A + (int64_t)(m0 + m) * a.sam + (int64_t)(k0 + k) * a.sak) : 0.f;
      const int n = B_NC ? e % TG_T : e / TG_K, kb = B_NC ? e / TG_T : e % TG_K;
      rb[i] = (n0 + n < a.N && k0 + kb < kend) ? __ldg(B + (int64_t)(k0 + kb) * a.sbk + (int64_t)(n0 + n) * a.sbn) : 0.f;
    }
  };
  auto stash = [&](int buf) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int e = tid + 256 * i;
      const int m = A_KC ? e / TG_K : e % TG_T, k = A_KC ? e % TG_K : e / TG_T;
      As[buf][k][m] = ra[i];
      const int n = B_NC ? e % TG_T : e / TG_K, kb = B_NC ? e / TG_T : e % TG_K;
      Bs[buf][kb][n] = rb[i];
    }
  };
  if (kbeg < kend) {
    fetch(kbeg);
    stash(0);
  }
  __syncthreads();
  int buf = 0;
  for (int k0 = kbeg; k0 < kend; k0 += TG_K, buf ^= 1) {
    const bool more = k0 + TG_K < kend;
    if (more) fetch(k0 + TG_K);
#pragma unroll
    for (int kk = 0; kk < TG_K; ++kk) {
      const float4 av = *reinterpret_cast<const float4*>(&As[buf][kk][ty * 4]);
      const float4 bv = *reinterpret_cast<const float4*>(&Bs[buf][kk][tx * 4]);
      const float ar[4] = {av.x, av.y, av.z, av.w}, br[4] = {bv.x, bv.y, bv.z, bv.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = __fmaf_rn(ar[i], br[j], acc[i][j]);
    }
    if (more) stash(buf ^ 1);
    __syncthreads();
  }
  float* C = (a.Cd[d] ? a.Cd[d] : a.C + d * a.c_dir) + z * a.c_split;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int m = m0 + ty * 4 + i;
    if (m >= a.M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int n = n0 + tx * 4 + j;
      if (n >= a.N) continue;
      float v = acc[i][j];
      if (a.bias && z == 0) v += a.bias[d * a.bias_dir + n];
      if (a.bias2 && z == 0) v += a.bias2[d * a.bias_dir + n];
      float* dst = C + (int64_t)m * a.ldc + n;
      *dst = a.accumulate ? *dst + v : v;
    }
  }
}

static int tg_gemm(TgArgs a, cudaStream_t st) {
  if (a.M <= 0 || a.N <= 0) return VFR_OK;
  if (a.dirs <= 0) a.dirs = 1;
  if (a.ksplit <= 0) a.ksplit = 1;
  dim3 grid((a.N + TG_T - 1) / TG_T, (a.M + TG_T - 1) / TG_T, a.dirs * a.ksplit);
  const bool akc = a.sak == 1, bnc = a.sbn == 1;
  if (akc && bnc) tg_gemm_kernel<true, true><<<grid, 256, 0, st>>>(a);
  else if (akc) tg_gemm_kernel<true, false><<<grid, 256, 0, st>>>(a);
  else if (bnc) tg_gemm_kernel<false, true><<<grid, 256, 0, st>>>(a);
  else tg_gemm_kernel<false, false><<<grid, 256, 0, st>>>(a);
  return check_launch("tg_gemm_kernel");
}

// column sums: out[d][n] = sum_m X[d][m][n]  (bias gradients); one block per 32 columns, fixed order (deterministic)
__global__ void tg_colsum_kernel(const float* __restrict__ x, int64_t rows, int n, int64_t ld, int64_t x_dir, float* __restrict__ out,
                                 int64_t out_dir) {
  __shared__ float part[8][32];
  const int d = blockIdx.y;
  const int col = blockIdx.x * 32 + (threadIdx.x & 31), w = threadIdx.x >> 5;
  float s = 0.f;
  if (col < n)
    for (int64_t r = w; r < rows; r += 8) s += x[d * x_dir + r * ld + col];
  part[w][threadIdx.x & 31] = s;
  __syncthreads();
  if (w == 0 && col < n) {
    float t = 0.f;
    for (int i = 0; i < 8; ++i) t += part[i][threadIdx.x & 31];
    out[d * out_dir + col] = t;
  }
}

// out[i] = act(sum_z part[z][i] + bias[i % n]) : the consumer of split-K partials (fixed summation order)
__global__ void tg_reduce_kernel(const float* __restrict__ part, int ks, int64_t total, int n, const float* __restrict__ bias,
                                 int relu, float* __restrict__ out) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  float v = part[i];
  for (int z = 1; z < ks; ++z) v += part[(int64_t)z * total + i];
  if (bias) v += bias[i % n];
  out[i] = relu ? fmaxf(v, 0.f) : v;
}

// ---------------------------------------------------------------------------------------------------
// text branch
// ---------------------------------------------------------------------------------------------------
// embeddings of all (t, b): X[t * B + b][E] (+ the normalised rows xn and 1/(|x|+eps) when the learnable length is on)
__global__ void tt_gather_kernel(const int64_t* __restrict__ tokens, int B, int T, const float* __restrict__ table, int64_t vocab,
                                 const float* __restrict__ length, int E, float* __restrict__ X, float* __restrict__ Xn,
                                 int* __restrict__ bad) {
  const int lane = threadIdx.x & 31;
  const int64_t w = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (w >= (int64_t)B * T) return;
  const int b = (int)(w / T), t = (int)(w % T);
  int64_t id = tokens[(int64_t)b * T + t];
  if (id < 0 || id >= vocab) { if (lane == 0) atomicExch(bad, 1); id = 0; }
  const float* row = table + id * E;
  float* dst = X + ((int64_t)t * B + b) * E;
  if (!length) {
    for (int k = lane; k < E; k += 32) dst[k] = row[k];
    return;
  }
  float ss = 0.f;
  for (int k = lane; k < E; k += 32) ss = __fmaf_rn(row[k], row[k], ss);
  ss = warp_sum(ss);
  const float denom = __fadd_rn(__fsqrt_rn(ss), VFR_NORM_EPS), len = length[id];
  float* dn = Xn + ((int64_t)t * B + b) * E;
  for (int k = lane; k < E; k += 32) {
    const float xn = __fdiv_rn(row[k], denom);
    dn[k] = xn;
    dst[k] = __fmul_rn(xn, len);
  }
}

struct TtCell {
  int B, H, T;
  const float* G;        // [2][ksplit][B][4H] recurrent products of this step (nullptr at the first step: h_prev = 0)
  int ksplit;
  const float* XP;       // [2][T*B][4H] input projections + both biases
  float* act;            // [2][T*B][4H] gate activations i, f, g, o
  float* C;              // [2][T*B][H]
  float* Hs;             // [2][T*B][H]
  int step;              // 0 .. T-1 in processing order
};
__device__ __forceinline__ float sigmoidf_(float x) { return 1.f / (1.f + expf(-x)); }

__global__ void tt_cell_fwd_kernel(const TtCell p) {
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int d = blockIdx.y;
  if (idx >= (int64_t)p.B * p.H) return;
  const int b = (int)(idx / p.H), j = (int)(idx % p.H);
  const int t = d == 0 ? p.step : p.T - 1 - p.step, tp = d == 0 ? t - 1 : t + 1;
  const int64_t row = (int64_t)t * p.B + b, rowp = (int64_t)tp * p.B + b;
  const int64_t TB = (int64_t)p.T * p.B, H4 = 4 * (int64_t)p.H;
  float pre[4];
#pragma unroll
  for (int g = 0; g < 4; ++g) {
    float v = p.XP[(d * TB + row) * H4 + g * p.H + j];
    if (p.G)
      for (int z = 0; z < p.ksplit; ++z) v += p.G[(((int64_t)d * p.ksplit + z) * p.B + b) * H4 + g * p.H + j];
    pre[g] = v;
  }
  const float i = sigmoidf_(pre[0]), f = sigmoidf_(pre[1]), g = tanhf(pre[2]), o = sigmoidf_(pre[3]);
  const float cp = p.step == 0 ? 0.f : p.C[(d * TB + rowp) * p.H + j];
  const float c = f * cp + i * g;
  const float h = o * tanhf(c);
  float* a = p.act + (d * TB + row) * H4 + j;
  a[0] = i; a[p.H] = f; a[2 * (int64_t)p.H] = g; a[3 * (int64_t)p.H] = o;
  p.C[(d * TB + row) * p.H + j] = c;
  p.Hs[(d * TB + row) * p.H + j] = h;
}

// [h_fwd(T-1) | h_bwd(0)]  (models.py:66)
__global__ void tt_hcat_kernel(const float* __restrict__ Hs, int B, int H, int T, float* __restrict__ hcat) {
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (int64_t)B * 2 * H) return;
  const int b = (int)(idx / (2 * H)), c = (int)(idx % (2 * H));
  const int d = c / H, j = c % H;
  const int t = d == 0 ? T - 1 : 0;
  hcat[idx] = Hs[(((int64_t)d * T + t) * B + b) * H + j];
}

struct TtCellBwd {
  int B, H, T, step, ksplit;
  const float* act; const float* C;
  const float* dhcat;    // [B][2H] gradient of the final hidden states (enters at the LAST processing step)
  const float* dHpart;   // [2][ksplit][B][H] split-K partials of dG(next step) * W_hh (nullptr at the last step)
  float* dC;             // [2][B][H] carried cell gradient (in place)
  float* dG;             // [2][T*B][4H] pre-activation gate gradients
};
__global__ void tt_cell_bwd_kernel(const TtCellBwd p) {
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int d = blockIdx.y;
  if (idx >= (int64_t)p.B * p.H) return;
  const int b = (int)(idx / p.H), j = (int)(idx % p.H);
  const int t = d == 0 ? p.step : p.T - 1 - p.step, tp = d == 0 ? t - 1 : t + 1;
  const int64_t TB = (int64_t)p.T * p.B, H4 = 4 * (int64_t)p.H;
  const int64_t row = (int64_t)t * p.B + b, rowp = (int64_t)tp * p.B + b;
  float dh = 0.f;
  if (p.step == p.T - 1) dh = p.dhcat[(int64_t)b * 2 * p.H + d * p.H + j];
  if (p.dHpart)
    for (int z = 0; z < p.ksplit; ++z) dh += p.dHpart[(((int64_t)d * p.ksplit + z) * p.B + b) * p.H + j];
  const float* a = p.act + (d * TB + row) * H4 + j;
  const float i = a[0], f = a[p.H], g = a[2 * (int64_t)p.H], o = a[3 * (int64_t)p.H];
  const float c = p.C[(d * TB + row) * p.H + j];
  const float cp = p.step == 0 ? 0.f : p.C[(d * TB + rowp) * p.H + j];
  const float tc = tanhf(c);
  float* dcp = p.dC + ((int64_t)d * p.B + b) * p.H + j;
  const float dc = (p.step == p.T - 1 ? 0.f : *dcp) + dh * o * (1.f - tc * tc);
  float* g_out = p.dG + (d * TB + row) * H4 + j;
  g_out[0] = dc * g * i * (1.f - i);
  g_out[p.H] = dc * cp * f * (1.f - f);
  g_out[2 * (int64_t)p.H] = dc * i * (1.f - g * g);
  g_out[3 * (int64_t)p.H] = dh * tc * o * (1.f - o);
  *dcp = dc * f;
}

// gradient of the learnable word length (models.py:36-38,62-64): x = xn * len[id]  =>  dlen[id] += dX . xn ; the padding
// row (padding_idx = 0) receives no gradient, as in nn.Embedding
__global__ void tt_length_grad_kernel(const int64_t* __restrict__ tokens, int B, int T, int E, const float* __restrict__ dX,
                                      const float* __restrict__ Xn, float* __restrict__ dlen) {
  const int lane = threadIdx.x & 31;
  const int64_t w = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (w >= (int64_t)B * T) return;
  const int b = (int)(w / T), t = (int)(w % T);
  const int64_t id = tokens[(int64_t)b * T + t];
  if (id == 0) return;
  const int64_t row = (int64_t)t * B + b;
  float s = 0.f;
  for (int k = lane; k < E; k += 32) s = __fmaf_rn(dX[row * E + k], Xn[row * E + k], s);
  s = warp_sum(s);
  if (lane == 0) atomicAdd(dlen + id, s);
}

struct TtWs {
  float *X, *Xn, *XP, *act, *C, *Hs, *hcat, *G, *dG, *dC, *dHpart, *dhcat, *dX;
  int* bad;
  size_t bytes;
};
constexpr int TT_KSPLIT_FWD = 2;      // K = H:  2 dirs x (B/64) x (4H/64) x 2 CTAs, 32 k-iterations each
constexpr int TT_KSPLIT_BWD = 8;      // K = 4H: 2 dirs x (B/64) x (H/64) x 8 CTAs, 32 k-iterations each
static TtWs tt_ws(void* base, int B, int T, int H, int E, int D) {
  TtWs w;
  uint8_t* p = reinterpret_cast<uint8_t*>(base);
  auto take = [&](size_t n) { float* q = reinterpret_cast<float*>(p); p += (n * 4 + 255) / 256 * 256; return q; };
  const size_t TB = (size_t)T * B;
  w.bad = reinterpret_cast<int*>(take(64));
  w.X = take(TB * E);
  w.Xn = take(TB * E);
  w.XP = take(2 * TB * 4 * H);
  w.act = take(2 * TB * 4 * H);
  w.C = take(2 * TB * H);
  w.Hs = take(2 * TB * H);
  w.hcat = take((size_t)B * 2 * H);
  w.G = take((size_t)2 * TT_KSPLIT_FWD * B * 4 * H);
  w.dG = take(2 * TB * 4 * H);
  w.dC = take((size_t)2 * B * H);
  w.dHpart = take((size_t)2 * TT_KSPLIT_BWD * B * H);
  w.dhcat = take((size_t)B * 2 * H);
  w.dX = take(TB * E);
  (void)D;
  w.bytes = (size_t)(p - reinterpret_cast<uint8_t*>(base));
  return w;
}

// ---------------------------------------------------------------------------------------------------
// visual branch backward
// ---------------------------------------------------------------------------------------------------
// dpre = (dE W2) * [hidden > 0]   (hidden is the post-ReLU activation saved by the forward)
__global__ void tv_relu_bwd_kernel(const float* __restrict__ hidden, float* __restrict__ dh, int64_t n) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n && !(hidden[i] > 0.f)) dh[i] = 0.f;
}

// ---------------------------------------------------------------------------------------------------
// fused multi-tensor Adam + gradient norms
// ---------------------------------------------------------------------------------------------------
constexpr int AD_MAX = 24;
struct AdamArgs {
  float* p[AD_MAX]; const float* g[AD_MAX]; float* m[AD_MAX]; float* v[AD_MAX];
  int64_t n[AD_MAX];
  int64_t start[AD_MAX + 1];     // prefix sums of the block counts
  int count;
  float lr_over_bias1, inv_sqrt_bias2, beta1, beta2, eps, wd;
};
constexpr int AD_BLOCK = 256, AD_PER = 4;
__global__ void __launch_bounds__(AD_BLOCK) adam_kernel(const AdamArgs a) {
  int ti = 0;
  while (ti + 1 < a.count && (int64_t)blockIdx.x >= a.start[ti + 1]) ++ti;
  const int64_t base = ((int64_t)blockIdx.x - a.start[ti]) * AD_BLOCK * AD_PER;
#pragma unroll
  for (int u = 0; u < AD_PER; ++u) {
    const int64_t i = base + (int64_t)u * AD_BLOCK + threadIdx.x;
    if (i >= a.n[ti]) return;
    const float p = a.p[ti][i];
    const float g = __fmaf_rn(a.wd, p, a.g[ti][i]);                          // L2 weight decay folded into the gradient
    const float m = __fmaf_rn(a.beta1, a.m[ti][i], (1.f - a.beta1) * g);
    const float v = __fmaf_rn(a.beta2, a.v[ti][i], (1.f - a.beta2) * g * g);
    a.m[ti][i] = m;
    a.v[ti][i] = v;
    a.p[ti][i] = p - a.lr_over_bias1 * (m / (__fsqrt_rn(v) * a.inv_sqrt_bias2 + a.eps));
  }
}

struct NormArgs { const float* g[AD_MAX]; int64_t n[AD_MAX]; int count; };
// out[i] = || g_i ||_2 : one block per tensor (these are ~14 tensors of <= 4 M elements; latency, not bandwidth)
__global__ void __launch_bounds__(1024) grad_norm_kernel(const NormArgs a, float* __restrict__ out) {
  __shared__ double part[32];
  const int ti = blockIdx.x;
  double s = 0.0;
  for (int64_t i = threadIdx.x; i < a.n[ti]; i += blockDim.x) { const double x = a.g[ti][i]; s += x * x; }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int i = 0; i < 32; ++i) t += part[i];
    out[ti] = (float)sqrt(t);
  }
}

}  // namespace vfr

using namespace vfr;

// ---- text branch ------------------------------------------------------------------------------------------------
extern "C" size_t vfr_text_train_bytes(int n_queries, int seq_len, int hidden, int emb, int dim) {
  if (n_queries <= 0 || seq_len <= 0 || hidden <= 0 || emb <= 0 || dim <= 0) return 0;
  return tt_ws(nullptr, n_queries, seq_len, hidden, emb, dim).bytes;
}

extern "C" int vfr_text_train_fwd(const int64_t* tokens, int n_queries, int seq_len, const float* table, int64_t vocab,
                                  const float* length_table, int emb, const float* const* w_ih, const float* const* w_hh,
                                  const float* const* b_ih, const float* const* b_hh, int hidden, const float* fc_w,
                                  const float* fc_b, int dim, void* workspace, float* out, vfr_stream_t stream) {
  VFR_REQUIRE(tokens && table && w_ih && w_hh && b_ih && b_hh && fc_w && fc_b && workspace && out, VFR_ERR_INVALID,
              "vfr_text_train_fwd: null pointer");
  VFR_REQUIRE(n_queries > 0 && seq_len > 0 && hidden > 0 && emb > 0 && dim > 0 && vocab > 0, VFR_ERR_INVALID,
              "vfr_text_train_fwd: bad shape");
  const int B = n_queries, T = seq_len, H = hidden, E = emb;
  const TtWs w = tt_ws(workspace, B, T, H, E, dim);
  cudaStream_t st = (cudaStream_t)stream;
  const int64_t TB = (int64_t)T * B, H4 = 4 * (int64_t)H;
  VFR_CUDA(cudaMemsetAsync(w.bad, 0, sizeof(int), st));
  tt_gather_kernel<<<(unsigned)((TB + 7) / 8), 256, 0, st>>>(tokens, B, T, table, vocab, length_table, E, w.X, w.Xn, w.bad);
  int rc = check_launch("tt_gather_kernel");
  if (rc) return rc;
  // input projections of all steps, one GEMM per direction (the two weight sets are separate allocations):
  // XP = X W_ih^T + b_ih + b_hh
  for (int d = 0; d < 2; ++d) {
    TgArgs g{};
    g.A = w.X; g.B = w_ih[d]; g.C = w.XP + d * TB * H4;
    g.M = (int)TB; g.N = (int)H4; g.K = E;
    g.sam = E; g.sak = 1; g.sbk = 1; g.sbn = E; g.ldc = H4;
    g.bias = b_ih[d];
    g.bias2 = b_hh[d];
    rc = tg_gemm(g, st);
    if (rc) return rc;
  }
  for (int s = 0; s < T; ++s) {
    if (s > 0) {
      TgArgs g{};                              // both directions in one launch: G[d] = h_prev[d] W_hh[d]^T (split-K partials)
      for (int d = 0; d < 2; ++d) {
        const int tp = d == 0 ? s - 1 : T - s;             // previous state's time index
        g.Ad[d] = w.Hs + (d * TB + (int64_t)tp * B) * H;
        g.Bd[d] = w_hh[d];
        g.Cd[d] = w.G + (int64_t)d * TT_KSPLIT_FWD * B * H4;
      }
      g.dirs = 2;
      g.M = B; g.N = (int)H4; g.K = H;
      g.sam = H; g.sak = 1; g.sbk = 1; g.sbn = H; g.ldc = H4;
      g.ksplit = TT_KSPLIT_FWD; g.c_split = (int64_t)B * H4;
      rc = tg_gemm(g, st);
      if (rc) return rc;
    }
    TtCell c{B, H, T, s > 0 ? w.G : nullptr, TT_KSPLIT_FWD, w.XP, w.act, w.C, w.Hs, s};
    tt_cell_fwd_kernel<<<dim3((unsigned)(((int64_t)B * H + 255) / 256), 2), 256, 0, st>>>(c);
    rc = check_launch("tt_cell_fwd_kernel");
    if (rc) return rc;
  }
  tt_hcat_kernel<<<(unsigned)(((int64_t)B * 2 * H + 255) / 256), 256, 0, st>>>(w.Hs, B, H, T, w.hcat);
  rc = check_launch("tt_hcat_kernel");
  if (rc) return rc;
  TgArgs g{};
  g.A = w.hcat; g.B = fc_w; g.C = out;
  g.M = B; g.N = dim; g.K = 2 * H;
  g.sam = 2 * H; g.sak = 1; g.sbk = 1; g.sbn = 2 * H; g.ldc = dim;
  g.bias = fc_b;
  return tg_gemm(g, st);
}


// gradients: d_w_ih / d_w_hh / d_b_ih / d_b_hh are HOST arrays of two device pointers (forward, reverse direction);
// d_length fp32 [vocab] (accumulated into: zero it first) or NULL.  The same workspace as the forward call.
extern "C" int vfr_text_train_bwd(const int64_t* tokens, int n_queries, int seq_len, int64_t vocab, int has_length, int emb,
                                  const float* const* w_ih, const float* const* w_hh, int hidden, const float* fc_w, int dim,
                                  void* workspace, const float* grad_out, float* const* d_w_ih, float* const* d_w_hh,
                                  float* const* d_b_ih, float* const* d_b_hh, float* d_fc_w, float* d_fc_b, float* d_length,
                                  vfr_stream_t stream) {
  VFR_REQUIRE(tokens && w_ih && w_hh && fc_w && workspace && grad_out && d_w_ih && d_w_hh && d_b_ih && d_b_hh && d_fc_w && d_fc_b,
              VFR_ERR_INVALID, "vfr_text_train_bwd: null pointer");
  VFR_REQUIRE(!has_length || d_length, VFR_ERR_INVALID, "vfr_text_train_bwd: the learnable length needs d_length");
  const int B = n_queries, T = seq_len, H = hidden, E = emb;
  const TtWs w = tt_ws(workspace, B, T, H, E, dim);
  cudaStream_t st = (cudaStream_t)stream;
  const int64_t TB = (int64_t)T * B, H4 = 4 * (int64_t)H;
  int rc;
  {   // final projection: d fc_w = dOut^T hcat ; d fc_b = colsum(dOut) ; d hcat = dOut fc_w
    TgArgs g{};
    g.A = grad_out; g.B = w.hcat; g.C = d_fc_w;
    g.M = dim; g.N = 2 * H; g.K = B;
    g.sam = 1; g.sak = dim; g.sbk = 2 * H; g.sbn = 1; g.ldc = 2 * H;
    rc = tg_gemm(g, st);
    if (rc) return rc;
    tg_colsum_kernel<<<dim3((dim + 31) / 32, 1), 256, 0, st>>>(grad_out, B, dim, dim, 0, d_fc_b, 0);
    rc = check_launch("tg_colsum_kernel");
    if (rc) return rc;
    TgArgs h{};
    h.A = grad_out; h.B = fc_w; h.C = w.dhcat;
    h.M = B; h.N = 2 * H; h.K = dim;
    h.sam = dim; h.sak = 1; h.sbk = 2 * H; h.sbn = 1; h.ldc = 2 * H;
    rc = tg_gemm(h, st);
    if (rc) return rc;
  }
  for (int s = T - 1; s >= 0; --s) {
    TtCellBwd c{B, H, T, s, TT_KSPLIT_BWD, w.act, w.C, w.dhcat, s == T - 1 ? nullptr : w.dHpart, w.dC, w.dG};
    tt_cell_bwd_kernel<<<dim3((unsigned)(((int64_t)B * H + 255) / 256), 2), 256, 0, st>>>(c);
    rc = check_launch("tt_cell_bwd_kernel");
    if (rc) return rc;
    if (s > 0) {
      TgArgs g{};                              // d h_prev = dG(step s) W_hh, both directions in one launch; split-K partials
      for (int d = 0; d < 2; ++d) {            //   are summed by the next cell kernel
        const int t = d == 0 ? s : T - 1 - s;
        g.Ad[d] = w.dG + (d * TB + (int64_t)t * B) * H4;
        g.Bd[d] = w_hh[d];
        g.Cd[d] = w.dHpart + (int64_t)d * TT_KSPLIT_BWD * B * H;
      }
      g.dirs = 2;
      g.M = B; g.N = H; g.K = (int)H4;
      g.sam = H4; g.sak = 1; g.sbk = H; g.sbn = 1; g.ldc = H;
      g.ksplit = TT_KSPLIT_BWD; g.c_split = (int64_t)B * H;
      rc = tg_gemm(g, st);
      if (rc) return rc;
    }
  }
  for (int d = 0; d < 2; ++d) {
    // weight gradients of all steps in one GEMM per matrix (K = (T-1) B resp. T B):
    //   d W_hh = dG[steps with a predecessor]^T H[their predecessors] ; d W_ih = dG^T X ; d b = colsum(dG)
    const float* dGd = w.dG + d * TB * H4;
    const float* Hd = w.Hs + d * TB * H;
    TgArgs g{};
    g.A = d == 0 ? dGd + (int64_t)B * H4 : dGd;                 // forward: t = 1 .. T-1 pairs with h(t-1); reverse: t = 0 .. T-2 with h(t+1)
    g.B = d == 0 ? Hd : Hd + (int64_t)B * H;
    g.C = d_w_hh[d];
    g.M = (int)H4; g.N = H; g.K = (int)((int64_t)(T - 1) * B);
    g.sam = 1; g.sak = H4; g.sbk = H; g.sbn = 1; g.ldc = H;
    if (T > 1) rc = tg_gemm(g, st);
    else VFR_CUDA(cudaMemsetAsync(d_w_hh[d], 0, (size_t)H4 * H * 4, st));
    if (rc) return rc;
    TgArgs x{};
    x.A = dGd; x.B = w.X; x.C = d_w_ih[d];
    x.M = (int)H4; x.N = E; x.K = (int)TB;
    x.sam = 1; x.sak = H4; x.sbk = E; x.sbn = 1; x.ldc = E;
    rc = tg_gemm(x, st);
    if (rc) return rc;
    tg_colsum_kernel<<<dim3((unsigned)((H4 + 31) / 32), 1), 256, 0, st>>>(dGd, TB, (int)H4, H4, 0, d_b_ih[d], 0);
    rc = check_launch("tg_colsum_kernel");
    if (rc) return rc;
    VFR_CUDA(cudaMemcpyAsync(d_b_hh[d], d_b_ih[d], (size_t)H4 * 4, cudaMemcpyDeviceToDevice, st));
  }
  if (has_length) {
    // d X = dG_f W_ih_f + dG_r W_ih_r ; d len[id] += d X . xn
    for (int d = 0; d < 2; ++d) {
      TgArgs g{};
      g.A = w.dG + d * TB * H4; g.B = w_ih[d]; g.C = w.dX;
      g.M = (int)TB; g.N = E; g.K = (int)H4;
      g.sam = H4; g.sak = 1; g.sbk = E; g.sbn = 1; g.ldc = E;
      g.accumulate = d;
      rc = tg_gemm(g, st);
      if (rc) return rc;
    }
    tt_length_grad_kernel<<<(unsigned)((TB + 7) / 8), 256, 0, st>>>(tokens, B, T, E, w.dX, w.Xn, d_length);
    rc = check_launch("tt_length_grad_kernel");
    if (rc) return rc;
  }
  (void)vocab;
  return VFR_OK;
}

// bad-token flag of the last vfr_text_train_fwd on this workspace: DEVICE int32 (1 = an id was outside [0, vocab))
extern "C" const int32_t* vfr_text_train_flag(const void* workspace) { return reinterpret_cast<const int32_t*>(workspace); }

// ---- visual branch -----------------------------------------------------------------------------------------------
// backward of  e = relu(x W1^T + b1) W2^T + b2  given x [n, in_dim], the saved post-ReLU hidden [n, hid] and dE [n, dim]:
//   d W2 = dE^T hidden ; d b2 = colsum(dE) ; d hidden = dE W2 (masked by hidden > 0) ; d W1 = dpre^T x ; d b1 = colsum(dpre)
// scratch: fp32 [n, hid].  d_x (optional, [n, in_dim]) = dpre W1.
extern "C" int vfr_visual_train_bwd(const float* x, int64_t n_rows, int in_dim, const float* hidden, int hid, const float* w1,
                                    const float* w2, int dim, const float* grad_out, float* scratch, float* d_w1, float* d_b1,
                                    float* d_w2, float* d_b2, float* d_x, vfr_stream_t stream) {
  VFR_REQUIRE(x && hidden && w1 && w2 && grad_out && scratch && d_w1 && d_b1 && d_w2 && d_b2, VFR_ERR_INVALID,
              "vfr_visual_train_bwd: null pointer");
  VFR_REQUIRE(n_rows > 0 && n_rows < (int64_t(1) << 31) && in_dim > 0 && hid > 0 && dim > 0, VFR_ERR_INVALID,
              "vfr_visual_train_bwd: bad shape");
  cudaStream_t st = (cudaStream_t)stream;
  const int n = (int)n_rows;
  int rc;
  TgArgs a{};
  a.A = grad_out; a.B = hidden; a.C = d_w2;
  a.M = dim; a.N = hid; a.K = n;
  a.sam = 1; a.sak = dim; a.sbk = hid; a.sbn = 1; a.ldc = hid;
  rc = tg_gemm(a, st);
  if (rc) return rc;
  tg_colsum_kernel<<<dim3((dim + 31) / 32, 1), 256, 0, st>>>(grad_out, n, dim, dim, 0, d_b2, 0);
  rc = check_launch("tg_colsum_kernel");
  if (rc) return rc;
  TgArgs b{};
  b.A = grad_out; b.B = w2; b.C = scratch;
  b.M = n; b.N = hid; b.K = dim;
  b.sam = dim; b.sak = 1; b.sbk = hid; b.sbn = 1; b.ldc = hid;
  rc = tg_gemm(b, st);
  if (rc) return rc;
  tv_relu_bwd_kernel<<<(unsigned)(((int64_t)n * hid + 255) / 256), 256, 0, st>>>(hidden, scratch, (int64_t)n * hid);
  rc = check_launch("tv_relu_bwd_kernel");
  if (rc) return rc;
  TgArgs c{};
  c.A = scratch; c.B = x; c.C = d_w1;
  c.M = hid; c.N = in_dim; c.K = n;
  c.sam = 1; c.sak = hid; c.sbk = in_dim; c.sbn = 1; c.ldc = in_dim;
  rc = tg_gemm(c, st);
  if (rc) return rc;
  tg_colsum_kernel<<<dim3((hid + 31) / 32, 1), 256, 0, st>>>(scratch, n, hid, hid, 0, d_b1, 0);
  rc = check_launch("tg_colsum_kernel");
  if (rc) return rc;
  if (d_x) {
    TgArgs e{};
    e.A = scratch; e.B = w1; e.C = d_x;
    e.M = n; e.N = in_dim; e.K = hid;
    e.sam = hid; e.sak = 1; e.sbk = in_dim; e.sbn = 1; e.ldc = in_dim;
    rc = tg_gemm(e, st);
  }
  return rc;
}

// forward of the same MLP for the training step (n ~ 120 rows: the evaluation kernels would run it on 4 CTAs): split-K
// SGEMMs over all SMs, partials summed in a fixed order together with the bias (and the ReLU).  hidden [n, hid] is kept
// for the backward call.  scratch: vfr_visual_train_fwd_bytes(n, hid, dim).
constexpr int TV_KS1 = 16, TV_KS2 = 4;
extern "C" size_t vfr_visual_train_fwd_bytes(int64_t n_rows, int hid, int dim) {
  if (n_rows <= 0 || hid <= 0 || dim <= 0) return 0;
  return (size_t)n_rows * std::max((size_t)TV_KS1 * hid, (size_t)TV_KS2 * dim) * sizeof(float);
}
extern "C" int vfr_visual_train_fwd(const float* x, int64_t n_rows, int in_dim, const float* w1, const float* b1, int hid,
                                    const float* w2, const float* b2, int dim, float* scratch, float* hidden, float* out,
                                    vfr_stream_t stream) {
  VFR_REQUIRE(x && w1 && b1 && w2 && b2 && scratch && hidden && out, VFR_ERR_INVALID, "vfr_visual_train_fwd: null pointer");
  VFR_REQUIRE(n_rows > 0 && n_rows < (int64_t(1) << 31) && in_dim > 0 && hid > 0 && dim > 0, VFR_ERR_INVALID,
              "vfr_visual_train_fwd: bad shape");
  cudaStream_t st = (cudaStream_t)stream;
  const int n = (int)n_rows;
  TgArgs a{};
  a.A = x; a.B = w1; a.C = scratch;
  a.M = n; a.N = hid; a.K = in_dim;
  a.sam = in_dim; a.sak = 1; a.sbk = 1; a.sbn = in_dim; a.ldc = hid;
  a.ksplit = TV_KS1; a.c_split = (int64_t)n * hid;
  int rc = tg_gemm(a, st);
  if (rc) return rc;
  tg_reduce_kernel<<<(unsigned)(((int64_t)n * hid + 255) / 256), 256, 0, st>>>(scratch, TV_KS1, (int64_t)n * hid, hid, b1, 1, hidden);
  rc = check_launch("tg_reduce_kernel");
  if (rc) return rc;
  TgArgs b{};
  b.A = hidden; b.B = w2; b.C = scratch;
  b.M = n; b.N = dim; b.K = hid;
  b.sam = hid; b.sak = 1; b.sbk = 1; b.sbn = hid; b.ldc = dim;
  b.ksplit = TV_KS2; b.c_split = (int64_t)n * dim;
  rc = tg_gemm(b, st);
  if (rc) return rc;
  tg_reduce_kernel<<<(unsigned)(((int64_t)n * dim + 255) / 256), 256, 0, st>>>(scratch, TV_KS2, (int64_t)n * dim, dim, b2, 0, out);
  return check_launch("tg_reduce_kernel");
}

// ---- optimiser ---------------------------------------------------------------------------------------------------
// ONE launch updates up to 24 parameter tensors (torch.optim.Adam semantics, main.py:358: L2 weight decay folded into
// the gradient, bias-corrected moments, eps after the sqrt).  params / grads / exp_avg / exp_avg_sq: HOST arrays of
// device pointers, numel HOST int64 [count]; step = 1-based step count of this update.
extern "C" int vfr_adam_step(float* const* params, const float* const* grads, float* const* exp_avg, float* const* exp_avg_sq,
                             const int64_t* numel, int count, int64_t step, float lr, float beta1, float beta2, float eps,
                             float weight_decay, vfr_stream_t stream) {
  VFR_REQUIRE(params && grads && exp_avg && exp_avg_sq && numel, VFR_ERR_INVALID, "vfr_adam_step: null pointer");
  VFR_REQUIRE(count >= 1 && count <= AD_MAX && step >= 1, VFR_ERR_INVALID, "vfr_adam_step: 1..%d tensors, step >= 1", AD_MAX);
  AdamArgs a{};
  a.count = count;
  int64_t blocks = 0;
  for (int i = 0; i < count; ++i) {
    VFR_REQUIRE(params[i] && grads[i] && exp_avg[i] && exp_avg_sq[i] && numel[i] > 0, VFR_ERR_INVALID, "vfr_adam_step: tensor %d", i);
    a.p[i] = params[i]; a.g[i] = grads[i]; a.m[i] = exp_avg[i]; a.v[i] = exp_avg_sq[i]; a.n[i] = numel[i];
    a.start[i] = blocks;
    blocks += (numel[i] + AD_BLOCK * AD_PER - 1) / (AD_BLOCK * AD_PER);
  }
  a.start[count] = blocks;
  const double b1 = 1.0 - pow((double)beta1, (double)step), b2 = 1.0 - pow((double)beta2, (double)step);
  a.lr_over_bias1 = (float)((double)lr / b1);
  a.inv_sqrt_bias2 = (float)(1.0 / sqrt(b2));
  a.beta1 = beta1; a.beta2 = beta2; a.eps = eps; a.wd = weight_decay;
  adam_kernel<<<(unsigned)blocks, AD_BLOCK, 0, (cudaStream_t)stream>>>(a);
  return check_launch("adam_kernel");
}

// out fp32 [count] (DEVICE) = the L2 norm of every gradient tensor (model/utils.py:85-92 takes their mean on the host)
extern "C" int vfr_grad_norms(const float* const* grads, const int64_t* numel, int count, float* out, vfr_stream_t stream) {
  VFR_REQUIRE(grads && numel && out && count >= 1 && count <= AD_MAX, VFR_ERR_INVALID, "vfr_grad_norms: bad argument");
  NormArgs a{};
  a.count = count;
  for (int i = 0; i < count; ++i) { a.g[i] = grads[i]; a.n[i] = numel[i]; }
  grad_norm_kernel<<<count, 1024, 0, (cudaStream_t)stream>>>(a, out);
  return check_launch("grad_norm_kernel");
}
