// K2 on tensor cores: the visual MLP  e = relu(x W1^T + b1) W2^T + b2  (reference model/models.py:21-27,55-56, eval mode)
// as split-fp16 tcgen05 GEMMs (vfr_gemm_tc.cuh), in two forms:
//
//  * vfr_visual_embed_split - the SPLIT-WEIGHT form of the reference's feature assembly (model/data.py:204-213): a clip row
//    is [segment | context | tef] with the context feature repeated for every clip of its video and tef = (i/n, (i+1)/n),
//    so   x W1^T = seg W1[:, :F]^T  +  ctx W1[:, F:2F]^T  +  tef W1[:, 2F:]^T .
//    Inputs are what K1 produces - seg fp32 [C, F], ctx fp32 [V, F], the CSR clip offsets - the 8194-wide concat is never
//    materialised, the context product is computed ONCE per video ([V, hid] instead of [C, hid]: half the FLOPs and bytes),
//    and the epilogue of the segment GEMM adds it per clip together with the two tef columns and the bias.
//  * vfr_visual_embed_tc - the general form on already assembled rows x fp32 [N, 2F+2] (what CALModel.forward receives from
//    the reference's iterators): one GEMM over K = 2F, tef columns and bias in the epilogue.
//
// Accuracy.  Operands are split into fp16 pairs (hi = fp16(x 2^e), lo = fp16(x 2^e - hi)): 22 significand bits, against
// the 16 of the split-bf16 operands of K3 - the product  Ah.Bh + Al.Bh + Ah.Bl  then carries a relative error of ~2^-22
// per term instead of ~2^-17, which is what holds K = 8194 all-positive features within 1e-5 of the fp32 reference.  The
// power-of-two exponents e (per operand tensor, from its maximum, computed on the device - no host round trip) keep hi
// AND lo inside fp16's range; the epilogue multiplies them out exactly.
#include "vfr_gemm_tc.cuh"
#include <algorithm>
#include <cuda_fp16.h>
#include <math_constants.h>
#include <stdlib.h>

namespace vfr {

struct VisDims {
  int F, hid, dim;
  int Fp;         // F rounded up to 64
  int K1;         // 2 * Fp : packed K of layer 1 = [segment | context]
  int Hp;         // hid rounded up to 64
  int N1, N2;     // output rows of the packed weights, rounded up to 256
};
static VisDims vis_dims(int F, int hid, int dim) {
  VisDims d;
  d.F = F; d.hid = hid; d.dim = dim;
  d.Fp = gt_kp(F);
  d.K1 = 2 * d.Fp;
  d.Hp = gt_kp(hid);
  d.N1 = (hid + GT_BN - 1) / GT_BN * GT_BN;
  d.N2 = (dim + GT_BN - 1) / GT_BN * GT_BN;
  return d;
}

// packed model blob: [W1 split-fp16 [N1, 2 K1] | W2 split-fp16 [N2, 2 Hp] | w1t fp32 [hid, 2] | b1 [hid] | b2 [dim] | exps int32 [4]]
struct VisBlob {
  __half* w1;
  __half* w2;
  float* w1t;
  float* b1;
  float* b2;
  int* exps;      // [0] = exponent of W1, [1] = of W2
  size_t bytes;
};
static VisBlob vis_blob(void* base, const VisDims& d) {
  VisBlob b;
  __half* p = reinterpret_cast<__half*>(base);
  b.w1 = p;
  b.w2 = p + (size_t)d.N1 * 2 * d.K1;
  float* f = reinterpret_cast<float*>(b.w2 + (size_t)d.N2 * 2 * d.Hp);
  b.w1t = f;
  b.b1 = f + 2 * (size_t)d.hid;
  b.b2 = b.b1 + d.hid;
  b.exps = reinterpret_cast<int*>(b.b2 + d.dim);
  b.bytes = ((size_t)d.N1 * 2 * d.K1 + (size_t)d.N2 * 2 * d.Hp) * 2 + (3 * (size_t)d.hid + d.dim) * 4 + 4 * sizeof(int);
  return b;
}

// ---- operand preparation -----------------------------------------------------------------------------------------
// max |x| over a strided fp32 matrix [rows, cols] (row pitch ld) -> atomicMax on the bit pattern (non-negative floats
// order like unsigned ints)
__global__ void vis_amax_kernel(const float* __restrict__ x, int64_t rows, int cols, int64_t ld, unsigned* __restrict__ amax) {
  float m = 0.f;
  const int64_t total = rows * cols;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = i / cols;
    m = fmaxf(m, fabsf(x[r * ld + (i - r * cols)]));
  }
  m = warp_max(m);
  if ((threadIdx.x & 31) == 0 && m > 0.f) atomicMax(amax, __float_as_uint(m));
}
// exponent e with max |x| 2^e in [2^13, 2^14): hi uses fp16's top binades, lo (<= 2^-11 |hi|) stays far above its subnormals
__global__ void vis_exp_kernel(const unsigned* __restrict__ amax, int* __restrict__ e) {
  const float m = __uint_as_float(*amax);
  int v = 0;
  if (m > 0.f && m < CUDART_INF_F) v = 13 - ilogbf(m);
  *e = max(-100, min(100, v));
}
__device__ __forceinline__ void split2h(float xs, __half& hi, __half& lo) {
  hi = __float2half_rn(xs);
  lo = __float2half_rn(xs - __half2float(hi));
}
// fp32 rows (column window [c0, c0 + k) of a matrix with row pitch ldx) -> columns [k0, k0 + kp) of split-fp16 packed rows
// (row pitch ldo halves, lo part lo_off columns after the hi part); one warp per row
__global__ void vis_split_rows_kernel(const float* __restrict__ x, int64_t rows, int c0, int k, int64_t ldx, int k0, int kp,
                                      int64_t ldo, int lo_off, const int* __restrict__ e, __half* __restrict__ out) {
  const int lane = threadIdx.x & 31;
  const int64_t r = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (r >= rows) return;
  const int ex = *e;
  const float* src = x + r * ldx + c0;
  __half* dst = out + r * ldo + k0;
  for (int c = lane; c < kp; c += 32) {
    __half hi, lo;
    split2h(c < k ? scalbnf(src[c], ex) : 0.f, hi, lo);
    dst[c] = hi;
    dst[lo_off + c] = lo;
  }
}
// The activations' form: ONE pass over HBM.  A warp owns a row, finds the row's max |x| over up to two column windows
// (the second sweep over the row hits L1 / L2), derives the ROW's exponent (max 2^e in [2^13, 2^14), stored in row_exp for
// the GEMM epilogue) and writes the split-fp16 packed row.  Window w = columns [c0[w], c0[w] + k) of x -> packed columns
// [k0[w], k0[w] + kp).  Replaces a per-tensor amax pass (403 MB at 1.1 TB/s = 0.37 ms of a 1.5 ms K2) + a split pass.
__global__ void vis_split_rows_auto_kernel(const float* __restrict__ x, int64_t rows, int n_win, int c0a, int c0b, int k,
                                           int64_t ldx, int k0a, int k0b, int kp, int64_t ldo, int lo_off,
                                           int* __restrict__ row_exp, __half* __restrict__ out) {
  const int lane = threadIdx.x & 31;
  const int64_t r = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (r >= rows) return;
  const float* row = x + r * ldx;
  const bool vec = ((reinterpret_cast<uintptr_t>(row + c0a) | reinterpret_cast<uintptr_t>(row + c0b)) & 15) == 0 && (k & 3) == 0;
  // one window of <= 4096 aligned columns (the split-weight form at F = 4096, the hidden layer): the row stays in
  // registers between the two sweeps (32 float4 per lane) - one read of the row instead of two
  if (vec && n_win == 1 && k <= 4096) {
    const float4* s4 = reinterpret_cast<const float4*>(row + c0a);
    const int n4 = k >> 2;
    float4 v[32];
    float m = 0.f;
#pragma unroll
    for (int u = 0; u < 32; ++u) {
      const int c = u * 32 + lane;
      v[u] = (c < n4) ? s4[c] : make_float4(0.f, 0.f, 0.f, 0.f);
      m = fmaxf(fmaxf(m, fmaxf(fabsf(v[u].x), fabsf(v[u].y))), fmaxf(fabsf(v[u].z), fabsf(v[u].w)));
    }
    m = warp_max(m);
    int ex = 0;
    if (m > 0.f && m < CUDART_INF_F) ex = max(-100, min(100, 13 - ilogbf(m)));
    if (lane == 0) row_exp[r] = ex;
    __half* dst = out + r * ldo + k0a;
#pragma unroll
    for (int u = 0; u < 32; ++u) {
      const int c = u * 32 + lane;
      if (4 * c < kp) {
        __half h[4], l[4];
        split2h(scalbnf(v[u].x, ex), h[0], l[0]);
        split2h(scalbnf(v[u].y, ex), h[1], l[1]);
        split2h(scalbnf(v[u].z, ex), h[2], l[2]);
        split2h(scalbnf(v[u].w, ex), h[3], l[3]);
        *reinterpret_cast<uint2*>(dst + 4 * c) = make_uint2((uint32_t)__half_as_ushort(h[0]) | ((uint32_t)__half_as_ushort(h[1]) << 16),
                                                            (uint32_t)__half_as_ushort(h[2]) | ((uint32_t)__half_as_ushort(h[3]) << 16));
        *reinterpret_cast<uint2*>(dst + lo_off + 4 * c) = make_uint2((uint32_t)__half_as_ushort(l[0]) | ((uint32_t)__half_as_ushort(l[1]) << 16),
                                                                     (uint32_t)__half_as_ushort(l[2]) | ((uint32_t)__half_as_ushort(l[3]) << 16));
      }
    }
    return;
  }
  float m = 0.f;
  for (int w = 0; w < n_win; ++w) {
    const float* src = row + (w ? c0b : c0a);
    if (vec) {
      const float4* s4 = reinterpret_cast<const float4*>(src);
      for (int c = lane; c < (k >> 2); c += 32) {
        const float4 v = s4[c];
        m = fmaxf(fmaxf(m, fmaxf(fabsf(v.x), fabsf(v.y))), fmaxf(fabsf(v.z), fabsf(v.w)));
      }
    } else {
      for (int c = lane; c < k; c += 32) m = fmaxf(m, fabsf(src[c]));
    }
  }
  m = warp_max(m);
  int ex = 0;
  if (m > 0.f && m < CUDART_INF_F) ex = max(-100, min(100, 13 - ilogbf(m)));
  if (lane == 0) row_exp[r] = ex;
  for (int w = 0; w < n_win; ++w) {
    const float* src = row + (w ? c0b : c0a);
    __half* dst = out + r * ldo + (w ? k0b : k0a);
    if (vec) {
      const float4* s4 = reinterpret_cast<const float4*>(src);
      for (int c = lane; c < (kp >> 2); c += 32) {
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (4 * c < k) v = s4[c];
        __half h[4], l[4];
        split2h(scalbnf(v.x, ex), h[0], l[0]);
        split2h(scalbnf(v.y, ex), h[1], l[1]);
        split2h(scalbnf(v.z, ex), h[2], l[2]);
        split2h(scalbnf(v.w, ex), h[3], l[3]);
        *reinterpret_cast<uint2*>(dst + 4 * c) = make_uint2((uint32_t)__half_as_ushort(h[0]) | ((uint32_t)__half_as_ushort(h[1]) << 16),
                                                            (uint32_t)__half_as_ushort(h[2]) | ((uint32_t)__half_as_ushort(h[3]) << 16));
        *reinterpret_cast<uint2*>(dst + lo_off + 4 * c) = make_uint2((uint32_t)__half_as_ushort(l[0]) | ((uint32_t)__half_as_ushort(l[1]) << 16),
                                                                     (uint32_t)__half_as_ushort(l[2]) | ((uint32_t)__half_as_ushort(l[3]) << 16));
      }
    } else {
      for (int c = lane; c < kp; c += 32) {
        __half hi, lo;
        split2h(c < k ? scalbnf(src[c], ex) : 0.f, hi, lo);
        dst[c] = hi;
        dst[lo_off + c] = lo;
      }
    }
  }
}
__global__ void vis_zero_rows_kernel(__half* __restrict__ out, int64_t n) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    out[i] = __float2half_rn(0.f);
}
__global__ void vis_small_pack_kernel(const float* __restrict__ w1, const float* __restrict__ b1, const float* __restrict__ b2,
                                      VisDims d, float* __restrict__ w1t, float* __restrict__ ob1, float* __restrict__ ob2) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const int ld = 2 * d.F + 2;
  if (i < d.hid) {
    w1t[2 * i] = w1[(int64_t)i * ld + 2 * d.F];
    w1t[2 * i + 1] = w1[(int64_t)i * ld + 2 * d.F + 1];
    ob1[i] = b1[i];
  }
  if (i < d.dim) ob2[i] = b2[i];
}
// clip -> (video, tef): tef = (i / n, (i + 1) / n) in fp32, the division model/data.py:208-210 performs
__global__ void vis_clip_meta_kernel(const int32_t* __restrict__ vid_off, int64_t n_videos, int32_t* __restrict__ clip_vid,
                                     float2* __restrict__ tef) {
  const int64_t v = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (v >= n_videos) return;
  const int c0 = vid_off[v], n = vid_off[v + 1] - c0;
  for (int i = 0; i < n; ++i) {
    clip_vid[c0 + i] = (int32_t)v;
    tef[c0 + i] = make_float2(__fdiv_rn((float)i, (float)n), __fdiv_rn((float)(i + 1), (float)n));
  }
}

// ---- epilogues ---------------------------------------------------------------------------------------------------
// layer 1: pre = acc 2^-(ea+ew) + add[row][n] (or b1[n]) + tef . w1t[n]  ->  relu  ->  hidden fp32
struct EpiVis1 {
  float* hidden;            // [M, hid]
  int hid;
  const int* ea;            // exponent of the A operand: ea[m] when ea_row, else ea[0]
  int ea_row;
  const int* ew;            // exponent of W1
  const float* b1;          // bias (used when add == nullptr)
  const float* add;         // optional per-VIDEO term [V, hid] (context product + bias), row = clip_vid[m]
  const int32_t* clip_vid;
  const float* tef;         // [M, tef_ld] the two tef values of row m start at tef + m * tef_ld
  int64_t tef_ld;
  const float* w1t;         // [hid, 2]
  struct Row {              // what a row needs for all of its column groups
    float s, t0, t1;
    const float* arow;
  };
  __device__ __forceinline__ Row row(int, int m) const {
    Row r;
    r.s = scalbnf(1.f, -(ea[ea_row ? m : 0] + *ew));
    r.t0 = tef ? tef[m * tef_ld] : 0.f;
    r.t1 = tef ? tef[m * tef_ld + 1] : 0.f;
    r.arow = add ? add + (int64_t)clip_vid[m] * hid : nullptr;
    return r;
  }
  __device__ __forceinline__ void operator()(int z, int m, int n0, const float (&v)[16]) const { (*this)(z, m, n0, v, row(z, m)); }
  __device__ __forceinline__ void operator()(int, int m, int n0, const float (&v)[16], const Row& rw) const {
    const float s = rw.s, t0 = rw.t0, t1 = rw.t1;
    const float* arow = rw.arow;
    if ((hid & 3) == 0) {
      // 16-byte accesses (every array here is 16-byte aligned and hid % 4 == 0): a thread owns a row, so each access of a
      // warp touches 32 rows - four columns per instruction instead of one
#pragma unroll
      for (int g = 0; g < 4; ++g) {
        const int n = n0 + 4 * g;
        if (n < hid) {
          const float4 a = arow ? *reinterpret_cast<const float4*>(arow + n) : __ldg(reinterpret_cast<const float4*>(b1 + n));
          const float4 wa = __ldg(reinterpret_cast<const float4*>(w1t + 2 * n)), wb = __ldg(reinterpret_cast<const float4*>(w1t + 2 * n + 4));
          float4 x;
          x.x = fmaxf(__fmaf_rn(t1, wa.y, __fmaf_rn(t0, wa.x, __fmaf_rn(v[4 * g + 0], s, a.x))), 0.f);
          x.y = fmaxf(__fmaf_rn(t1, wa.w, __fmaf_rn(t0, wa.z, __fmaf_rn(v[4 * g + 1], s, a.y))), 0.f);
          x.z = fmaxf(__fmaf_rn(t1, wb.y, __fmaf_rn(t0, wb.x, __fmaf_rn(v[4 * g + 2], s, a.z))), 0.f);
          x.w = fmaxf(__fmaf_rn(t1, wb.w, __fmaf_rn(t0, wb.z, __fmaf_rn(v[4 * g + 3], s, a.w))), 0.f);
          *reinterpret_cast<float4*>(hidden + (int64_t)m * hid + n) = x;
        }
      }
    } else {
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        const int n = n0 + j;
        if (n < hid) {
          float x = __fmaf_rn(v[j], s, arow ? arow[n] : __ldg(b1 + n));
          x = __fmaf_rn(t0, __ldg(w1t + 2 * n), x);
          x = __fmaf_rn(t1, __ldg(w1t + 2 * n + 1), x);
          x = fmaxf(x, 0.f);
          hidden[(int64_t)m * hid + n] = x;
        }
      }
    }
  }
};
// plain scaled output:  out = acc 2^-(ea+ew) + bias
struct EpiVisOut {
  float* out;
  int64_t ldo;
  int N;
  const int* ea;            // ea[m] when ea_row, else ea[0]
  int ea_row;
  const int* ew;
  const float* bias;
  __device__ __forceinline__ void operator()(int, int m, int n0, const float (&v)[16]) const {
    const float s = scalbnf(1.f, -(ea[ea_row ? m : 0] + *ew));
    if (((N | (int)ldo) & 3) == 0 && (reinterpret_cast<uintptr_t>(out) & 15) == 0) {
#pragma unroll
      for (int g = 0; g < 4; ++g) {
        const int n = n0 + 4 * g;
        if (n < N) {
          const float4 b = bias ? __ldg(reinterpret_cast<const float4*>(bias + n)) : make_float4(0.f, 0.f, 0.f, 0.f);
          *reinterpret_cast<float4*>(out + (int64_t)m * ldo + n) =
              make_float4(__fmaf_rn(v[4 * g], s, b.x), __fmaf_rn(v[4 * g + 1], s, b.y), __fmaf_rn(v[4 * g + 2], s, b.z),
                          __fmaf_rn(v[4 * g + 3], s, b.w));
        }
      }
      return;
    }
#pragma unroll
    for (int j = 0; j < 16; ++j) {
      const int n = n0 + j;
      if (n < N) out[(int64_t)m * ldo + n] = __fmaf_rn(v[j], s, bias ? __ldg(bias + n) : 0.f);
    }
  }
};

// workspace carve-up (all sizes for `rows` clip rows and `vids` videos)
struct VisWs {
  __half* a1;        // [rows, 2 K1] (general form) or [rows, 2 Fp] (split form)
  __half* ac;        // [vids, 2 Fp] (split form)
  __half* ah;        // [rows, 2 Hp]
  float* hidden;     // [rows, hid]
  float* cvec;       // [vids, hid]
  int32_t* clip_vid; // [rows]
  float2* tef;       // [rows]
  int* row_exp;      // [rows] exponents of the clip rows (layer 1), then of the hidden rows (layer 2)
  int* vid_exp;      // [vids] exponents of the context rows (split form)
  float* flush;      // [max(rows, vids), N1] fp32: K-segmented accumulation of the layer-1 GEMMs (vfr_gemm_tc.cuh)
  size_t bytes;
};
// K per accumulation segment of the layer-1 GEMMs.  The tensor core's fp32 accumulator truncates: integrated over the
// 512 accumulation steps of K = 8192 that is a ~1.5e-5 relative bias, 2.7e-5 of the embedding scale after layer 2
// (measured; the fp32 reference needs 1e-5).  Segments are drained into an fp32 buffer with round-to-nearest adds; measured on
// B200 (tools/k2_accuracy.py, 24 576 rows x 8194, error of the embeddings / their scale, time of the general form):
//   no flush 2.6e-5, 1.98 ms | K 2048: 7.7e-6, 2.08 | 1024: 4.4e-6, 2.24 | 512: 3.0e-6, 2.56 | 256: 2.5e-6, 3.22 | fp32 SGEMM 1.7e-6, 5.04
static int vis_flush_k() {
  const char* e = getenv("VFR_VIS_FLUSH");
  const int v = e ? atoi(e) : 512;
  return v <= 0 ? 0 : std::max(32, v / 32 * 32);
}
static VisWs vis_ws(void* base, const VisDims& d, int64_t rows, int64_t vids, bool split) {
  VisWs w;
  uint8_t* p = reinterpret_cast<uint8_t*>(base);
  auto take = [&](size_t n) { uint8_t* q = p; p += (n + 255) / 256 * 256; return q; };
  w.a1 = reinterpret_cast<__half*>(take((size_t)rows * 2 * (split ? d.Fp : d.K1) * 2));
  w.ac = reinterpret_cast<__half*>(take(split ? (size_t)vids * 2 * d.Fp * 2 : 0));
  w.ah = reinterpret_cast<__half*>(take((size_t)rows * 2 * d.Hp * 2));
  w.hidden = reinterpret_cast<float*>(take((size_t)rows * d.hid * 4));
  w.cvec = reinterpret_cast<float*>(take(split ? (size_t)vids * d.hid * 4 : 0));
  w.clip_vid = reinterpret_cast<int32_t*>(take(split ? (size_t)rows * 4 : 0));
  w.tef = reinterpret_cast<float2*>(take(split ? (size_t)rows * 8 : 0));
  w.row_exp = reinterpret_cast<int*>(take((size_t)rows * 4));
  w.vid_exp = reinterpret_cast<int*>(take(split ? (size_t)vids * 4 : 0));
  // (whole 256-row tiles: the CTA-pair kernel keeps a tile's partial sums in a tile-major layout)
  w.flush = reinterpret_cast<float*>(take((size_t)((std::max(rows, split ? vids : (int64_t)0) + 255) / 256 * 256) * d.N1 * 4));
  w.bytes = (size_t)(p - reinterpret_cast<uint8_t*>(base));
  return w;
}

static int vis_amax(const float* x, int64_t rows, int cols, int64_t ld, unsigned* amax, cudaStream_t st) {
  const int64_t total = rows * cols;
  const unsigned blocks = (unsigned)std::min<int64_t>((total + 255) / 256, 148 * 8);
  vis_amax_kernel<<<blocks, 256, 0, st>>>(x, rows, cols, ld, amax);
  return check_launch("vis_amax_kernel");
}

// layer 2 shared by both forms: hidden fp32 [rows, hid] -> out [rows, dim]
static int vis_layer2(const VisDims& d, const VisBlob& blob, const VisWs& w, int64_t rows, float* out, cudaStream_t st) {
  // (per-row exponents again - the layer-1 exponents in row_exp are dead once its GEMM has run)
  vis_split_rows_auto_kernel<<<(unsigned)((rows + 7) / 8), 256, 0, st>>>(w.hidden, rows, 1, 0, 0, d.hid, d.hid, 0, 0, d.Hp,
                                                                         2 * (int64_t)d.Hp, d.Hp, w.row_exp, w.ah);
  int rc = check_launch("vis_split_rows_auto_kernel");
  if (rc) return rc;
  const void* a[1] = {w.ah};
  const void* b[1] = {blob.w2};
  EpiVisOut epi{out, d.dim, d.dim, w.row_exp, 1, blob.exps + 1, blob.b2};
  return launch_gemm_tc(a, b, 1, (int)rows, d.dim, d.Hp, 2 * (int64_t)d.Hp, 2 * (int64_t)d.Hp, epi, st, nullptr, true);
}

}  // namespace vfr

using namespace vfr;

extern "C" size_t vfr_visual_pack_bytes(int feat_dim, int hid, int dim) {
  if (feat_dim <= 0 || hid <= 0 || dim <= 0) return 0;
  return vis_blob(nullptr, vis_dims(feat_dim, hid, dim)).bytes;
}

extern "C" int vfr_visual_pack(const float* w1, const float* b1, const float* w2, const float* b2, int feat_dim, int hid,
                               int dim, void* packed, vfr_stream_t stream) {
  VFR_REQUIRE(w1 && b1 && w2 && b2 && packed, VFR_ERR_INVALID, "vfr_visual_pack: null pointer");
  VFR_REQUIRE(feat_dim > 0 && hid > 0 && dim > 0, VFR_ERR_INVALID, "vfr_visual_pack: bad shape");
  const VisDims d = vis_dims(feat_dim, hid, dim);
  const VisBlob b = vis_blob(packed, d);
  cudaStream_t st = (cudaStream_t)stream;
  const int64_t ld1 = 2 * (int64_t)feat_dim + 2;
  VFR_CUDA(cudaMemsetAsync(b.exps, 0, 4 * sizeof(int), st));
  // (the exponent cells double as the amax scratch: amax is written as a bit pattern, then replaced by the exponent)
  unsigned* amax = reinterpret_cast<unsigned*>(b.exps) + 2;
  int rc = vis_amax(w1, hid, 2 * feat_dim, ld1, amax, st);
  if (rc) return rc;
  rc = vis_amax(w2, dim, hid, hid, amax + 1, st);
  if (rc) return rc;
  vis_exp_kernel<<<1, 1, 0, st>>>(amax, b.exps);
  vis_exp_kernel<<<1, 1, 0, st>>>(amax + 1, b.exps + 1);
  rc = check_launch("vis_exp_kernel");
  if (rc) return rc;
  // rows >= hid / dim of the packed weights stay zero (their outputs are masked by the epilogues anyway)
  vis_zero_rows_kernel<<<148 * 4, 256, 0, st>>>(b.w1, (int64_t)d.N1 * 2 * d.K1 + (int64_t)d.N2 * 2 * d.Hp);
  rc = check_launch("vis_zero_rows_kernel");
  if (rc) return rc;
  const unsigned g1 = (unsigned)((hid + 7) / 8);
  vis_split_rows_kernel<<<g1, 256, 0, st>>>(w1, hid, 0, feat_dim, ld1, 0, d.Fp, 2 * (int64_t)d.K1, d.K1, b.exps, b.w1);
  vis_split_rows_kernel<<<g1, 256, 0, st>>>(w1, hid, feat_dim, feat_dim, ld1, d.Fp, d.Fp, 2 * (int64_t)d.K1, d.K1, b.exps, b.w1);
  vis_split_rows_kernel<<<(unsigned)((dim + 7) / 8), 256, 0, st>>>(w2, dim, 0, hid, hid, 0, d.Hp, 2 * (int64_t)d.Hp, d.Hp,
                                                                   b.exps + 1, b.w2);
  rc = check_launch("vis_split_rows_kernel");
  if (rc) return rc;
  vis_small_pack_kernel<<<(std::max(hid, dim) + 255) / 256, 256, 0, st>>>(w1, b1, b2, d, b.w1t, b.b1, b.b2);
  return check_launch("vis_small_pack_kernel");
}

extern "C" size_t vfr_visual_embed_tc_bytes(int64_t n_rows, int64_t n_videos, int feat_dim, int hid, int dim, int split) {
  if (n_rows <= 0 || feat_dim <= 0 || hid <= 0 || dim <= 0) return 0;
  return vis_ws(nullptr, vis_dims(feat_dim, hid, dim), n_rows, n_videos > 0 ? n_videos : 1, split != 0).bytes;
}

extern "C" int vfr_visual_embed_tc(const float* x, int64_t n_rows, int feat_dim, const void* packed, int hid, int dim,
                                   void* workspace, float* out, vfr_stream_t stream) {
  VFR_REQUIRE(x && packed && workspace && out, VFR_ERR_INVALID, "vfr_visual_embed_tc: null pointer");
  VFR_REQUIRE(n_rows > 0 && n_rows < (int64_t(1) << 31) - 256 && feat_dim > 0 && hid > 0 && dim > 0, VFR_ERR_INVALID,
              "vfr_visual_embed_tc: bad shape");
  const VisDims d = vis_dims(feat_dim, hid, dim);
  const VisBlob blob = vis_blob(const_cast<void*>(packed), d);
  const VisWs w = vis_ws(workspace, d, n_rows, 1, false);
  cudaStream_t st = (cudaStream_t)stream;
  const int64_t ldx = 2 * (int64_t)feat_dim + 2;
  vis_split_rows_auto_kernel<<<(unsigned)((n_rows + 7) / 8), 256, 0, st>>>(x, n_rows, 2, 0, feat_dim, feat_dim, ldx, 0, d.Fp, d.Fp,
                                                                           2 * (int64_t)d.K1, d.K1, w.row_exp, w.a1);
  int rc = check_launch("vis_split_rows_auto_kernel");
  if (rc) return rc;
  {
    const void* a[1] = {w.a1};
    const void* b[1] = {blob.w1};
    EpiVis1 epi{w.hidden, hid, w.row_exp, 1, blob.exps, blob.b1, nullptr, nullptr, x + 2 * feat_dim, ldx, blob.w1t};
    const int fk = vis_flush_k();
    rc = launch_gemm_tc(a, b, 1, (int)n_rows, hid, d.K1, 2 * (int64_t)d.K1, 2 * (int64_t)d.K1, epi, st, nullptr, true, 0, 0, fk,
                        fk ? w.flush : nullptr, d.N1);
    if (rc) return rc;
  }
  return vis_layer2(d, blob, w, n_rows, out, st);
}

extern "C" int vfr_visual_embed_split(const float* seg, const float* ctx, const int32_t* vid_off, int64_t n_clips,
                                      int64_t n_videos, int feat_dim, const void* packed, int hid, int dim, void* workspace,
                                      float* out, vfr_stream_t stream) {
  VFR_REQUIRE(seg && ctx && vid_off && packed && workspace && out, VFR_ERR_INVALID, "vfr_visual_embed_split: null pointer");
  VFR_REQUIRE(n_clips > 0 && n_clips < (int64_t(1) << 31) - 256 && n_videos > 0 && n_videos <= n_clips && feat_dim > 0 &&
                  hid > 0 && dim > 0,
              VFR_ERR_INVALID, "vfr_visual_embed_split: bad shape");
  const VisDims d = vis_dims(feat_dim, hid, dim);
  const VisBlob blob = vis_blob(const_cast<void*>(packed), d);
  const VisWs w = vis_ws(workspace, d, n_clips, n_videos, true);
  cudaStream_t st = (cudaStream_t)stream;
  int rc;
  vis_clip_meta_kernel<<<(unsigned)((n_videos + 255) / 256), 256, 0, st>>>(vid_off, n_videos, w.clip_vid, w.tef);
  rc = check_launch("vis_clip_meta_kernel");
  if (rc) return rc;
  vis_split_rows_auto_kernel<<<(unsigned)((n_clips + 7) / 8), 256, 0, st>>>(seg, n_clips, 1, 0, 0, feat_dim, feat_dim, 0, 0, d.Fp,
                                                                            2 * (int64_t)d.Fp, d.Fp, w.row_exp, w.a1);
  vis_split_rows_auto_kernel<<<(unsigned)((n_videos + 7) / 8), 256, 0, st>>>(ctx, n_videos, 1, 0, 0, feat_dim, feat_dim, 0, 0, d.Fp,
                                                                             2 * (int64_t)d.Fp, d.Fp, w.vid_exp, w.ac);
  rc = check_launch("vis_split_rows_auto_kernel");
  if (rc) return rc;
  {
    // once per VIDEO: cvec = ctx . W1[:, F:2F]^T + b1   (the K-segment [Fp, 2 Fp) of the packed W1)
    const void* a[1] = {w.ac};
    const void* b[1] = {blob.w1 + d.Fp};
    EpiVisOut epi{w.cvec, hid, hid, w.vid_exp, 1, blob.exps, blob.b1};
    const int fk = vis_flush_k();
    rc = launch_gemm_tc(a, b, 1, (int)n_videos, hid, d.Fp, 2 * (int64_t)d.Fp, 2 * (int64_t)d.K1, epi, st, nullptr, true, d.Fp,
                        d.K1, fk, fk ? w.flush : nullptr, d.N1);
    if (rc) return rc;
  }
  {
    // per clip: seg . W1[:, :F]^T + cvec[video] + tef . W1[:, 2F:]^T -> relu
    const void* a[1] = {w.a1};
    const void* b[1] = {blob.w1};
    EpiVis1 epi{w.hidden, hid, w.row_exp, 1, blob.exps, blob.b1, w.cvec, w.clip_vid, reinterpret_cast<const float*>(w.tef), 2,
                blob.w1t};
    const int fk = vis_flush_k();
    rc = launch_gemm_tc(a, b, 1, (int)n_clips, hid, d.Fp, 2 * (int64_t)d.Fp, 2 * (int64_t)d.K1, epi, st, nullptr, true, d.Fp,
                        d.K1, fk, fk ? w.flush : nullptr, d.N1);
    if (rc) return rc;
  }
  return vis_layer2(d, blob, w, n_clips, out, st);
}
