// K1: frame -> segment pooling (+ whole-video context feature), L2-normalised with +1e-5.
// Replaces reference model/data.py:142-188 (CustomDataset.load_video_features): the .npy branch
// (:163-181, avg or max over 25-frame windows, ragged last window, context = pool over ALL frames)
// and the "preprocessed .h5" branch (:144-161, six windows of the first 150 rows, drop an all-zero
// 6th segment, context = mean of the un-normalised segment means).
//
// HBM-bound streaming kernel: the frames (2.46 MB per 150-frame video) are read exactly once.
//   pass A  grid (video, 1024-column chunk): each thread owns 4 adjacent columns (128-bit coalesced
//           loads, 5 frames in flight), walks the frames in order (NumPy's axis-0 reduction order,
//           fp32), writes the un-normalised pooled rows and its chunk's partial sums of squares;
//   pass B  same grid over the small pooled rows (L2-resident): fixed-order sum of the partials ->
//           norm -> x / (|x| + 1e-5).  No atomics: results are run-to-run deterministic.
#include "vfr_common.cuh"
#include <math_constants.h>

namespace vfr {

constexpr int P_THREADS = 256;

struct PoolParams {
  const float* frames;       // all videos' frames concatenated [sum F, dim]
  const int64_t* frame_off;  // [V+1]
  int dim;
  int window;       // 25
  int mode;         // 0 = avg, 1 = max, 2 = preprocessed-h5 variant
  float* seg;       // [V, seg_stride, dim]
  float* ctx;       // [V, dim]
  int32_t* n_seg;   // [V]
  int seg_stride;
  float* sqpart;    // [V, n_chunks, seg_stride + 1]
  int n_chunks;
};

__device__ __forceinline__ float4 f4_add(float4 a, float4 b) { return make_float4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w); }
__device__ __forceinline__ float4 f4_max(float4 a, float4 b) {
  return make_float4(fmaxf(a.x, b.x), fmaxf(a.y, b.y), fmaxf(a.z, b.z), fmaxf(a.w, b.w));
}
__device__ __forceinline__ float4 f4_div(float4 a, float d) {
  return make_float4(__fdiv_rn(a.x, d), __fdiv_rn(a.y, d), __fdiv_rn(a.z, d), __fdiv_rn(a.w, d));
}
__device__ __forceinline__ float f4_sq(float4 a) { return a.x * a.x + a.y * a.y + a.z * a.z + a.w * a.w; }

__device__ __forceinline__ float block_sum1(float v, float* red) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  v = warp_sum(v);
  __syncthreads();
  if (lane == 0) red[warp] = v;
  __syncthreads();
  float s = 0.f;
#pragma unroll
  for (int w = 0; w < P_THREADS / 32; ++w) s += red[w];
  return s;
}

__device__ __forceinline__ int video_segments(const PoolParams& p, int F) {
  return p.mode == 2 ? 6 : (F + p.window - 1) / p.window;
}

__global__ void __launch_bounds__(P_THREADS) pool_pass_a(const PoolParams p) {
  __shared__ float red[P_THREADS / 32];
  const int v = blockIdx.x, chunk = blockIdx.y;
  const int64_t f0 = p.frame_off[v];
  int F = (int)(p.frame_off[v + 1] - f0);
  if (p.mode == 2) F = min(F, 6 * p.window);
  const int n = min(video_segments(p, F), p.seg_stride);
  const int dim4 = p.dim >> 2;
  const int c4 = chunk * P_THREADS + threadIdx.x;
  const bool active = c4 < dim4;
  const float4* base = reinterpret_cast<const float4*>(p.frames + f0 * p.dim) + (active ? c4 : 0);
  const bool is_max = p.mode == 1;
  const float4 ident = is_max ? make_float4(-CUDART_INF_F, -CUDART_INF_F, -CUDART_INF_F, -CUDART_INF_F)
                              : make_float4(0.f, 0.f, 0.f, 0.f);
  float4 tot = ident;  // avg: running sum over ALL frames; max: running max; h5: sum of segment means
  float* sq = p.sqpart + ((int64_t)v * p.n_chunks + chunk) * (p.seg_stride + 1);

  for (int s = 0; s < n; ++s) {
    const int a = s * p.window, b = min(a + p.window, F);
    float4 acc = ident;
    if (active) {
      int f = a;
      for (; f + 5 <= b; f += 5) {
        const float4 x0 = __ldg(base + (int64_t)(f + 0) * dim4);
        const float4 x1 = __ldg(base + (int64_t)(f + 1) * dim4);
        const float4 x2 = __ldg(base + (int64_t)(f + 2) * dim4);
        const float4 x3 = __ldg(base + (int64_t)(f + 3) * dim4);
        const float4 x4 = __ldg(base + (int64_t)(f + 4) * dim4);
        if (is_max) {
          acc = f4_max(f4_max(f4_max(f4_max(f4_max(acc, x0), x1), x2), x3), x4);
        } else {
          acc = f4_add(f4_add(f4_add(f4_add(f4_add(acc, x0), x1), x2), x3), x4);
          if (p.mode == 0) tot = f4_add(f4_add(f4_add(f4_add(f4_add(tot, x0), x1), x2), x3), x4);
        }
      }
      for (; f < b; ++f) {
        const float4 x = __ldg(base + (int64_t)f * dim4);
        if (is_max) acc = f4_max(acc, x);
        else {
          acc = f4_add(acc, x);
          if (p.mode == 0) tot = f4_add(tot, x);
        }
      }
    }
    float4 val;
    if (b <= a) val = make_float4(0.f, 0.f, 0.f, 0.f);  // h5: window beyond the clip stays zero
    else if (is_max) val = acc;
    else val = f4_div(acc, (float)(b - a));
    if (is_max) tot = f4_max(tot, val);
    if (p.mode == 2) tot = f4_add(tot, val);
    if (active) *(reinterpret_cast<float4*>(p.seg + ((int64_t)v * p.seg_stride + s) * p.dim) + c4) = val;
    const float part = block_sum1(active ? f4_sq(val) : 0.f, red);
    if (threadIdx.x == 0) sq[s] = part;
  }
  float4 cval = tot;                                     // max: done; h5: divided by kept count in pass B
  if (p.mode == 0) cval = f4_div(tot, (float)F);
  if (active) *(reinterpret_cast<float4*>(p.ctx + (int64_t)v * p.dim) + c4) = cval;
  const float part = block_sum1(active ? f4_sq(cval) : 0.f, red);
  if (threadIdx.x == 0) sq[p.seg_stride] = part;
}

__global__ void __launch_bounds__(P_THREADS) pool_pass_b(const PoolParams p) {
  __shared__ float s_div[VFR_MAX_SEG + 1];
  __shared__ int s_n;
  const int v = blockIdx.x, chunk = blockIdx.y;
  int F = (int)(p.frame_off[v + 1] - p.frame_off[v]);
  if (p.mode == 2) F = min(F, 6 * p.window);
  const int n = min(video_segments(p, F), p.seg_stride);
  if (threadIdx.x <= p.seg_stride) {
    const int s = threadIdx.x;   // s == seg_stride is the context row
    float tot = 0.f;
    if (s < n || s == p.seg_stride)
      for (int c = 0; c < p.n_chunks; ++c) tot += p.sqpart[((int64_t)v * p.n_chunks + c) * (p.seg_stride + 1) + s];
    s_div[s] = tot;              // sum of squares
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    int n_eff = n;
    float ctx_scale = 1.f;
    if (p.mode == 2) {
      // data.py:155-158: drop an all-zero 6th segment, context = mean of the kept segment means
      if (s_div[5] == 0.f) n_eff = 5;
      ctx_scale = (float)n_eff;
    }
    s_n = n_eff;
    for (int s = 0; s < n; ++s) s_div[s] = __fadd_rn(__fsqrt_rn(s_div[s]), VFR_NORM_EPS);
    // context row holds the pooled value (avg, max) or the un-divided sum of segment means (h5):
    // x/k / (|x|/k + eps) == x / (|x| + k*eps)
    s_div[p.seg_stride] = __fadd_rn(__fsqrt_rn(s_div[p.seg_stride]), __fmul_rn(VFR_NORM_EPS, ctx_scale));
    if (chunk == 0) p.n_seg[v] = n_eff;
  }
  __syncthreads();
  const int dim4 = p.dim >> 2;
  const int c4 = chunk * P_THREADS + threadIdx.x;
  if (c4 >= dim4) return;
  const int n_eff = s_n;
  for (int s = 0; s < p.seg_stride; ++s) {
    float4* ptr = reinterpret_cast<float4*>(p.seg + ((int64_t)v * p.seg_stride + s) * p.dim) + c4;
    if (s < n_eff) *ptr = f4_div(*ptr, s_div[s]);
    else *ptr = make_float4(0.f, 0.f, 0.f, 0.f);
  }
  float4* cptr = reinterpret_cast<float4*>(p.ctx + (int64_t)v * p.dim) + c4;
  *cptr = f4_div(*cptr, s_div[p.seg_stride]);
}

}  // namespace vfr

using namespace vfr;

extern "C" size_t vfr_segment_pool_bytes(int64_t n_videos, int dim, int seg_stride) {
  if (n_videos <= 0 || dim <= 0 || seg_stride <= 0) return 0;
  const size_t chunks = (size_t)(dim / 4 + P_THREADS - 1) / P_THREADS;
  return (size_t)n_videos * chunks * (seg_stride + 1) * sizeof(float);
}

extern "C" int vfr_segment_pool(const float* frames, const int64_t* frame_off, int64_t n_videos, int dim, int window,
                                int mode, float* seg, int seg_stride, float* ctx, int32_t* n_seg, void* workspace,
                                vfr_stream_t stream) {
  VFR_REQUIRE(frames && frame_off && seg && ctx && n_seg && workspace, VFR_ERR_INVALID, "vfr_segment_pool: null pointer");
  VFR_REQUIRE(n_videos > 0 && n_videos < (int64_t(1) << 31), VFR_ERR_INVALID, "vfr_segment_pool: n_videos");
  VFR_REQUIRE(dim > 0 && dim % 4 == 0, VFR_ERR_UNSUPPORTED, "vfr_segment_pool: dim=%d must be a multiple of 4", dim);
  VFR_REQUIRE(window > 0 && mode >= 0 && mode <= 2 && seg_stride >= 1 && seg_stride <= VFR_MAX_SEG, VFR_ERR_INVALID,
              "vfr_segment_pool: bad window/mode/seg_stride");
  VFR_REQUIRE(mode != 2 || seg_stride >= 6, VFR_ERR_INVALID, "vfr_segment_pool: h5 mode needs seg_stride >= 6");
  const int chunks = (dim / 4 + P_THREADS - 1) / P_THREADS;
  VFR_REQUIRE(chunks <= 65535, VFR_ERR_UNSUPPORTED, "vfr_segment_pool: dim too large");
  PoolParams p{frames, frame_off, dim, window, mode, seg, ctx, n_seg, seg_stride,
               reinterpret_cast<float*>(workspace), chunks};
  cudaStream_t st = (cudaStream_t)stream;
  dim3 grid((unsigned)n_videos, chunks);
  pool_pass_a<<<grid, P_THREADS, 0, st>>>(p);
  int rc = check_launch("pool_pass_a");
  if (rc) return rc;
  pool_pass_b<<<grid, P_THREADS, 0, st>>>(p);
  return check_launch("pool_pass_b");
}
