// Training-batch construction on the device:
//  * vfr_sample_negatives - the per-query draws of the reference's CustomBatchSampler.__iter__ (model/data.py:275-337): a
//    positive annotation, an intra-video negative (same-length or any non-annotated segment) and an inter-video negative
//    (a random other video with enough segments), one thread per query of the epoch, counter-based RNG (a pure function of
//    (seed, epoch, query, draw): reproducible, order-free).  Same DISTRIBUTIONS as the reference; not its Mersenne-Twister
//    stream, which is inherently sequential.
//  * vfr_gather_clip_rows - assembles the [segment | context | tef] rows of model/data.py:204-213 for a list of
//    (video, clip) pairs straight from the pooled features resident in HBM (K1's outputs).
//  * vfr_moment_pool - the MCN-style POOLED-MOMENT features (an additional, non-reference scoring variant, north-star
//    item (1)): the mean segment feature of every candidate moment (s, e) of every video from per-column prefix sums.
#include "vfr_common.cuh"
#include <algorithm>

namespace vfr {

__device__ __forceinline__ unsigned long long mix64(unsigned long long x) {   // splitmix64 finaliser
  x += 0x9e3779b97f4a7c15ull;
  x = (x ^ (x >> 30)) * 0xbf58476d1ce4e5b9ull;
  x = (x ^ (x >> 27)) * 0x94d049bb133111ebull;
  return x ^ (x >> 31);
}
struct Rng {
  unsigned long long key;
  unsigned ctr;
  __device__ unsigned below(unsigned n) {      // uniform in [0, n), n > 0 (multiply-shift; bias < 2^-32 n)
    const unsigned long long r = mix64(key + 0x632be59bd9b4e019ull * (++ctr));
    return (unsigned)(((r >> 32) * (unsigned long long)n) >> 32);
  }
};

struct SampleParams {
  const int32_t* times;     // [Q, A, 2] inclusive (start, end), absent annotators (-1, -1)
  int n_annot;
  const int32_t* q_video;   // [Q]
  const int32_t* nseg;      // [V]
  int64_t n_queries, n_videos;
  int same_length;
  unsigned long long seed, epoch;
  int32_t* out;             // [Q, 8] video_pos, video_neg, start_t, end_t, start_tn, end_tn, status, 0
};

__global__ void sample_negatives_kernel(const SampleParams p) {
  const int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= p.n_queries) return;
  Rng rng{mix64(p.seed ^ mix64(p.epoch * 0x100000001b3ull + (unsigned long long)q)), 0u};
  const int32_t* t = p.times + q * p.n_annot * 2;
  const int vp = p.q_video[q];
  const int n = p.nseg[vp];
  int status = 0;
  // positive: random.choice over the annotations with end <= num_segments (data.py:262-266,289)
  int cnt = 0;
  for (int a = 0; a < p.n_annot; ++a) cnt += (t[2 * a] >= 0 && t[2 * a + 1] <= n) ? 1 : 0;
  int st = 0, en = 0;
  if (cnt == 0) status = 3;
  else {
    int pick = (int)rng.below((unsigned)cnt);
    for (int a = 0; a < p.n_annot; ++a)
      if (t[2 * a] >= 0 && t[2 * a + 1] <= n && pick-- == 0) { st = t[2 * a]; en = t[2 * a + 1]; }
  }
  // intra-video negative (data.py:299-313)
  int sn = 0, enn = 0;
  if (p.same_length) {
    const int len = en - st;
    const int c = (n - len) - 1;                 // segments of that length, minus the positive itself
    if (c <= 0) status = status ? status : 1;    // random.choice([]) raises IndexError in the reference
    else {
      const int r = (int)rng.below((unsigned)c);
      sn = r < st ? r : r + 1;
      enn = sn + len;
    }
  } else {
    int c = 0;
    for (int s = 0; s < n; ++s)
      for (int e = s; e < n; ++e) {
        bool annotated = false;                  // IoU == 1 with some annotator <=> identical segment
        for (int a = 0; a < p.n_annot; ++a) annotated = annotated || (t[2 * a] == s && t[2 * a + 1] == e);
        c += annotated ? 0 : 1;
      }
    if (c <= 0) status = status ? status : 1;
    else {
      int r = (int)rng.below((unsigned)c);
      for (int s = 0; s < n && r >= 0; ++s)
        for (int e = s; e < n && r >= 0; ++e) {
          bool annotated = false;
          for (int a = 0; a < p.n_annot; ++a) annotated = annotated || (t[2 * a] == s && t[2 * a + 1] == e);
          if (!annotated && r-- == 0) { sn = s; enn = e; }
        }
    }
  }
  // inter-video negative: a random OTHER video with at least end_t + 1 segments (data.py:317-323)
  int vn = -1;
  for (int tries = 0; tries < 64 && vn < 0; ++tries) {
    const int v = (int)rng.below((unsigned)p.n_videos);
    if (v != vp && p.nseg[v] >= en + 1) vn = v;
  }
  if (vn < 0) {                                   // rare (few long videos): uniform among the valid ones by counting
    int c = 0;
    for (int64_t v = 0; v < p.n_videos; ++v) c += (v != vp && p.nseg[v] >= en + 1) ? 1 : 0;
    if (c == 0) status = status ? status : 2;
    else {
      int r = (int)rng.below((unsigned)c);
      for (int64_t v = 0; v < p.n_videos && vn < 0; ++v)
        if (v != vp && p.nseg[v] >= en + 1 && r-- == 0) vn = (int)v;
    }
  }
  int32_t* o = p.out + q * 8;
  o[0] = vp; o[1] = vn; o[2] = st; o[3] = en; o[4] = sn; o[5] = enn; o[6] = status; o[7] = 0;
}

// one block per output row: x[r] = [seg[vid_off[v] + c] | ctx[v] | c / n, (c + 1) / n]
__global__ void gather_clip_rows_kernel(const float* __restrict__ seg, const float* __restrict__ ctx,
                                        const int32_t* __restrict__ vid_off, const int32_t* __restrict__ row_video,
                                        const int32_t* __restrict__ row_clip, int feat, float* __restrict__ out) {
  const int64_t r = blockIdx.x;
  const int v = row_video[r], c = row_clip[r];
  const int c0 = vid_off[v], n = vid_off[v + 1] - c0;
  const float4* s4 = reinterpret_cast<const float4*>(seg + (int64_t)(c0 + c) * feat);
  const float4* c4 = reinterpret_cast<const float4*>(ctx + (int64_t)v * feat);
  float* o = out + r * (2 * (int64_t)feat + 2);
  // (rows of 2F + 2 floats are only 8-byte aligned: 64-bit stores)
  for (int i = threadIdx.x; i < feat / 4; i += blockDim.x) {
    const float4 a = __ldg(s4 + i), b = __ldg(c4 + i);
    float2* oa = reinterpret_cast<float2*>(o + 4 * i);
    float2* ob = reinterpret_cast<float2*>(o + feat + 4 * i);
    oa[0] = make_float2(a.x, a.y); oa[1] = make_float2(a.z, a.w);
    ob[0] = make_float2(b.x, b.y); ob[1] = make_float2(b.z, b.w);
  }
  if (threadIdx.x == 0) {
    o[2 * feat] = __fdiv_rn((float)c, (float)n);
    o[2 * feat + 1] = __fdiv_rn((float)(c + 1), (float)n);
  }
}

// pooled-moment features: out[mom_off[v] + m][j] = mean_{c = s..e} seg[vid_off[v] + c][j], m = moment_index(n, s, e).
// One thread per (video, 4 consecutive columns): the prefix sums over the <= 32 clips live in shared memory
// ([clip][thread] float4: conflict-free), loads and stores are 128-bit and coalesced along the feature dimension; every
// moment is then one subtraction and one multiply per column.
__global__ void moment_pool_kernel(const float* __restrict__ seg, const int32_t* __restrict__ vid_off,
                                   const int64_t* __restrict__ mom_off, int feat, float* __restrict__ out) {
  extern __shared__ float4 mp_pre[];               // [(n_max + 1)][blockDim.x]
  const int64_t v = blockIdx.y;
  const int j4 = blockIdx.x * blockDim.x + threadIdx.x;
  if (j4 >= feat / 4) return;
  const int c0 = vid_off[v], n = vid_off[v + 1] - c0;
  const int64_t m0 = mom_off[v];
  float4 run = make_float4(0.f, 0.f, 0.f, 0.f);
  mp_pre[threadIdx.x] = run;
  for (int c = 0; c < n; ++c) {
    const float4 x = __ldg(reinterpret_cast<const float4*>(seg + (int64_t)(c0 + c) * feat) + j4);
    run = make_float4(run.x + x.x, run.y + x.y, run.z + x.z, run.w + x.w);
    mp_pre[(c + 1) * blockDim.x + threadIdx.x] = run;
  }
  for (int s = 0; s < n; ++s) {
    const float4 b = mp_pre[s * blockDim.x + threadIdx.x];
    for (int e = s; e < n; ++e) {
      const float inv = __fdiv_rn(1.f, (float)(e - s + 1));
      const float4 a = mp_pre[(e + 1) * blockDim.x + threadIdx.x];
      reinterpret_cast<float4*>(out + (m0 + moment_index(n, s, e)) * feat)[j4] =
          make_float4((a.x - b.x) * inv, (a.y - b.y) * inv, (a.z - b.z) * inv, (a.w - b.w) * inv);
    }
  }
}

}  // namespace vfr

using namespace vfr;

extern "C" int vfr_sample_negatives(const int32_t* times, int n_annot, const int32_t* q_video, const int32_t* nseg,
                                    int64_t n_queries, int64_t n_videos, int same_length, uint64_t seed, uint64_t epoch,
                                    int32_t* out, vfr_stream_t stream) {
  VFR_REQUIRE(times && q_video && nseg && out, VFR_ERR_INVALID, "vfr_sample_negatives: null pointer");
  VFR_REQUIRE(n_annot >= 1 && n_queries > 0 && n_videos > 0 && n_videos < (int64_t(1) << 31), VFR_ERR_INVALID,
              "vfr_sample_negatives: bad shape");
  SampleParams p{times, n_annot, q_video, nseg, n_queries, n_videos, same_length, seed, epoch, out};
  sample_negatives_kernel<<<(unsigned)((n_queries + 127) / 128), 128, 0, (cudaStream_t)stream>>>(p);
  return check_launch("sample_negatives_kernel");
}

extern "C" int vfr_gather_clip_rows(const float* seg, const float* ctx, const int32_t* vid_off, const int32_t* row_video,
                                    const int32_t* row_clip, int64_t n_rows, int feat_dim, float* out, vfr_stream_t stream) {
  VFR_REQUIRE(seg && ctx && vid_off && row_video && row_clip && out, VFR_ERR_INVALID, "vfr_gather_clip_rows: null pointer");
  VFR_REQUIRE(n_rows >= 0 && n_rows < (int64_t(1) << 31) && feat_dim > 0 && feat_dim % 4 == 0, VFR_ERR_INVALID,
              "vfr_gather_clip_rows: bad shape (feat_dim %% 4 == 0)");
  if (n_rows == 0) return VFR_OK;
  gather_clip_rows_kernel<<<(unsigned)n_rows, 256, 0, (cudaStream_t)stream>>>(seg, ctx, vid_off, row_video, row_clip, feat_dim, out);
  return check_launch("gather_clip_rows_kernel");
}

extern "C" int vfr_moment_pool(const float* seg, const int32_t* vid_off, const int64_t* mom_off, int64_t n_videos, int n_max,
                               int feat_dim, float* out, vfr_stream_t stream) {
  VFR_REQUIRE(seg && vid_off && mom_off && out, VFR_ERR_INVALID, "vfr_moment_pool: null pointer");
  VFR_REQUIRE(n_videos > 0 && feat_dim > 0 && feat_dim % 4 == 0 && n_max >= 1 && n_max <= VFR_MAX_SEG, VFR_ERR_INVALID,
              "vfr_moment_pool: bad shape (feat_dim %% 4 == 0, n_max <= 32)");
  const int threads = 128;
  const size_t smem = (size_t)(n_max + 1) * threads * sizeof(float4);          // <= 67.6 KB at n_max = 32
  VFR_CUDA(cudaFuncSetAttribute(moment_pool_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  dim3 grid((unsigned)((feat_dim / 4 + threads - 1) / threads), 1);
  for (int64_t v0 = 0; v0 < n_videos; v0 += 65535) {          // grid.y is limited to 65535 videos per launch
    grid.y = (unsigned)std::min<int64_t>(65535, n_videos - v0);
    moment_pool_kernel<<<grid, threads, smem, (cudaStream_t)stream>>>(seg, vid_off + v0, mom_off + v0, feat_dim, out);
    int rc = check_launch("moment_pool_kernel");
    if (rc) return rc;
  }
  return VFR_OK;
}
