// K6: ranking loss forward + backward.
// Replaces reference model/main.py:214-232 (Trainer.ranking_loss) and its autograd graph:
//   per sample i:  c_p = mean_r ||posit_r - lang_i + 1e-6||   (rows with maskp == i)
//                  c_n = mean_r ||intra_r - lang_i + 1e-6||   (rows with maskn == i)
//                  c_t = mean_r ||inter_r - lang_i + 1e-6||   (rows with maskp == i)
//   loss = sum_i relu(c_p - c_n + b) + lamb * relu(c_p - c_t + b)          (a SUM, main.py:231)
// optional pre-normalisation x / (|x| + 1e-5) of all four inputs (main.py:219-223).
// Row runs are contiguous and ascending (model/data.py:347-348) but the kernels only rely on the
// mask values.  Tiny problem (~360 x 100 floats): one warp per row for the distances, a single CTA
// for the per-sample reduction (fixed order -> deterministic), latency-bound by design.
#include "vfr_common.cuh"

namespace vfr {

struct LossParams {
  const float* x[3];      // posit, intra, inter  [R_s, D]
  const int64_t* mask[3]; // maskp, maskn, maskp
  int rows[3];
  const float* lang;      // [B, D]
  int n_samples;
  int dim;
  int normalize;
  float b, lamb;
  // workspace
  float* dist[3];         // [R_s]
  float* inv[4];          // 1 / (|x| + eps) per row of posit/intra/inter/lang (normalize only)
  float* cost;            // [3, B] mean distances
  float* coef;            // [3, B] dL/dc_s divided by the run length
  float* loss;            // [1]
  // backward outputs
  float* grad[4];         // d posit, d intra, d inter, d lang
  const float* grad_out;  // scalar upstream gradient (device), may be null (= 1)
};

__device__ __forceinline__ float row_inv_norm(const float* r, int dim, int lane) {
  float ss = 0.f;
  for (int k = lane; k < dim; k += 32) ss = __fmaf_rn(r[k], r[k], ss);
  ss = warp_sum(ss);
  return __fdiv_rn(1.f, __fadd_rn(__fsqrt_rn(ss), VFR_NORM_EPS));
}

// phase 0: inverse norms (normalize only).  one warp per row of the 4 tensors
__global__ void loss_norm_kernel(const LossParams p) {
  const int lane = threadIdx.x & 31;
  int w = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  for (int s = 0; s < 4; ++s) {
    const int rows = s < 3 ? p.rows[s] : p.n_samples;
    if (w < rows) {
      const float* r = (s < 3 ? p.x[s] : p.lang) + (int64_t)w * p.dim;
      const float inv = row_inv_norm(r, p.dim, lane);
      if (lane == 0) p.inv[s][w] = inv;
      return;
    }
    w -= rows;
  }
}

// phase 1: clip distances, one warp per row
__global__ void loss_dist_kernel(const LossParams p) {
  const int lane = threadIdx.x & 31;
  int w = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  for (int s = 0; s < 3; ++s) {
    if (w < p.rows[s]) {
      const int64_t i = p.mask[s][w];
      // a row whose sample id is outside [0, n_samples) is never selected by the reference's `mask == i` (main.py:226-230)
      if (i < 0 || i >= p.n_samples) { if (lane == 0) p.dist[s][w] = 0.f; return; }
      const float* xr = p.x[s] + (int64_t)w * p.dim;
      const float* lr = p.lang + i * p.dim;
      const float sx = p.normalize ? p.inv[s][w] : 1.f;
      const float sl = p.normalize ? p.inv[3][i] : 1.f;
      float ss = 0.f;
      for (int k = lane; k < p.dim; k += 32) {
        const float d = __fadd_rn(__fsub_rn(xr[k] * sx, lr[k] * sl), VFR_PAIRWISE_EPS);
        ss = __fmaf_rn(d, d, ss);
      }
      ss = warp_sum(ss);
      if (lane == 0) p.dist[s][w] = __fsqrt_rn(ss);
      return;
    }
    w -= p.rows[s];
  }
}

// phase 2: per-sample means, hinges, loss (single CTA)
__global__ void loss_reduce_kernel(const LossParams p) {
  const int B = p.n_samples;
  for (int i = threadIdx.x; i < B; i += blockDim.x) {
    float c[3];
    int cnt[3];
    for (int s = 0; s < 3; ++s) {
      float sum = 0.f;
      int n = 0;
      for (int r = 0; r < p.rows[s]; ++r)
        if (p.mask[s][r] == i) { sum = __fadd_rn(sum, p.dist[s][r]); ++n; }
      cnt[s] = n;
      c[s] = __fdiv_rn(sum, (float)n);   // empty run -> nan, as torch's mean of an empty tensor
      p.cost[s * B + i] = c[s];
    }
    const float h1 = c[0] - c[1] + p.b, h2 = c[0] - c[2] + p.b;
    // NaN (an empty run, diverged embeddings) must reach the loss as F.relu(nan) = nan does in the reference
    // (main.py:231): fmaxf(nan, 0) would return 0 and hide the divergence
    const float a1 = h1 > 0.f ? 1.f : (h1 != h1 ? h1 : 0.f), a2 = h2 > 0.f ? 1.f : (h2 != h2 ? h2 : 0.f);
    p.coef[0 * B + i] = __fdiv_rn(a1 + p.lamb * a2, (float)cnt[0]);
    p.coef[1 * B + i] = __fdiv_rn(-a1, (float)cnt[1]);
    p.coef[2 * B + i] = __fdiv_rn(-p.lamb * a2, (float)cnt[2]);
    const float r1 = h1 > 0.f ? h1 : (h1 != h1 ? h1 : 0.f), r2 = h2 > 0.f ? h2 : (h2 != h2 ? h2 : 0.f);
    p.cost[3 * B + i] = r1 + p.lamb * r2;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    float loss = 0.f;   // sequential sum over samples, as the python loop does
    for (int i = 0; i < B; ++i) loss = __fadd_rn(loss, p.cost[3 * B + i]);
    *p.loss = loss;
  }
}

// backward, one warp per row of posit/intra/inter: dx and the row's contribution to d lang
__global__ void loss_bwd_rows_kernel(const LossParams p) {
  const int lane = threadIdx.x & 31;
  int w = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const float go = p.grad_out ? *p.grad_out : 1.f;
  const int B = p.n_samples;
  for (int s = 0; s < 3; ++s) {
    if (w < p.rows[s]) {
      const int64_t i = p.mask[s][w];
      float* gx = p.grad[s] + (int64_t)w * p.dim;
      if (i < 0 || i >= p.n_samples) {       // row of no sample: no gradient (and no out-of-range access)
        for (int k = lane; k < p.dim; k += 32) gx[k] = 0.f;
        return;
      }
      const float* xr = p.x[s] + (int64_t)w * p.dim;
      const float* lr = p.lang + i * p.dim;
      const float sx = p.normalize ? p.inv[s][w] : 1.f;
      const float sl = p.normalize ? p.inv[3][i] : 1.f;
      const float d = p.dist[s][w];
      const float g = (d > 0.f || d != d) ? go * p.coef[s * B + i] / d : 0.f;   // dL/dd / d (NaN propagates)
      // u_k = g * (xhat_k - lhat_k + eps) is the gradient wrt xhat (and minus that wrt lhat)
      float dot = 0.f;   // sum_k u_k * xhat_k  (needed to back-propagate through the normalisation)
      for (int k = lane; k < p.dim; k += 32) {
        const float xh = xr[k] * sx;
        const float u = g * __fadd_rn(__fsub_rn(xh, lr[k] * sl), VFR_PAIRWISE_EPS);
        dot = __fmaf_rn(u, xh, dot);
      }
      dot = warp_sum(dot);
      // xhat = x * sx with sx = 1/(|x|+eps):  dx = sx * (u - xhat * (u . xhat) * |x| * sx) ;  |x| = 1/sx - eps
      const float nx = p.normalize ? (__fdiv_rn(1.f, sx) - VFR_NORM_EPS) : 0.f;
      for (int k = lane; k < p.dim; k += 32) {
        const float xh = xr[k] * sx;
        const float u = g * __fadd_rn(__fsub_rn(xh, lr[k] * sl), VFR_PAIRWISE_EPS);
        float gxk = u;
        if (p.normalize) gxk = (nx > 0.f) ? sx * (u - xh * dot / (sx * nx)) : sx * u;
        gx[k] = gxk;
        atomicAdd(p.grad[3] + i * p.dim + k, -u);   // accumulates d lhat (normalised space)
      }
      return;
    }
    w -= p.rows[s];
  }
}

// backward through the language normalisation: grad[3] holds d lhat; convert to d lang in place
__global__ void loss_bwd_lang_norm_kernel(const LossParams p) {
  const int lane = threadIdx.x & 31;
  const int i = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (i >= p.n_samples) return;
  const float sl = p.inv[3][i];
  const float* lr = p.lang + (int64_t)i * p.dim;
  float* g = p.grad[3] + (int64_t)i * p.dim;
  float dot = 0.f;
  for (int k = lane; k < p.dim; k += 32) dot = __fmaf_rn(g[k], lr[k] * sl, dot);
  dot = warp_sum(dot);
  const float nl = __fdiv_rn(1.f, sl) - VFR_NORM_EPS;
  for (int k = lane; k < p.dim; k += 32) {
    const float lh = lr[k] * sl;
    g[k] = (nl > 0.f) ? sl * (g[k] - lh * dot / (sl * nl)) : sl * g[k];
  }
}

}  // namespace vfr

using namespace vfr;

extern "C" size_t vfr_ranking_loss_bytes(int rows_posit, int rows_intra, int rows_inter, int n_samples) {
  if (rows_posit < 0 || rows_intra < 0 || rows_inter < 0 || n_samples <= 0) return 0;
  const size_t r = (size_t)rows_posit + rows_intra + rows_inter;
  return (2 * r + 8 * (size_t)n_samples + 8) * sizeof(float);
}

static int fill_loss(LossParams& p, const float* posit, const float* intra, const float* inter, const float* lang,
                     const int64_t* maskp, const int64_t* maskn, int rows_posit, int rows_intra, int rows_inter,
                     int n_samples, int dim, int normalize, float b, float lamb, void* workspace) {
  VFR_REQUIRE(posit && intra && inter && lang && maskp && maskn && workspace, VFR_ERR_INVALID, "ranking_loss: null pointer");
  VFR_REQUIRE(rows_posit > 0 && rows_intra > 0 && rows_inter == rows_posit && n_samples > 0 && dim > 0, VFR_ERR_INVALID,
              "ranking_loss: bad shape (posit %d, intra %d, inter %d rows, %d samples)", rows_posit, rows_intra,
              rows_inter, n_samples);
  p = LossParams{};
  p.x[0] = posit; p.x[1] = intra; p.x[2] = inter;
  p.mask[0] = maskp; p.mask[1] = maskn; p.mask[2] = maskp;
  p.rows[0] = rows_posit; p.rows[1] = rows_intra; p.rows[2] = rows_inter;
  p.lang = lang;
  p.n_samples = n_samples;
  p.dim = dim;
  p.normalize = normalize;
  p.b = b;
  p.lamb = lamb;
  float* ws = reinterpret_cast<float*>(workspace);
  for (int s = 0; s < 3; ++s) { p.dist[s] = ws; ws += p.rows[s]; }
  for (int s = 0; s < 3; ++s) { p.inv[s] = ws; ws += p.rows[s]; }
  p.inv[3] = ws; ws += n_samples;
  p.cost = ws; ws += 4 * (size_t)n_samples;
  p.coef = ws; ws += 3 * (size_t)n_samples;
  p.loss = ws;
  return VFR_OK;
}

extern "C" int vfr_ranking_loss_fwd(const float* posit, const float* intra, const float* inter, const float* lang,
                                    const int64_t* maskp, const int64_t* maskn, int rows_posit, int rows_intra,
                                    int rows_inter, int n_samples, int dim, int normalize, float b, float lamb,
                                    void* workspace, float* loss_out, vfr_stream_t stream) {
  LossParams p;
  int rc = fill_loss(p, posit, intra, inter, lang, maskp, maskn, rows_posit, rows_intra, rows_inter, n_samples, dim,
                     normalize, b, lamb, workspace);
  if (rc) return rc;
  VFR_REQUIRE(loss_out, VFR_ERR_INVALID, "ranking_loss_fwd: null loss_out");
  cudaStream_t st = (cudaStream_t)stream;
  const int rows3 = rows_posit + rows_intra + rows_inter;
  if (normalize) {
    loss_norm_kernel<<<(rows3 + n_samples + 7) / 8, 256, 0, st>>>(p);
    rc = check_launch("loss_norm_kernel");
    if (rc) return rc;
  }
  loss_dist_kernel<<<(rows3 + 7) / 8, 256, 0, st>>>(p);
  rc = check_launch("loss_dist_kernel");
  if (rc) return rc;
  loss_reduce_kernel<<<1, 256, 0, st>>>(p);
  rc = check_launch("loss_reduce_kernel");
  if (rc) return rc;
  VFR_CUDA(cudaMemcpyAsync(loss_out, p.loss, sizeof(float), cudaMemcpyDeviceToDevice, st));
  return VFR_OK;
}

extern "C" int vfr_ranking_loss_bwd(const float* posit, const float* intra, const float* inter, const float* lang,
                                    const int64_t* maskp, const int64_t* maskn, int rows_posit, int rows_intra,
                                    int rows_inter, int n_samples, int dim, int normalize, float b, float lamb,
                                    void* workspace, const float* grad_out, float* grad_posit, float* grad_intra,
                                    float* grad_inter, float* grad_lang, vfr_stream_t stream) {
  LossParams p;
  int rc = fill_loss(p, posit, intra, inter, lang, maskp, maskn, rows_posit, rows_intra, rows_inter, n_samples, dim,
                     normalize, b, lamb, workspace);
  if (rc) return rc;
  VFR_REQUIRE(grad_posit && grad_intra && grad_inter && grad_lang, VFR_ERR_INVALID, "ranking_loss_bwd: null grad");
  p.grad[0] = grad_posit; p.grad[1] = grad_intra; p.grad[2] = grad_inter; p.grad[3] = grad_lang;
  p.grad_out = grad_out;
  cudaStream_t st = (cudaStream_t)stream;
  VFR_CUDA(cudaMemsetAsync(grad_lang, 0, (size_t)n_samples * dim * sizeof(float), st));
  const int rows3 = rows_posit + rows_intra + rows_inter;
  loss_bwd_rows_kernel<<<(rows3 + 7) / 8, 256, 0, st>>>(p);
  rc = check_launch("loss_bwd_rows_kernel");
  if (rc) return rc;
  if (normalize) {
    loss_bwd_lang_norm_kernel<<<(n_samples + 7) / 8, 256, 0, st>>>(p);
    rc = check_launch("loss_bwd_lang_norm_kernel");
  }
  return rc;
}
