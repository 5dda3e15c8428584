// K2 (exact-fp32 path): visual embedding  Linear(in,hid) -> ReLU -> Linear(hid,D).
// Replaces reference model/models.py:21-27,55-56 (eval mode: Dropout is the identity; in train mode
// the Python shim applies the dropout mask, see models.py).
#include "vfr_gemm.cuh"

namespace vfr {

struct EpiBiasAct {
  float* out;
  int ldc;
  int N;
  const float* bias;
  int relu;
  __device__ __forceinline__ void operator()(int, int m, int n0, const float (&v)[4]) const {
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int n = n0 + j;
      if (n < N) {
        float x = v[j] + (bias ? __ldg(bias + n) : 0.f);
        if (relu) x = fmaxf(x, 0.f);
        out[(int64_t)m * ldc + n] = x;
      }
    }
  }
};

}  // namespace vfr

using namespace vfr;

extern "C" int vfr_linear(const float* x, int64_t n_rows, int in_dim, int ldx, const float* w, const float* bias,
                          int out_dim, int relu, float* out, int ldo, vfr_stream_t stream) {
  VFR_REQUIRE(x && w && out, VFR_ERR_INVALID, "vfr_linear: null pointer");
  VFR_REQUIRE(n_rows >= 0 && n_rows < (int64_t(1) << 31) && in_dim > 0 && out_dim > 0 && ldx >= in_dim && ldo >= out_dim,
              VFR_ERR_INVALID, "vfr_linear: bad shape");
  if (n_rows == 0) return VFR_OK;
  GemmBatch ops{};
  ops.op[0] = GemmOperand{x, w};
  EpiBiasAct epi{out, ldo, out_dim, bias, relu};
  return launch_sgemm_nt(ops, 1, ldx, in_dim, (int)n_rows, out_dim, in_dim, epi, (cudaStream_t)stream);
}

extern "C" int vfr_visual_embed(const float* x, int64_t n_rows, int in_dim, const float* w1, const float* b1, int hid,
                                const float* w2, const float* b2, int dim, float* hidden, float* out,
                                vfr_stream_t stream) {
  VFR_REQUIRE(hidden, VFR_ERR_INVALID, "vfr_visual_embed: null hidden workspace");
  int rc = vfr_linear(x, n_rows, in_dim, in_dim, w1, b1, hid, 1, hidden, hid, stream);
  if (rc) return rc;
  return vfr_linear(hidden, n_rows, hid, hid, w2, b2, dim, 0, out, dim, stream);
}
