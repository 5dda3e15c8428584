// K3 (exact-fp32 path): query embedding  GloVe gather -> (optional length-normalise) -> BiLSTM over
// all L positions (padding included, h0=c0=0) -> [h_fwd(L-1) | h_bwd(0)] -> Linear(2H, D).
// Replaces reference model/models.py:33-48,61-66.
//
// B200-first layout: per direction one buffer hx[L+1][B][H+E]; slot t holds [h_{t-1} | x_t], so one
// recurrent step is ONE GEMM  gates = [h|x] * [W_hh|W_ih]^T  (K = H+E) whose epilogue adds the
// bias, applies the LSTM cell and writes h_t straight into slot t+1 - no separate input-projection
// buffer, no pointwise kernel, and every h_t stays resident for a later backward pass.  Weight rows
// are interleaved (row 4j+g <- gate g of unit j, PyTorch gate order i,f,g,o) so the four gates of a
// unit land in one thread's 4-column epilogue group.  Both directions run in one launch (grid z).
#include "vfr_gemm.cuh"

namespace vfr {

__global__ void lstm_pack_kernel(const float* __restrict__ w_ih, const float* __restrict__ w_hh,
                                 const float* __restrict__ b_ih, const float* __restrict__ b_hh, int H, int E,
                                 float* __restrict__ wcat, float* __restrict__ bias) {
  const int K = H + E;
  const int64_t total = (int64_t)4 * H * K;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int rp = (int)(i / K), k = (int)(i % K);
    const int j = rp >> 2, g = rp & 3;
    const int r = g * H + j;
    wcat[i] = (k < H) ? w_hh[(int64_t)r * H + k] : w_ih[(int64_t)r * E + (k - H)];
    if (k == 0) bias[rp] = b_ih[r] + b_hh[r];
  }
}

// fill the x part of every slot (both directions) and zero h_{-1}; one warp per (b, t)
__global__ void lstm_gather_kernel(const int64_t* __restrict__ tokens, int64_t B, int L, const float* __restrict__ table,
                                   int64_t vocab, const float* __restrict__ length, int E, int H,
                                   float* __restrict__ hx_f, float* __restrict__ hx_b, int* __restrict__ bad_token) {
  const int lane = threadIdx.x & 31;
  const int64_t w = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int K = H + E;
  if (w < B * L) {
    const int64_t b = w / L;
    const int t = (int)(w % L);
    int64_t id = tokens[b * L + t];
    if (id < 0 || id >= vocab) { if (lane == 0) atomicExch(bad_token, 1); id = 0; }
    const float* row = table + id * E;
    float scale = 1.f;
    if (length) {
      float ss = 0.f;
      for (int k = lane; k < E; k += 32) ss = __fmaf_rn(row[k], row[k], ss);
      ss = warp_sum(ss);
      scale = __fadd_rn(__fsqrt_rn(ss), VFR_NORM_EPS);   // |x| + 1e-5
    }
    const float len = length ? length[id] : 1.f;
    // forward direction consumes x_t at step t; backward direction consumes x_{L-1-t} at step t
    float* df = hx_f + ((int64_t)t * B + b) * K + H;
    float* db = hx_b + ((int64_t)(L - 1 - t) * B + b) * K + H;
    for (int k = lane; k < E; k += 32) {
      float v = row[k];
      if (length) v = __fmul_rn(__fdiv_rn(row[k], scale), len);
      df[k] = v;
      db[k] = v;
    }
  }
  // h_{-1} = 0 in slot 0 of both directions
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < B * H; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t b = i / H;
    const int k = (int)(i % H);
    hx_f[b * K + k] = 0.f;
    hx_b[b * K + k] = 0.f;
  }
}

__device__ __forceinline__ float sigmoid_acc(float x) { return __fdiv_rn(1.f, 1.f + expf(-x)); }

struct EpiLstmCell {
  const float* bias[2];   // [4H] interleaved, per direction
  float* c[2];            // [B, H] cell state, per direction (in place)
  float* h_next[2];       // slot t+1 base, per direction: row stride ldh, first H columns
  int ldh;
  int H;
  int first;              // step 0: c_prev = 0 (c buffer not yet initialised)
  __device__ __forceinline__ void operator()(int z, int m, int n0, const float (&v)[4]) const {
    const int j = n0 >> 2;
    if (j >= H) return;
    const float* bz = bias[z];
    const float gi = v[0] + __ldg(bz + n0), gf = v[1] + __ldg(bz + n0 + 1);
    const float gg = v[2] + __ldg(bz + n0 + 2), go = v[3] + __ldg(bz + n0 + 3);
    float* cp = c[z] + (int64_t)m * H + j;
    const float c_prev = first ? 0.f : *cp;
    const float c_new = sigmoid_acc(gf) * c_prev + sigmoid_acc(gi) * tanhf(gg);
    *cp = c_new;
    h_next[z][(int64_t)m * ldh + j] = sigmoid_acc(go) * tanhf(c_new);
  }
};

struct EpiFc {
  float* out;
  int ldc;
  int N;
  const float* bias;
  int accumulate;
  __device__ __forceinline__ void operator()(int, int m, int n0, const float (&v)[4]) const {
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int n = n0 + j;
      if (n < N) {
        float* o = out + (int64_t)m * ldc + n;
        *o = accumulate ? (*o + v[j]) : (v[j] + __ldg(bias + n));
      }
    }
  }
};

}  // namespace vfr

using namespace vfr;

extern "C" size_t vfr_lstm_pack_bytes(int hidden, int emb) {
  if (hidden <= 0 || emb <= 0) return 0;
  return ((size_t)4 * hidden * (hidden + emb) + (size_t)4 * hidden) * sizeof(float);
}

extern "C" int vfr_lstm_pack(const float* w_ih, const float* w_hh, const float* b_ih, const float* b_hh, int hidden,
                             int emb, float* packed, vfr_stream_t stream) {
  VFR_REQUIRE(w_ih && w_hh && b_ih && b_hh && packed, VFR_ERR_INVALID, "vfr_lstm_pack: null pointer");
  VFR_REQUIRE(hidden > 0 && emb > 0 && (hidden + emb) % 4 == 0, VFR_ERR_UNSUPPORTED,
              "vfr_lstm_pack: hidden+emb must be a multiple of 4");
  float* wcat = packed;
  float* bias = packed + (size_t)4 * hidden * (hidden + emb);
  lstm_pack_kernel<<<592, 256, 0, (cudaStream_t)stream>>>(w_ih, w_hh, b_ih, b_hh, hidden, emb, wcat, bias);
  return check_launch("lstm_pack_kernel");
}

extern "C" size_t vfr_text_embed_bytes(int64_t n_queries, int seq_len, int hidden, int emb) {
  if (n_queries <= 0 || seq_len <= 0 || hidden <= 0 || emb <= 0) return 0;
  const size_t hx = (size_t)(seq_len + 1) * n_queries * (hidden + emb);
  const size_t c = (size_t)n_queries * hidden;
  return (2 * hx + 2 * c) * sizeof(float) + 16;
}

extern "C" int vfr_text_embed(const int64_t* tokens, int64_t n_queries, int seq_len, const float* table, int64_t vocab,
                              const float* length_table, int emb, const float* packed_fwd, const float* packed_bwd,
                              int hidden, const float* fc_w, const float* fc_b, int dim, void* workspace, float* out,
                              vfr_stream_t stream) {
  VFR_REQUIRE(tokens && table && packed_fwd && packed_bwd && fc_w && fc_b && workspace && out, VFR_ERR_INVALID,
              "vfr_text_embed: null pointer");
  VFR_REQUIRE(n_queries > 0 && n_queries < (int64_t(1) << 31) && seq_len > 0 && hidden > 0 && emb > 0 && dim > 0 && vocab > 0,
              VFR_ERR_INVALID, "vfr_text_embed: bad shape");
  VFR_REQUIRE((hidden + emb) % 4 == 0 && hidden % 4 == 0, VFR_ERR_UNSUPPORTED,
              "vfr_text_embed: hidden and hidden+emb must be multiples of 4");
  cudaStream_t st = (cudaStream_t)stream;
  const int B = (int)n_queries, L = seq_len, H = hidden, E = emb, K = H + E;
  int* bad = reinterpret_cast<int*>(workspace);          // first 16 bytes: out-of-range-token flag
  float* ws = reinterpret_cast<float*>(workspace) + 4;
  const size_t hx_sz = (size_t)(L + 1) * B * K;
  float* hx[2] = {ws, ws + hx_sz};
  float* c[2] = {ws + 2 * hx_sz, ws + 2 * hx_sz + (size_t)B * H};
  VFR_CUDA(cudaMemsetAsync(bad, 0, sizeof(int), st));
  {
    const int64_t warps = (int64_t)B * L;
    const int64_t blocks = (warps + 7) / 8;
    lstm_gather_kernel<<<(unsigned)blocks, 256, 0, st>>>(tokens, B, L, table, vocab, length_table, E, H, hx[0], hx[1], bad);
    int rc = check_launch("lstm_gather_kernel");
    if (rc) return rc;
  }
  const float* wcat[2] = {packed_fwd, packed_bwd};
  const float* bias[2] = {packed_fwd + (size_t)4 * H * K, packed_bwd + (size_t)4 * H * K};
  for (int t = 0; t < L; ++t) {
    GemmBatch ops{};
    EpiLstmCell epi{};
    for (int z = 0; z < 2; ++z) {
      ops.op[z] = GemmOperand{hx[z] + (size_t)t * B * K, wcat[z]};
      epi.bias[z] = bias[z];
      epi.c[z] = c[z];
      epi.h_next[z] = hx[z] + (size_t)(t + 1) * B * K;
    }
    epi.ldh = K;
    epi.H = H;
    epi.first = (t == 0);
    int rc = launch_sgemm_nt(ops, 2, K, K, B, 4 * H, K, epi, st);
    if (rc) return rc;
  }
  // lang_fc on [h_fwd(L-1) | h_bwd(0)] = first H columns of slot L of each direction
  for (int z = 0; z < 2; ++z) {
    GemmBatch ops{};
    ops.op[0] = GemmOperand{hx[z] + (size_t)L * B * K, fc_w + (size_t)z * H};
    EpiFc epi{out, dim, dim, fc_b, z};
    int rc = launch_sgemm_nt(ops, 1, K, 2 * H, B, dim, H, epi, st);
    if (rc) return rc;
  }
  // out-of-range token ids (the reference raises IndexError) are clamped to the pad row and flagged
  // in the FIRST int32 of the workspace; the call stays asynchronous, the caller inspects the flag.
  return VFR_OK;
}
