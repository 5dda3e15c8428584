// K3 (tensor-core path): query embedding  GloVe gather -> BiLSTM -> Linear  with every GEMM on
// tcgen05 as a split-bf16 product (vfr_gemm_tc.cuh), fp32 accumulation and fp32 cell state.
// Same contract as vfr_text_embed (reference model/models.py:33-48,61-66).
//
// Layout: per direction L slots [B][2*Kp] bf16, Kp = Hp + Ep (H, E rounded up to 64); slot t holds
// the split operand [h_{t-1} | x_t] = (hi | lo), so one recurrent step is ONE GEMM against the packed
// weights [W_hh | W_ih] (rows interleaved 4j+g) whose epilogue adds the bias, applies the LSTM cell and
// writes h_t, already split into bf16 hi/lo, straight into slot t+1 (the last step writes into the
// [h_fwd | h_bwd] operand of the final projection).  Both directions share a launch (grid z).
#include "vfr_gemm_tc.cuh"

namespace vfr {

struct TextTcDims {
  int H, E, D, L;
  int Hp, Ep, Kp;        // padded widths
  int Np;                // 4H rounded up to 256 rows
  int Dp;                // D rounded up to 256 rows
  int Kf;                // 2*Hp : K of the final projection
};

// The 20 step GEMMs of a batch overlap through programmatic dependent launch (see vfr_text_embed_tc) for batches up to
// 20 000 rows - measured on B200 (tools/k3_overlap_check.py, K3 alone): 4 736 queries 3.7 -> 3.1 ms, 9 472: 6.8 -> 6.3,
// 18 944: 13.05 -> 12.9, 37 888: 26.1 -> 26.7 (64 rounds per launch leave little tail to hide, and every tile pays two
// release atomics).  VFR_K3_OVERLAP = 0 / 1 forces it off / on.
static inline bool text_overlap(int64_t rows) {
  const char* e = getenv("VFR_K3_OVERLAP");
  if (e) return atoi(e) != 0;
  return rows <= 20000;
}

static TextTcDims text_dims(int hidden, int emb, int dim, int seq_len) {
  TextTcDims d;
  d.H = hidden; d.E = emb; d.D = dim; d.L = seq_len;
  d.Hp = gt_kp(hidden); d.Ep = gt_kp(emb); d.Kp = d.Hp + d.Ep;
  d.Np = (4 * hidden + GT_BN - 1) / GT_BN * GT_BN;
  d.Dp = (dim + GT_BN - 1) / GT_BN * GT_BN;
  d.Kf = 2 * d.Hp;
  return d;
}

// packed model blob: [W fwd | W bwd | fc | bias fwd | bias bwd | fc bias]
struct TextTcBlob {
  const __nv_bfloat16* w[2];
  const __nv_bfloat16* fc;
  const float* bias[2];
  const float* fc_b;
  size_t bytes;
};
static TextTcBlob text_blob(const void* base, const TextTcDims& d) {
  TextTcBlob b;
  const __nv_bfloat16* p = reinterpret_cast<const __nv_bfloat16*>(base);
  const size_t wsz = (size_t)d.Np * 2 * d.Kp;
  b.w[0] = p;
  b.w[1] = p + wsz;
  b.fc = p + 2 * wsz;
  const float* f = reinterpret_cast<const float*>(b.fc + (size_t)d.Dp * 2 * d.Kf);
  b.bias[0] = f;
  b.bias[1] = f + 4 * d.H;
  b.fc_b = f + 8 * d.H;
  b.bytes = (2 * wsz + (size_t)d.Dp * 2 * d.Kf) * 2 + ((size_t)8 * d.H + d.D) * 4;
  return b;
}

__global__ void tc_text_pack_lstm_kernel(const float* __restrict__ w_ih, const float* __restrict__ w_hh,
                                         const float* __restrict__ b_ih, const float* __restrict__ b_hh, TextTcDims d,
                                         __nv_bfloat16* __restrict__ w, float* __restrict__ bias) {
  const int64_t total = (int64_t)4 * d.H * d.Kp;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int rp = (int)(i / d.Kp), k = (int)(i % d.Kp);
    const int j = rp >> 2, g = rp & 3, r = g * d.H + j;
    float x = 0.f;
    if (k < d.H) x = w_hh[(int64_t)r * d.H + k];
    else if (k >= d.Hp && k < d.Hp + d.E) x = w_ih[(int64_t)r * d.E + (k - d.Hp)];
    __nv_bfloat16 hi, lo;
    split2(x, hi, lo);
    w[(int64_t)rp * 2 * d.Kp + k] = hi;
    w[(int64_t)rp * 2 * d.Kp + d.Kp + k] = lo;
    if (k == 0) bias[rp] = b_ih[r] + b_hh[r];
  }
}

__global__ void tc_text_pack_fc_kernel(const float* __restrict__ fc_w, const float* __restrict__ fc_b, TextTcDims d,
                                       __nv_bfloat16* __restrict__ w, float* __restrict__ bias) {
  const int64_t total = (int64_t)d.D * d.Kf;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int r = (int)(i / d.Kf), k = (int)(i % d.Kf);
    const int dir = k / d.Hp, kk = k % d.Hp;
    float x = 0.f;
    if (kk < d.H) x = fc_w[(int64_t)r * 2 * d.H + dir * d.H + kk];
    __nv_bfloat16 hi, lo;
    split2(x, hi, lo);
    w[(int64_t)r * 2 * d.Kf + k] = hi;
    w[(int64_t)r * 2 * d.Kf + d.Kf + k] = lo;
    if (k == 0) bias[r] = fc_b[r];
  }
}

// ---- padding-aware row order ------------------------------------------------------------------------
// The reference feeds the zero padding through the LSTM (models.py:65, no packing).  In the BACKWARD direction
// a query of length len starts with L - len padding steps from the zero state: that prefix is the same for
// every query, so it is computed ONCE (row 0 of the batch is an all-padding query) and a query joins the
// recurrence at its first real token with the state row 0 has reached.  Rows are ordered by descending length,
// so the live rows of backward step t are a prefix [0, n_active[t]) and the GEMM tiles past it retire at once:
// the backward direction costs sum(len) instead of B * L cell updates (~ -30 % of K3 at DiDeMo's lengths).
struct TextOrder {
  int* len;        // [B]     tokens up to the last non-zero id
  int* perm;       // [B + 1] row -> query (row 0 = the padding row, -1)
  int* hist;       // [L + 1] queries per length
  int* cursor;     // [L + 1] next free row of each length bucket
  int* n_active;   // [L]     live backward rows of step t (padding row included)
  int* limits;     // [L][2]  per-step row limits of the two GEMM problems {forward, backward}
};

__global__ void tc_text_len_kernel(const int64_t* __restrict__ tokens, int64_t B, int L, TextOrder o) {
  const int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  int len = 0;
  for (int t = 0; t < L; ++t)
    if (tokens[b * L + t] != 0) len = t + 1;
  o.len[b] = len;
  atomicAdd(&o.hist[len], 1);
}

__global__ void tc_text_scan_kernel(int64_t B, int L, TextOrder o) {
  if (threadIdx.x != 0) return;
  int rows = 1;                                    // row 0 = the padding row
  for (int len = L; len >= 0; --len) {             // descending length
    o.cursor[len] = rows;
    rows += o.hist[len];
    if (len >= 1) {                                // step t = L - len is the first one that needs length-len rows
      const int t = L - len;
      o.n_active[t] = rows;
      o.limits[2 * t] = (int)B + 1;
      o.limits[2 * t + 1] = rows;
    }
  }
  o.perm[0] = -1;
}

__global__ void tc_text_perm_kernel(int64_t B, TextOrder o) {
  const int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  o.perm[atomicAdd(&o.cursor[o.len[b]], 1)] = (int)b;
}

// x part of every slot of both directions; one warp per (row, t); row r holds query perm[r]
__global__ void tc_text_gather_kernel(const int64_t* __restrict__ tokens, int64_t B, TextTcDims d,
                                      const float* __restrict__ table, int64_t vocab, const float* __restrict__ length,
                                      const int* __restrict__ perm, __nv_bfloat16* __restrict__ slots_f,
                                      __nv_bfloat16* __restrict__ slots_b, int* __restrict__ bad_token) {
  const int lane = threadIdx.x & 31;
  const int64_t w = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int64_t Bp = B + 1;
  if (w >= Bp * d.L) return;
  const int64_t r = w / d.L;
  const int t = (int)(w % d.L);
  const int qi = perm[r];
  int64_t id = (qi < 0) ? 0 : tokens[(int64_t)qi * d.L + t];
  if (id < 0 || id >= vocab) { if (lane == 0) atomicExch(bad_token, 1); id = 0; }
  const float* row = table + id * d.E;
  float denom = 1.f, len = 1.f;
  if (length) {
    float ss = 0.f;
    for (int k = lane; k < d.E; k += 32) ss = __fmaf_rn(row[k], row[k], ss);
    ss = warp_sum(ss);
    denom = __fadd_rn(__fsqrt_rn(ss), VFR_NORM_EPS);
    len = length[id];
  }
  const int64_t ld = 2 * (int64_t)d.Kp;
  __nv_bfloat16* df = slots_f + ((int64_t)t * Bp + r) * ld + d.Hp;              // forward consumes x_t at step t
  __nv_bfloat16* db = slots_b + ((int64_t)(d.L - 1 - t) * Bp + r) * ld + d.Hp;  // backward consumes x_{L-1-t}
  for (int k = lane; k < d.E; k += 32) {
    float v = row[k];
    if (length) v = __fmul_rn(__fdiv_rn(v, denom), len);
    __nv_bfloat16 hi, lo;
    split2(v, hi, lo);
    df[k] = hi; df[d.Kp + k] = lo;
    db[k] = hi; db[d.Kp + k] = lo;
  }
}

// What the recurrence needs zeroed in the operand slots (instead of a memset of all of them, 3.5 GB at 18 944
// queries): the whole h part of slot 0 (h_{-1} = 0) and, in every slot, the padding columns between the real
// widths and the 32-aligned ones (they meet zero weights, but 0 x NaN from uninitialised memory is NaN).
// One warp per (direction, slot, row).
__global__ void tc_text_zero_kernel(__nv_bfloat16* __restrict__ slots_f, __nv_bfloat16* __restrict__ slots_b, int64_t rows,
                                    TextTcDims d) {
  const int lane = threadIdx.x & 31;
  const int64_t w = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (w >= 2 * d.L * rows) return;
  const int z = (int)(w / (d.L * rows));
  const int64_t rem = w % (d.L * rows);
  const int t = (int)(rem / rows);
  __nv_bfloat16* row = (z ? slots_b : slots_f) + rem * 2 * (int64_t)d.Kp;
  const __nv_bfloat16 zero = __float2bfloat16_rn(0.f);
  const int h0 = (t == 0) ? 0 : d.H;
  for (int seg = 0; seg < 2; ++seg) {
    __nv_bfloat16* r = row + seg * d.Kp;
    for (int k = h0 + lane; k < d.Hp; k += 32) r[k] = zero;
    for (int k = d.Hp + d.E + lane; k < d.Kp; k += 32) r[k] = zero;
  }
}

// rows [lo, hi) take over the state of row 0 (the padding row): h (split bf16, hi and lo segments) into dst,
// c into the cell array.  lo / hi come from device memory (n_active); one warp per row.
__global__ void tc_text_join_kernel(const int* __restrict__ lo_p, const int* __restrict__ hi_p, int hi_default,
                                    __nv_bfloat16* __restrict__ dst, int64_t ld, int col0, int lo_off, int Hp,
                                    float* __restrict__ c, int H) {
  const int lane = threadIdx.x & 31;
  const int lo = lo_p ? *lo_p : 0;
  const int hi = hi_p ? *hi_p : hi_default;
  const int r = lo + blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (r >= hi || r == 0) return;
  // (all loads of a row in flight before the first store: the copy sits between two dependent GEMM launches, so it is
  //  its latency that counts - 14 us -> 5 us per step)
  const uint4* s_hi = reinterpret_cast<const uint4*>(dst + col0);
  const uint4* s_lo = reinterpret_cast<const uint4*>(dst + col0 + lo_off);
  uint4* d_hi = reinterpret_cast<uint4*>(dst + (int64_t)r * ld + col0);
  uint4* d_lo = reinterpret_cast<uint4*>(dst + (int64_t)r * ld + col0 + lo_off);
  const float4* sc = reinterpret_cast<const float4*>(c);
  float4* dc = c ? reinterpret_cast<float4*>(c + (int64_t)r * H) : nullptr;
  const int nh = Hp / 8, nc = c ? H / 4 : 0;
  for (int k0 = 0; k0 < max(nh, (nc + 1) / 2); k0 += 128) {
    uint4 a[4], b[4];
    float4 x[8];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int k = k0 + u * 32 + lane;
      if (k < nh) { a[u] = s_hi[k]; b[u] = s_lo[k]; }
    }
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      const int k = 2 * k0 + u * 32 + lane;
      if (k < nc) x[u] = sc[k];
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int k = k0 + u * 32 + lane;
      if (k < nh) { d_hi[k] = a[u]; d_lo[k] = b[u]; }
    }
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      const int k = 2 * k0 + u * 32 + lane;
      if (k < nc) dc[k] = x[u];
    }
  }
}

__device__ __forceinline__ uint32_t pack_bf16x2(__nv_bfloat16 a, __nv_bfloat16 b) {
  return (uint32_t)__bfloat16_as_ushort(a) | ((uint32_t)__bfloat16_as_ushort(b) << 16);
}
__device__ __forceinline__ float sigmoid_fast(float x) { return __fdividef(1.f, 1.f + __expf(-x)); }
__device__ __forceinline__ float tanh_fast(float x) { return 2.f * sigmoid_fast(2.f * x) - 1.f; }

// (members indexed by the problem z are selected with ?: - a runtime index into a by-value kernel parameter makes the
//  compiler copy the whole functor to LOCAL memory and fetch every pointer with an LDL in front of the loads it feeds:
//  6 % of all warp samples of the step GEMM sat on that, ncu source page of round 2)
struct EpiLstmTc {
  const float* bias[2];
  float* c[2];                 // [B, H] fp32
  __nv_bfloat16* dst[2];       // where h_t goes: row stride ld, hi at column col0 + j, lo at + lo_off
  int64_t ld;
  int col0[2];
  int lo_off;
  int H;
  // Backward direction: the rows [*join_lo, *join_hi) consume their first real token at the NEXT step and start from the
  // state the padding row (row 0) has reached (see TextOrder).  The warp that has just written row 0's 4 hidden units of a
  // column group copies them into those rows (h hi / lo in the next slot, c) - what a copy kernel between two dependent
  // GEMM launches did before (14 us per step, whatever the batch).  null: nothing joins.
  const int* join_lo;
  const int* join_hi;
  struct After {};
  __device__ __forceinline__ void after(int z, int m_warp, int n0, int lane) const {
    if (z != 1 || m_warp != 0 || join_lo == nullptr) return;           // warp-uniform
    const int j0 = n0 >> 2;
    if (j0 >= H) return;
    const int lo = max(*join_lo, 1), hi = *join_hi;
    if (hi <= lo) return;
    __syncwarp();                                                      // lane 0's stores of row 0 -> L2, then read from L2
    const float4 cv = __ldcg(reinterpret_cast<const float4*>(c[1] + j0));
    const __nv_bfloat16* o0 = dst[1] + col0[1] + j0;
    const uint2 hv = __ldcg(reinterpret_cast<const uint2*>(o0)), lv = __ldcg(reinterpret_cast<const uint2*>(o0 + lo_off));
    for (int r = lo + lane; r < hi; r += 32) {
      *reinterpret_cast<float4*>(c[1] + (int64_t)r * H + j0) = cv;
      __nv_bfloat16* o = dst[1] + (int64_t)r * ld + col0[1] + j0;
      *reinterpret_cast<uint2*>(o) = hv;
      *reinterpret_cast<uint2*>(o + lo_off) = lv;
    }
  }
  // the operand of one 16-column group (4 hidden units x 4 gates) that comes from DRAM: the previous cell state.  The
  // CTA-pair kernel requests all of a tile's groups BEFORE it waits for the accumulator (the biases are L1 hits).
  struct Pre {
    float4 c_prev;
  };
  __device__ __forceinline__ Pre prefetch(int z, int m, int n0) const {
    Pre p;
    const int j0 = n0 >> 2;
    if (j0 >= H) { p.c_prev = make_float4(0.f, 0.f, 0.f, 0.f); return p; }
    // (from L2: with overlapped step launches the row may have been written by another SM after this SM last cached it)
    p.c_prev = __ldcg(reinterpret_cast<const float4*>((z ? c[1] : c[0]) + (int64_t)m * H + j0));      // (the cell arrays start zeroed)
    return p;
  }
  __device__ __forceinline__ void operator()(int z, int m, int n0, const float (&v)[16], const Pre& p) const {
    const int j0 = n0 >> 2;
    if (j0 >= H) return;
    const float4* bz = reinterpret_cast<const float4*>((z ? bias[1] : bias[0]) + n0);
    const float cpv[4] = {p.c_prev.x, p.c_prev.y, p.c_prev.z, p.c_prev.w};
    float cn[4];
    __nv_bfloat16 hh[4], hl[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const float4 bb = __ldg(bz + u);
      const float gi = v[4 * u + 0] + bb.x, gf = v[4 * u + 1] + bb.y, gg = v[4 * u + 2] + bb.z, go = v[4 * u + 3] + bb.w;
      cn[u] = sigmoid_fast(gf) * cpv[u] + sigmoid_fast(gi) * tanh_fast(gg);
      const float h = sigmoid_fast(go) * tanh_fast(cn[u]);
      split2(h, hh[u], hl[u]);
    }
    *reinterpret_cast<float4*>((z ? c[1] : c[0]) + (int64_t)m * H + j0) = make_float4(cn[0], cn[1], cn[2], cn[3]);
    __nv_bfloat16* o = (z ? dst[1] : dst[0]) + (int64_t)m * ld + (z ? col0[1] : col0[0]) + j0;
    *reinterpret_cast<uint2*>(o) = make_uint2(pack_bf16x2(hh[0], hh[1]), pack_bf16x2(hh[2], hh[3]));
    *reinterpret_cast<uint2*>(o + lo_off) = make_uint2(pack_bf16x2(hl[0], hl[1]), pack_bf16x2(hl[2], hl[3]));
  }
};

struct EpiBiasOutTc {
  float* out;
  int64_t ldo;
  int N;
  const float* bias;
  int relu;
  const int* row_map;          // optional: output row of GEMM row m (negative = no output)
  __device__ __forceinline__ void operator()(int, int m, int n0, const float (&v)[16]) const {
    if (row_map) {
      m = row_map[m];
      if (m < 0) return;
    }
#pragma unroll
    for (int j = 0; j < 16; ++j) {
      const int n = n0 + j;
      if (n < N) {
        float x = v[j] + (bias ? __ldg(bias + n) : 0.f);
        if (relu) x = fmaxf(x, 0.f);
        out[(int64_t)m * ldo + n] = x;
      }
    }
  }
};

// fp32 rows -> split bf16 packed rows (generic operand preparation, one warp per row)
__global__ void tc_split_rows_kernel(const float* __restrict__ x, int64_t rows, int k, int64_t ldx, int kp,
                                     __nv_bfloat16* __restrict__ out) {
  const int lane = threadIdx.x & 31;
  const int64_t r = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (r >= rows) return;
  const float* src = x + r * ldx;
  __nv_bfloat16* dst = out + r * 2 * (int64_t)kp;
  for (int c = lane; c < kp; c += 32) {
    __nv_bfloat16 hi, lo;
    split2(c < k ? src[c] : 0.f, hi, lo);
    dst[c] = hi;
    dst[kp + c] = lo;
  }
}

}  // namespace vfr

using namespace vfr;

// ---------------------------------------------------------------------------------------------
// generic split-bf16 linear layer (building block of the K2 tensor-core path; also the GEMM self-test)
// ---------------------------------------------------------------------------------------------
extern "C" size_t vfr_tc_weight_bytes(int out_dim, int in_dim) {
  if (out_dim <= 0 || in_dim <= 0) return 0;
  return (size_t)out_dim * 2 * gt_kp(in_dim) * 2;
}

extern "C" int vfr_tc_weight_pack(const float* w, int out_dim, int in_dim, void* packed, vfr_stream_t stream) {
  VFR_REQUIRE(w && packed && out_dim > 0 && in_dim > 0, VFR_ERR_INVALID, "vfr_tc_weight_pack: bad argument");
  tc_split_rows_kernel<<<(out_dim + 7) / 8, 256, 0, (cudaStream_t)stream>>>(w, out_dim, in_dim, in_dim, gt_kp(in_dim),
                                                                            reinterpret_cast<__nv_bfloat16*>(packed));
  return check_launch("tc_split_rows_kernel");
}

extern "C" size_t vfr_linear_tc_bytes(int64_t n_rows, int in_dim) {
  if (n_rows <= 0 || in_dim <= 0) return 0;
  return (size_t)n_rows * 2 * gt_kp(in_dim) * 2;
}

extern "C" int vfr_linear_tc(const float* x, int64_t n_rows, int in_dim, int64_t ldx, const void* w_packed,
                             const float* bias, int out_dim, int relu, float* out, int64_t ldo, void* workspace,
                             vfr_stream_t stream) {
  VFR_REQUIRE(x && w_packed && out && workspace, VFR_ERR_INVALID, "vfr_linear_tc: null pointer");
  VFR_REQUIRE(n_rows >= 0 && n_rows < (int64_t(1) << 31) && in_dim > 0 && out_dim > 0 && ldx >= in_dim && ldo >= out_dim,
              VFR_ERR_INVALID, "vfr_linear_tc: bad shape");
  if (n_rows == 0) return VFR_OK;
  cudaStream_t st = (cudaStream_t)stream;
  const int kp = gt_kp(in_dim);
  __nv_bfloat16* xs = reinterpret_cast<__nv_bfloat16*>(workspace);
  tc_split_rows_kernel<<<(unsigned)((n_rows + 7) / 8), 256, 0, st>>>(x, n_rows, in_dim, ldx, kp, xs);
  int rc = check_launch("tc_split_rows_kernel");
  if (rc) return rc;
  const void* a[1] = {xs};
  const void* b[1] = {w_packed};
  EpiBiasOutTc epi{out, ldo, out_dim, bias, relu, nullptr};
  return launch_gemm_tc(a, b, 1, (int)n_rows, out_dim, kp, 2 * (int64_t)kp, 2 * (int64_t)kp, epi, st);
}

// ---------------------------------------------------------------------------------------------
// K3
// ---------------------------------------------------------------------------------------------
extern "C" size_t vfr_text_pack_tc_bytes(int hidden, int emb, int dim) {
  if (hidden <= 0 || emb <= 0 || dim <= 0) return 0;
  const TextTcDims d = text_dims(hidden, emb, dim, 1);
  return text_blob(nullptr, d).bytes;
}

extern "C" int vfr_text_pack_tc(const float* w_ih_f, const float* w_hh_f, const float* b_ih_f, const float* b_hh_f,
                                const float* w_ih_b, const float* w_hh_b, const float* b_ih_b, const float* b_hh_b,
                                const float* fc_w, const float* fc_b, int hidden, int emb, int dim, void* packed,
                                vfr_stream_t stream) {
  VFR_REQUIRE(w_ih_f && w_hh_f && b_ih_f && b_hh_f && w_ih_b && w_hh_b && b_ih_b && b_hh_b && fc_w && fc_b && packed,
              VFR_ERR_INVALID, "vfr_text_pack_tc: null pointer");
  VFR_REQUIRE(hidden > 0 && hidden % 4 == 0 && emb > 0 && dim > 0, VFR_ERR_UNSUPPORTED,
              "vfr_text_pack_tc: hidden must be a positive multiple of 4");
  const TextTcDims d = text_dims(hidden, emb, dim, 1);
  const TextTcBlob blob = text_blob(packed, d);
  cudaStream_t st = (cudaStream_t)stream;
  VFR_CUDA(cudaMemsetAsync(packed, 0, blob.bytes, st));
  const float* wi[2] = {w_ih_f, w_ih_b};
  const float* wh[2] = {w_hh_f, w_hh_b};
  const float* bi[2] = {b_ih_f, b_ih_b};
  const float* bh[2] = {b_hh_f, b_hh_b};
  for (int z = 0; z < 2; ++z) {
    tc_text_pack_lstm_kernel<<<1184, 256, 0, st>>>(wi[z], wh[z], bi[z], bh[z], d, const_cast<__nv_bfloat16*>(blob.w[z]),
                                                   const_cast<float*>(blob.bias[z]));
    int rc = check_launch("tc_text_pack_lstm_kernel");
    if (rc) return rc;
  }
  tc_text_pack_fc_kernel<<<296, 256, 0, st>>>(fc_w, fc_b, d, const_cast<__nv_bfloat16*>(blob.fc),
                                              const_cast<float*>(blob.fc_b));
  return check_launch("tc_text_pack_fc_kernel");
}

extern "C" size_t vfr_text_embed_tc_bytes(int64_t n_queries, int seq_len, int hidden, int emb) {
  if (n_queries <= 0 || seq_len <= 0 || hidden <= 0 || emb <= 0) return 0;
  const TextTcDims d = text_dims(hidden, emb, 1, seq_len);
  const size_t rows = (size_t)n_queries + 1;                                // + the padding row
  const size_t slots = (size_t)2 * seq_len * rows * 2 * d.Kp * 2;          // bf16
  const size_t hcat = rows * 2 * d.Kf * 2;
  const size_t c = (size_t)2 * rows * hidden * 4;
  const size_t order = ((size_t)2 * rows + 6 * ((size_t)seq_len + 1) + 16) * 4;
  const size_t deps = (size_t)seq_len * (2 * ((rows + GT_BM - 1) / GT_BM) + 1) * 4;      // row-block counters of the step GEMMs
  return 16 + slots + hcat + c + order + deps;
}

extern "C" int vfr_text_embed_tc(const int64_t* tokens, int64_t n_queries, int seq_len, const float* table,
                                 int64_t vocab, const float* length_table, int emb, const void* packed, int hidden,
                                 int dim, void* workspace, float* out, vfr_stream_t stream) {
  VFR_REQUIRE(tokens && table && packed && workspace && out, VFR_ERR_INVALID, "vfr_text_embed_tc: null pointer");
  VFR_REQUIRE(n_queries > 0 && n_queries < (int64_t(1) << 31) - 1 && seq_len > 0 && hidden > 0 && hidden % 4 == 0 &&
                  emb > 0 && dim > 0 && vocab > 0,
              VFR_ERR_INVALID, "vfr_text_embed_tc: bad shape");
  cudaStream_t st = (cudaStream_t)stream;
  const TextTcDims d = text_dims(hidden, emb, dim, seq_len);
  const TextTcBlob blob = text_blob(packed, d);
  const int64_t B = n_queries, Bp = n_queries + 1;
  const int L = seq_len;
  const int64_t ld = 2 * (int64_t)d.Kp;
  uint8_t* base = reinterpret_cast<uint8_t*>(workspace);
  int* bad = reinterpret_cast<int*>(base);
  __nv_bfloat16* slots[2];
  slots[0] = reinterpret_cast<__nv_bfloat16*>(base + 16);
  slots[1] = slots[0] + (size_t)L * Bp * ld;
  __nv_bfloat16* hcat = slots[1] + (size_t)L * Bp * ld;
  float* c0 = reinterpret_cast<float*>(hcat + (size_t)Bp * 2 * d.Kf);
  float* c[2] = {c0, c0 + (size_t)Bp * hidden};
  int* ints = reinterpret_cast<int*>(c0 + (size_t)2 * Bp * hidden);
  TextOrder o;
  o.len = ints;
  o.perm = o.len + Bp;
  o.hist = o.perm + Bp;
  o.cursor = o.hist + (L + 1);
  o.n_active = o.cursor + (L + 1);
  o.limits = o.n_active + (L + 1);
  const int n_mb = (int)((Bp + GT_BM - 1) / GT_BM);
  int* dep = ints + ((size_t)2 * Bp + 6 * ((size_t)L + 1) + 16);       // [L][2 n_mb + 1], zeroed with the rest below
  // zero: flag; h_{-1} and the pad columns of the operand slots; the final operand, the cell states, the counters
  VFR_CUDA(cudaMemsetAsync(base, 0, 16, st));
  VFR_CUDA(cudaMemsetAsync(hcat, 0, base + vfr_text_embed_tc_bytes(n_queries, seq_len, hidden, emb) - reinterpret_cast<uint8_t*>(hcat), st));
  {
    const int64_t warps = 2 * (int64_t)L * Bp;
    tc_text_zero_kernel<<<(unsigned)((warps + 7) / 8), 256, 0, st>>>(slots[0], slots[1], Bp, d);
    int rc0 = check_launch("tc_text_zero_kernel");
    if (rc0) return rc0;
  }
  tc_text_len_kernel<<<(unsigned)((B + 255) / 256), 256, 0, st>>>(tokens, B, L, o);
  int rc = check_launch("tc_text_len_kernel");
  if (rc) return rc;
  tc_text_scan_kernel<<<1, 32, 0, st>>>(B, L, o);
  rc = check_launch("tc_text_scan_kernel");
  if (rc) return rc;
  tc_text_perm_kernel<<<(unsigned)((B + 255) / 256), 256, 0, st>>>(B, o);
  rc = check_launch("tc_text_perm_kernel");
  if (rc) return rc;
  {
    const int64_t warps = Bp * L;
    tc_text_gather_kernel<<<(unsigned)((warps + 7) / 8), 256, 0, st>>>(tokens, B, d, table, vocab, length_table, o.perm,
                                                                     slots[0], slots[1], bad);
    rc = check_launch("tc_text_gather_kernel");
    if (rc) return rc;
  }
  const unsigned join_blocks = (unsigned)((Bp + 7) / 8);
  // the CTA-pair GEMM copies the padding row's state into the joining rows from its epilogue (EpiLstmTc::after); the
  // one-CTA kernel (VFR_GEMM2=0) keeps the copy kernel between the steps
  const bool fused_join = g2_enabled() && d.Kp % 64 == 0;
  for (int t = 0; t < L; ++t) {
    if (t > 0 && !fused_join) {
      // rows whose first real token is consumed by backward step t join with the padding row's state
      tc_text_join_kernel<<<join_blocks, 256, 0, st>>>(o.n_active + (t - 1), o.n_active + t, 0,
                                                       slots[1] + (size_t)t * Bp * ld, ld, 0, d.Kp, d.Hp, c[1], hidden);
      rc = check_launch("tc_text_join_kernel");
      if (rc) return rc;
    }
    const void* a[2] = {slots[0] + (size_t)t * Bp * ld, slots[1] + (size_t)t * Bp * ld};
    const void* b[2] = {blob.w[0], blob.w[1]};
    EpiLstmTc epi{};
    const bool last = (t == L - 1);
    for (int z = 0; z < 2; ++z) {
      epi.bias[z] = blob.bias[z];
      epi.c[z] = c[z];
      epi.dst[z] = last ? hcat : slots[z] + (size_t)(t + 1) * Bp * ld;
      epi.col0[z] = last ? z * d.Hp : 0;
    }
    epi.ld = last ? 2 * (int64_t)d.Kf : ld;
    epi.lo_off = last ? d.Kf : d.Kp;
    epi.H = hidden;
    epi.join_lo = (fused_join && !last) ? o.n_active + t : nullptr;
    epi.join_hi = (fused_join && !last) ? o.n_active + t + 1 : nullptr;
    if (fused_join && text_overlap(Bp)) {
      // consecutive step GEMMs overlap (programmatic dependent launch): a tile of step t waits for the row block of step
      // t - 1 it reads (and, backward direction, for row block 0, the source of the joining rows) instead of for the
      // whole previous grid - no exposed epilogue tail, prologue and round quantisation between the 20 launches
      G2Deps dp{};
      dp.prev = t > 0 ? dep + (size_t)(t - 1) * (2 * n_mb + 1) : nullptr;
      dp.cur = dep + (size_t)t * (2 * n_mb + 1);
      dp.prev_limit = t > 0 ? o.limits + 2 * (t - 1) : nullptr;
      dp.n_blocks = n_mb;
      dp.src_block[0] = -1;
      dp.src_block[1] = 0;
      rc = launch_gemm_tc2(a, b, 2, (int)Bp, 4 * hidden, d.Kp, ld, ld, epi, st, o.limits + 2 * t, false, d.Kp, d.Kp, 0, nullptr, 0,
                           &dp, t > 0);
    } else {
      rc = launch_gemm_tc(a, b, 2, (int)Bp, 4 * hidden, d.Kp, ld, ld, epi, st, o.limits + 2 * t);
    }
    if (rc) return rc;
  }
  // all-padding queries never joined: their backward state is the padding row's final one
  tc_text_join_kernel<<<join_blocks, 256, 0, st>>>(o.n_active + (L - 1), nullptr, (int)Bp, hcat, 2 * (int64_t)d.Kf, d.Hp, d.Kf,
                                                   d.Hp, nullptr, hidden);
  rc = check_launch("tc_text_join_kernel");
  if (rc) return rc;
  const void* a[1] = {hcat};
  const void* b[1] = {blob.fc};
  EpiBiasOutTc epi{out, dim, dim, blob.fc_b, 0, o.perm};
  return launch_gemm_tc(a, b, 1, (int)Bp, dim, d.Kf, 2 * (int64_t)d.Kf, 2 * (int64_t)d.Kf, epi, st);
}
