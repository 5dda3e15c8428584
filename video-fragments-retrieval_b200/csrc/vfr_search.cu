// Retrieval step = K3 (query embedding) -> K4 (fused score + top-k) behind one C call, with a
// HOST-buffer flavour (pinned or pageable host pointers in, host pointers out; the H2D / D2H copies
// are issued on the caller's stream inside the call).  This is the serving form of the hot loop of
// reference model/evaluate.py:42-80: one batch of tokenised queries against the resident bank.
#include "vfr_common.cuh"
#include <algorithm>

using namespace vfr;

static int check_plan(const vfr_search_plan* p, int64_t n_queries, int k) {
  VFR_REQUIRE(p, VFR_ERR_INVALID, "vfr_search: null plan");
  VFR_REQUIRE(p->table && p->vid_off && p->mom_off, VFR_ERR_INVALID, "vfr_search: null model/bank pointer in plan");
  VFR_REQUIRE(p->text_engine == 0 || p->text_engine == 3, VFR_ERR_INVALID, "vfr_search: text_engine=%d", p->text_engine);
  if (p->text_engine == 0)
    VFR_REQUIRE(p->lstm_fwd && p->lstm_bwd && p->fc_w && p->fc_b, VFR_ERR_INVALID, "vfr_search: null text weights");
  else
    VFR_REQUIRE(p->text_tc, VFR_ERR_INVALID, "vfr_search: text engine 3 needs text_tc");
  VFR_REQUIRE(p->tokens_dev && p->q_emb && p->text_ws && p->topk_ws && p->out_scores_dev && p->out_ids_dev,
              VFR_ERR_INVALID, "vfr_search: null scratch pointer in plan");
  VFR_REQUIRE(p->engine == 0 || p->engine == 1 || p->engine == 3 || p->engine == 4, VFR_ERR_INVALID, "vfr_search: engine=%d",
              p->engine);
  if (p->engine == 4) VFR_REQUIRE(p->n_clips > 0, VFR_ERR_INVALID, "vfr_search: engine 4 needs n_clips");
  if (p->engine == 0)
    VFR_REQUIRE(p->bank_packed && p->q_packed, VFR_ERR_INVALID, "vfr_search: engine 0 needs bank_packed and q_packed");
  else
    VFR_REQUIRE(p->bank_tc && p->bank_clips && p->q_tc, VFR_ERR_INVALID, "vfr_search: tensor-core engine needs bank_tc, bank_clips, q_tc");
  VFR_REQUIRE(n_queries > 0 && n_queries <= p->max_queries, VFR_ERR_INVALID,
              "vfr_search: n_queries=%lld exceeds plan capacity %lld", (long long)n_queries, (long long)p->max_queries);
  VFR_REQUIRE(k >= 1 && k <= VFR_TOPK_MAX, VFR_ERR_UNSUPPORTED, "vfr_search: k=%d", k);
  return VFR_OK;
}

// stage 1: K3 only - tokens -> query embeddings (q_emb_out fp32 [n_queries, dim], may be a slice of plan->q_emb)
extern "C" int vfr_search_embed_device(const vfr_search_plan* p, const int64_t* tokens_dev, int64_t n_queries,
                                       float* q_emb_out, vfr_stream_t stream) {
  int rc = check_plan(p, n_queries, 1);
  if (rc) return rc;
  VFR_REQUIRE(tokens_dev && q_emb_out, VFR_ERR_INVALID, "vfr_search_embed_device: null pointer");
  if (p->text_engine == 3)
    return vfr_text_embed_tc(tokens_dev, n_queries, p->seq_len, p->table, p->vocab, p->length_table, p->emb, p->text_tc,
                             p->hidden, p->dim, p->text_ws, q_emb_out, stream);
  return vfr_text_embed(tokens_dev, n_queries, p->seq_len, p->table, p->vocab, p->length_table, p->emb, p->lstm_fwd,
                        p->lstm_bwd, p->hidden, p->fc_w, p->fc_b, p->dim, p->text_ws, q_emb_out, stream);
}

// stage 2: K4 only - the first n_queries rows of plan->q_emb against the resident bank
extern "C" int vfr_search_score_device(const vfr_search_plan* p, int64_t n_queries, int k, float* out_scores_dev,
                                       int64_t* out_ids_dev, vfr_stream_t stream) {
  int rc = check_plan(p, n_queries, k);
  if (rc) return rc;
  VFR_REQUIRE(out_scores_dev && out_ids_dev, VFR_ERR_INVALID, "vfr_search_score_device: null pointer");
  if (p->engine == 4) {
    rc = vfr_sel_query_pack(p->q_emb, n_queries, p->dim, p->bank_tc, p->n_clips, p->q_tc, stream);
    if (rc) return rc;
    return vfr_sel_topk(p->bank_tc, p->bank_clips, p->vid_off, p->mom_off, p->n_videos, p->n_clips, p->n_max, p->dim,
                        p->q_tc, p->q_emb, n_queries, k, p->id_base, out_scores_dev, out_ids_dev, p->topk_ws, p->n_split,
                        stream);
  }
  if (p->engine != 0) {
    rc = vfr_tc_query_pack(p->q_emb, n_queries, p->dim, p->engine, p->q_tc, stream);
    if (rc) return rc;
    return vfr_score_topk_tc(p->bank_tc, p->bank_clips, p->vid_off, p->mom_off, p->n_videos, p->uniform6, p->dim,
                             p->engine, p->q_tc, p->q_emb, n_queries, k, p->id_base, out_scores_dev, out_ids_dev,
                             p->topk_ws, p->n_split, stream);
  }
  rc = vfr_query_pack(p->q_emb, n_queries, p->dim, p->q_packed, stream);
  if (rc) return rc;
  return vfr_score_topk(p->bank_packed, p->vid_off, p->mom_off, p->n_videos, p->n_max, p->dim, p->q_packed, n_queries,
                        k, p->id_base, out_scores_dev, out_ids_dev, p->topk_ws, p->n_split, stream);
}

extern "C" int vfr_search_device(const vfr_search_plan* p, const int64_t* tokens_dev, int64_t n_queries, int k,
                                 float* out_scores_dev, int64_t* out_ids_dev, vfr_stream_t stream) {
  int rc = check_plan(p, n_queries, k);
  if (rc) return rc;
  rc = vfr_search_embed_device(p, tokens_dev, n_queries, p->q_emb, stream);
  if (rc) return rc;
  return vfr_search_score_device(p, n_queries, k, out_scores_dev, out_ids_dev, stream);
}

// copy stream + events of the host-buffer step (created on first use, one set per process)
namespace {
struct HostCopy {
  cudaStream_t stream = nullptr;
  cudaEvent_t ready[8] = {};
  cudaEvent_t done = nullptr;
  bool ok = false;
};
// one set per device (a process may serve several GPUs); not for concurrent calls on the same device from several threads
HostCopy& host_copy() {
  static HostCopy per_device[16];
  static HostCopy none;
  int dev = -1;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 16) return none;
  HostCopy& h = per_device[dev];
  if (!h.stream) {
    bool ok = cudaStreamCreateWithFlags(&h.stream, cudaStreamNonBlocking) == cudaSuccess;
    for (int i = 0; i < 8 && ok; ++i) ok = cudaEventCreateWithFlags(&h.ready[i], cudaEventDisableTiming) == cudaSuccess;
    ok = ok && cudaEventCreateWithFlags(&h.done, cudaEventDisableTiming) == cudaSuccess;
    h.ok = ok;
    if (!ok) cudaGetLastError();       // fall back to the plain path, do not leave a sticky-looking error behind
  }
  return h;
}
}  // namespace

extern "C" int vfr_search_host(const vfr_search_plan* p, const int64_t* tokens_host, int64_t n_queries, int k,
                               float* out_scores_host, int64_t* out_ids_host, vfr_stream_t stream) {
  int rc = check_plan(p, n_queries, k);
  if (rc) return rc;
  VFR_REQUIRE(tokens_host && out_scores_host && out_ids_host, VFR_ERR_INVALID, "vfr_search_host: null pointer");
  cudaStream_t st = (cudaStream_t)stream;
  VFR_CUDA(cudaMemcpyAsync(p->tokens_dev, tokens_host, (size_t)n_queries * p->seq_len * sizeof(int64_t),
                           cudaMemcpyHostToDevice, st));
  // Engine 4, large batches: stage 2 (exact re-scoring, ~3 ms per 37 888 queries) runs in query chunks, and a finished
  // chunk's rows travel to the host on a second stream while the next chunk is re-scored - of the 45 MB of results only
  // the last chunk's copy (instead of all 0.8 ms) is exposed at the end of the step.
  HostCopy& hc = host_copy();
  const int chunks = (p->engine == 4 && hc.ok && n_queries >= 4096) ? 8 : 1;
  if (chunks > 1) {
    rc = vfr_search_embed_device(p, p->tokens_dev, n_queries, p->q_emb, stream);
    if (rc) return rc;
    rc = vfr_sel_query_pack(p->q_emb, n_queries, p->dim, p->bank_tc, p->n_clips, p->q_tc, stream);
    if (rc) return rc;
    rc = vfr_sel_filter(p->bank_tc, p->n_clips, p->dim, p->q_tc, n_queries, k, p->topk_ws, p->n_split, 0, -1, 0, stream);
    if (rc) return rc;
    const int64_t per = (n_queries + chunks - 1) / chunks;
    for (int c = 0; c < chunks; ++c) {
      const int64_t q0 = c * per, n = std::min<int64_t>(per, n_queries - q0);
      if (n <= 0) break;
      rc = vfr_sel_refine_range(p->bank_clips, p->vid_off, p->mom_off, p->n_videos, p->n_clips, p->n_max, p->dim, p->q_tc,
                                p->q_emb, n_queries, k, p->id_base, p->out_scores_dev, p->out_ids_dev, p->topk_ws, p->n_split,
                                q0, n, stream);
      if (rc) return rc;
      VFR_CUDA(cudaEventRecord(hc.ready[c], st));
      VFR_CUDA(cudaStreamWaitEvent(hc.stream, hc.ready[c], 0));
      VFR_CUDA(cudaMemcpyAsync(out_scores_host + q0 * k, p->out_scores_dev + q0 * k, (size_t)n * k * sizeof(float),
                               cudaMemcpyDeviceToHost, hc.stream));
      VFR_CUDA(cudaMemcpyAsync(out_ids_host + q0 * k, p->out_ids_dev + q0 * k, (size_t)n * k * sizeof(int64_t),
                               cudaMemcpyDeviceToHost, hc.stream));
    }
    VFR_CUDA(cudaEventRecord(hc.done, hc.stream));
    VFR_CUDA(cudaStreamWaitEvent(st, hc.done, 0));       // the caller's stream stays the one thing to wait for
    VFR_CUDA(cudaStreamSynchronize(st));
    return VFR_OK;
  }
  rc = vfr_search_device(p, p->tokens_dev, n_queries, k, p->out_scores_dev, p->out_ids_dev, stream);
  if (rc) return rc;
  VFR_CUDA(cudaMemcpyAsync(out_scores_host, p->out_scores_dev, (size_t)n_queries * k * sizeof(float),
                           cudaMemcpyDeviceToHost, st));
  VFR_CUDA(cudaMemcpyAsync(out_ids_host, p->out_ids_dev, (size_t)n_queries * k * sizeof(int64_t),
                           cudaMemcpyDeviceToHost, st));
  VFR_CUDA(cudaStreamSynchronize(st));
  return VFR_OK;
}
