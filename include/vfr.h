/* libvfr - C ABI of the B200-native moment-scoring hot path (sm_100a).
 *
 * This header is the drop-in boundary.  The reference (mariyashcheg/video-fragments-retrieval) is
 * pure Python with no FFI of its own, so each entry point below names the reference lines whose
 * work it replaces; the Python shim in video-fragments-retrieval_b200/ binds them with ctypes
 * (see INTEGRATION.md for the stub a reference maintainer would add).
 *
 * Conventions
 *  - plain C: pointers + sizes only.  Unless a parameter is marked HOST every pointer is a DEVICE
 *    pointer owned by the caller; the library never allocates or frees caller-visible memory and
 *    takes caller-supplied workspaces (sizes from the vfr_*_bytes queries).
 *  - every call takes the CUDA stream to enqueue on (cudaStream_t passed as void*, NULL = legacy
 *    default stream) and is asynchronous with respect to the host unless stated otherwise.
 *  - return value: 0 = ok, negative = error (VFR_ERR_*); vfr_last_error() gives the text for the
 *    calling thread.  Nothing throws or aborts.
 *  - moments of an n-clip video are indexed as the reference enumerates them
 *    (model/utils.py:71-75): the n single clips, then all (s<e) pairs in lexicographic order.
 *  - a "bank" is the set of clip embeddings of V videos: fp32 [C, D] row-major, with CSR clip
 *    offsets vid_off int32 [V+1] and moment offsets mom_off int64 [V+1]
 *    (mom_off[v+1]-mom_off[v] = n_v (n_v+1)/2, n_v <= 32).
 */
#ifndef VFR_H_
#define VFR_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define VFR_OK 0
#define VFR_ERR_INVALID (-1)     /* bad argument (null pointer, negative size, misalignment) */
#define VFR_ERR_UNSUPPORTED (-2) /* shape outside what the kernels handle */
#define VFR_ERR_CUDA (-3)        /* a CUDA runtime call / launch failed */

#define VFR_TILE_Q 128    /* queries per scoring tile */
#define VFR_TILE_C 96     /* clip columns per scoring tile */
#define VFR_TOPK_MAX 128  /* largest k of the fused top-k */
#define VFR_TOPK_CAP 512  /* candidate slots per (query, part) list in the top-k workspace */
#define VFR_MAX_TAU 16    /* thresholds per query in count mode */

typedef void* vfr_stream_t;

const char* vfr_last_error(void);
int vfr_version(void);
/* number of SMs of the current device (grid sizing), or negative error */
int vfr_device_sms(void);
/* kernels this library has launched in this process so far (what bench.py reports as gpu_launches) */
int64_t vfr_launch_count(void);

/* ---- K4 : query x clip distance -> moment means -> full / count / top-k --------------------
 * replaces model/evaluate.py:49-58 (and evaluate_single.py:48-53, main.py:148-157):
 *   d[q,c] = || v_c - q + 1e-6 ||_2 ;  score[q,(v,s,e)] = mean(d[q, v, s..e])  (fp32)
 * Exact-fp32 CUDA-core path (direct-difference form, k-sequential FFMA chain, IEEE sqrt/div);
 * every mode evaluates the SAME arithmetic, so thresholds taken from one mode compare exactly
 * against scores of another.
 */

/* Packed operand layouts (k-major tiles streamed by cp.async.bulk).
 * videos per tile VT = VFR_TILE_C / n_max ; tiles = ceil(V / VT). */
size_t vfr_bank_pack_bytes(int64_t n_videos, int n_max, int dim);
int vfr_bank_pack(const float* bank, const int32_t* vid_off, int64_t n_videos, int n_max, int dim,
                  float* packed, vfr_stream_t stream);
size_t vfr_query_pack_bytes(int64_t n_queries, int dim);
int vfr_query_pack(const float* queries, int64_t n_queries, int dim, float* packed, vfr_stream_t stream);

/* all scores, reference order (video-major, then moment index): out fp32 [Q, M_total] */
int vfr_score_full(const float* bank_packed, const int32_t* vid_off, const int64_t* mom_off,
                   int64_t n_videos, int n_max, int dim, const float* query_packed, int64_t n_queries,
                   float* out, int64_t m_total, vfr_stream_t stream);

/* scores of each query against ONE video (its own): out fp32 [Q, m_stride], +inf padded.
 * bank is the UNPACKED fp32 [C, D] array, queries the unpacked [Q, D]. (evaluate_single.py:48-53) */
int vfr_score_own(const float* bank, const int32_t* vid_off, int dim, const float* queries,
                  int64_t n_queries, const int32_t* q_video, float* out, int m_stride, vfr_stream_t stream);

/* rank counting (replaces the argsort of model/evaluate.py:71,77 - only the position of the
 * first positive is consumed): for each query q and threshold t<n_tau,
 *   cnt_lt[q,t]  += #{ moments : score <  tau[q,t] }
 *   cnt_eqb[q,t] += #{ moments : score == tau[q,t] and video < q_video[q] }   (deterministic ties)
 * counters are uint32 [Q, n_tau] and must be zeroed by the caller. */
int vfr_score_count(const float* bank_packed, const int32_t* vid_off, int64_t n_videos, int n_max, int dim,
                    const float* query_packed, int64_t n_queries, const float* tau, int n_tau,
                    const int32_t* q_video, uint32_t* cnt_lt, uint32_t* cnt_eqb, int n_split,
                    vfr_stream_t stream);

/* fused top-k (k <= VFR_TOPK_MAX) by ascending (score, moment id):
 * out_scores fp32 [Q, k], out_ids int64 [Q, k] = id_base + local moment id; unused slots
 * (+inf, -1).  workspace: vfr_score_topk_bytes(Q, n_split). */
size_t vfr_score_topk_bytes(int64_t n_queries, int n_split);
int vfr_score_topk(const float* bank_packed, const int32_t* vid_off, const int64_t* mom_off,
                   int64_t n_videos, int n_max, int dim, const float* query_packed, int64_t n_queries,
                   int k, int64_t id_base, float* out_scores, int64_t* out_ids, void* workspace,
                   int n_split, vfr_stream_t stream);

/* K7 : merge P per-shard top-k lists (after the NCCL all-gather): in_scores fp32 [P, Q, k],
 * in_ids int64 [P, Q, k] -> out [Q, k] by ascending (score, id).  P * k <= 4096. */
int vfr_topk_merge(const float* in_scores, const int64_t* in_ids, int n_parts, int64_t n_queries, int k,
                   float* out_scores, int64_t* out_ids, vfr_stream_t stream);

/* ---- K4, tensor-core path (tcgen05 / TMEM / TMA) ----------------------------------------------------
 * Same outputs as vfr_score_topk / vfr_score_full at the north-star tolerance (fp32 scores within
 * 1e-5 of the reference): the query x clip contraction runs as a split-bf16 GEMM (n_terms = 3:
 * qh.vh + ql.vh + qh.vl, fp32 accumulation in TMEM) with an exact-fp32 recompute of near-duplicate
 * pairs; n_terms = 1 is the plain-bf16 variant (1e-2 tolerance).  Videos of at most 6 clips
 * (fixed 6-slot layout), dim <= 125.  Means are sum * fl(1/len), sqrt is sqrt.approx (<= 2 ulp).
 * Packed operands: vfr_tc_bank_bytes / vfr_tc_query_bytes; `bank`, `queries`, `vid_off`, `mom_off`
 * are the unpacked arrays (read only on the exact-fallback / append paths); uniform != 0 asserts that
 * every video has exactly 6 clips. */
size_t vfr_tc_bank_bytes(int64_t n_videos);
int vfr_tc_bank_pack(const float* bank, const int32_t* vid_off, int64_t n_videos, int n_max, int dim,
                     int n_terms, void* packed, vfr_stream_t stream);
size_t vfr_tc_query_bytes(int64_t n_queries);
int vfr_tc_query_pack(const float* queries, int64_t n_queries, int dim, int n_terms, void* packed,
                      vfr_stream_t stream);
size_t vfr_score_topk_tc_bytes(int64_t n_queries, int64_t n_videos, int n_split);
int vfr_score_topk_tc(const void* bank_packed, const float* bank, const int32_t* vid_off,
                      const int64_t* mom_off, int64_t n_videos, int uniform, int dim, int n_terms,
                      const void* query_packed, const float* queries, int64_t n_queries, int k,
                      int64_t id_base, float* out_scores, int64_t* out_ids, void* workspace, int n_split,
                      vfr_stream_t stream);
int vfr_score_full_tc(const void* bank_packed, const float* bank, const int32_t* vid_off,
                      const int64_t* mom_off, int64_t n_videos, int uniform, int dim, int n_terms,
                      const void* query_packed, const float* queries, int64_t n_queries, float* out,
                      int64_t m_total, vfr_stream_t stream);

/* ---- K4, filter + refine top-k (one fp16 tcgen05 pass + exact fp32 re-scoring) ------------------------
 * Same contract and the SAME BITS as vfr_score_topk (reference model/evaluate.py:49-58,71-80): a
 * moment's score is the mean of its clips' distances, so the k best moments all live in videos that own
 * one of the ~k closest clips.  Stage 1 finds those clips with a single fp16 GEMM pass (fp32
 * accumulation in TMEM) whose rounding error is bounded rigorously per query; every clip inside the
 * bound of the k-th smallest squared distance survives.  Stage 2 re-scores all moments of the surviving
 * videos with the exact-fp32 arithmetic of vfr_score_topk and selects the k best by (score, moment id).
 * Any number of clips per video (<= 32), dim <= 1085; the bank is packed by clip rows (no video padding): rows of 128 fp16
 * while dim + 3 <= 128 (query tiles resident in shared memory), rows of ceil((dim + 3) / 64) * 64 fp16 beyond that (BASELINE
 * configs[2]: the 1024-d joint space) - there the kernel streams 64-column chunks of the query tiles AND the bank tile
 * through the TMA ring and the two TMEM accumulators integrate over the chunks (same epilogue, same bits out).
 *   packed bank    vfr_sel_bank_bytes(n_clips, dim)   <- vfr_sel_bank_pack (once per bank)
 *   packed queries vfr_sel_query_bytes(n_queries, dim) <- vfr_sel_query_pack (per batch; reads the bank's scales)
 *   workspace      vfr_sel_topk_bytes(max n_queries, n_clips, n_split)
 * vfr_sel_flags(query_packed, n_queries, dim) -> device pointer to int32 [n_queries] written by the last
 * vfr_sel_query_pack / vfr_sel_topk on that buffer: 0 = result guaranteed exact; 1 = the query's or the
 * bank's magnitudes do not fit the fp16 operand scales, 2 / 3 = more candidates inside the error band than
 * the stage-1 lists / stage-2 buffer hold (mass duplicates); 4 = the scan started from a SAMPLED threshold
 * (large banks: the j-th smallest distance of a strided sample of the bank, j chosen so that the bank
 * holds k clips under it except with probability < 1e-10 for a bank in no particular order) and stage 2
 * found fewer than k clips under it.  Flagged queries must be re-run through vfr_score_topk by the caller
 * (vfr_b200.retrieval does): a flag costs time, never correctness.  VFR_SEL_SAMPLE=0 turns the sample
 * pass off. */
size_t vfr_sel_bank_bytes(int64_t n_clips, int dim);
int vfr_sel_bank_pack(const float* bank, int64_t n_clips, int dim, void* packed, vfr_stream_t stream);
size_t vfr_sel_query_bytes(int64_t n_queries, int dim);
int vfr_sel_query_pack(const float* queries, int64_t n_queries, int dim, const void* bank_packed,
                       int64_t n_clips, void* packed, vfr_stream_t stream);
size_t vfr_sel_topk_bytes(int64_t n_queries, int64_t n_clips, int n_split);
int vfr_sel_topk(const void* bank_packed, const float* bank, const int32_t* vid_off, const int64_t* mom_off,
                 int64_t n_videos, int64_t n_clips, int n_max, int dim, void* query_packed,
                 const float* queries, int64_t n_queries, int k, int64_t id_base, float* out_scores,
                 int64_t* out_ids, void* workspace, int n_split, vfr_stream_t stream);
const int32_t* vfr_sel_flags(const void* query_packed, int64_t n_queries, int dim);
/* bf16 embedding path (BASELINE configs[2]): vfr_sel_topk with the bank's rows stored as bf16 [n_clips, dim] - stage 2 then
 * reads half the bytes.  The packed operand must have been built (vfr_sel_bank_pack) from the same bf16-representable
 * values, and the query rows are expected to be bf16-representable too; the scores are the exact engine's scores OF THOSE
 * ROUNDED EMBEDDINGS bit for bit, i.e. within the stated 1e-2 of the fp32 embeddings' scores. */
int vfr_sel_topk_b16(const void* bank_packed, const void* bank_b16, const int32_t* vid_off, const int64_t* mom_off,
                     int64_t n_videos, int64_t n_clips, int n_max, int dim, void* query_packed, const float* queries,
                     int64_t n_queries, int k, int64_t id_base, float* out_scores, int64_t* out_ids, void* workspace,
                     int n_split, vfr_stream_t stream);
/* The two stages separately, for a bank sharded over several GPUs.  Every shard's local top-k only has to contain
 * what can reach the GLOBAL top-k, so the shards exchange a bound half way: vfr_sel_filter over a first slice of
 * the shard's bank tiles (tiles of 256 clips, vfr_sel_tiles(n_clips) in total; resume = 0 starts fresh lists) ->
 * vfr_sel_bound_get (bound[q] = a CERTIFIED upper bound of the exact k-th smallest squared clip distance of THIS
 * shard: from k keys really seen, never from the sampled threshold; +inf until a list has been compacted) ->
 * all-reduce(min) over the shards -> vfr_sel_bound_put -> vfr_sel_filter over the remaining tiles with
 * resume = 1 -> vfr_sel_refine.  vfr_sel_topk == filter(all tiles) + refine. */
/* A tighter protocol for large sharded banks (what vfr_b200.retrieval uses when every shard is big enough to sample):
 * the shards pool their SAMPLES instead of their bounds.  vfr_sel_sample starts fresh lists and writes, per query and
 * candidate list (vfr_sel_sample_lists of them), the 32 smallest sampled 64-clip minima as upper bounds of exact squared
 * distances, ascending (+inf padded); *n_sampled = clips sampled per query (0: shard too small, use the protocol above).
 * All-gather the 32 smallest per shard, take T[q] = the j-th smallest of the union with
 * j = vfr_sel_sample_rank(k, total sampled, total clips) (0: no rank <= 32 is safe), vfr_sel_bound_put(T), vfr_sel_filter
 * over ALL tiles with resume = 1, then CHECK the guess: vfr_sel_count_under(T) counts the shard's clips that are certainly
 * within T; where the all-reduced (sum) count is < k the query must be treated as flagged (the sample promised k clips the
 * bank does not have).  Then vfr_sel_refine.  Every shard then keeps ~k/P candidates instead of ~k. */
int vfr_sel_sample_rank(int k, int64_t n_sampled, int64_t n_total);
int vfr_sel_sample_lists(int64_t n_queries, int64_t n_clips, int n_split, int dim);
int64_t vfr_sel_sample_clips(int64_t n_queries, int64_t n_clips, int k, int n_split, int dim);   /* what vfr_sel_sample will report */
int vfr_sel_sample(const void* bank_packed, int64_t n_clips, int dim, void* query_packed, int64_t n_queries, int k,
                   void* workspace, int n_split, float* out, int64_t* n_sampled, vfr_stream_t stream);
int vfr_sel_count_under(void* query_packed, int64_t n_queries, int64_t n_clips, int dim, int k, void* workspace,
                        int n_split, const float* bound, int32_t* count, vfr_stream_t stream);
int64_t vfr_sel_tiles(int64_t n_clips);
/* The shard-level glue of the pooled-sample protocol as kernels (no host round trip, no library ops):
 *  vfr_sel_pool_levels   pooled fp32 [n_src, Q, width] (the all-gathered samples, +inf padded) -> levels fp32 [n_levels, Q],
 *                        levels[l][q] = the ranks[l]-th smallest pooled value of query q (ranks: HOST int32 [n_levels <= 4],
 *                        1-based, descending = loosest level first; n_src * width <= 1024; sorted_runs != 0 asserts that
 *                        every run of 32 values is ascending, as vfr_sel_sample exports them: a shortcut, same result);
 *  vfr_sel_count_levels  count int32 [n_levels, Q] = the shard's clips certainly within each level (vfr_sel_count_under for
 *                        all levels in one pass) - all-reduce(sum) them over the shards;
 *  vfr_sel_pick_put      the tightest level whose all-reduced count is >= k becomes the certified bound (vfr_sel_bound_put);
 *                        queries whose loosest level fails are flagged 4;
 *  vfr_sel_refine_blocks vfr_sel_refine writing the QUERY-SLICE records of the exchange: queries are cut into slices of
 *                        `per` (slice j belongs to rank j), record j = {ids int64 [per, k] | scores fp32 [per, k] | flags
 *                        int32 [per]} at out_blocks + j * vfr_topk_block_bytes(per, k); per * (3k + 1) must be even.
 *                        ONE all-to-all then hands every rank the P shard records of its own slice (vfr_topk_merge_blocks). */
int vfr_sel_pool_levels(const float* pooled, int n_src, int64_t n_queries, int width, const int32_t* ranks, int n_levels,
                        int sorted_runs, float* levels, vfr_stream_t stream);
int vfr_sel_count_levels(void* query_packed, int64_t n_queries, int64_t n_clips, int dim, int k, void* workspace, int n_split,
                         const float* levels, int n_levels, int32_t* count, vfr_stream_t stream);
int vfr_sel_pick_put(void* query_packed, int64_t n_queries, int64_t n_clips, int dim, int k, void* workspace, int n_split,
                     const float* levels, const int32_t* count, int n_levels, vfr_stream_t stream);
size_t vfr_topk_block_bytes(int64_t per, int k);
int vfr_sel_refine_blocks(const float* bank, const int32_t* vid_off, const int64_t* mom_off, int64_t n_videos, int64_t n_clips,
                          int n_max, int dim, void* query_packed, const float* queries, int64_t n_queries, int k,
                          int64_t id_base, int64_t per, void* out_blocks, void* workspace, int n_split, vfr_stream_t stream);
/* K7 on the records of the query-slice exchange: merges the first n_rows queries of n_parts records of one slice into
 * out_scores / out_ids [n_rows, k]; out_flags int32 [n_rows] = OR of the shards' flags, *n_flagged (DEVICE counter, may be
 * NULL) += number of flagged queries. */
int vfr_topk_merge_blocks(const void* blocks, int n_parts, int64_t per, int64_t n_rows, int k, float* out_scores,
                          int64_t* out_ids, int32_t* out_flags, int64_t* n_flagged, vfr_stream_t stream);
/* diagnostics of the last filter on this workspace (what bench.py reports next to the throughput): out = DEVICE
 * int64 [8] = {sum of the candidate-list lengths, longest list, flagged queries, lists, warp-level compaction events,
 * lists compacted, 0, 0}. */
int vfr_sel_stats(void* query_packed, int64_t n_queries, int64_t n_clips, int dim, int k, void* workspace, int n_split,
                  int64_t* out, vfr_stream_t stream);
int vfr_sel_filter(const void* bank_packed, int64_t n_clips, int dim, void* query_packed, int64_t n_queries, int k,
                   void* workspace, int n_split, int64_t tile_lo, int64_t tile_hi, int resume, vfr_stream_t stream);
int vfr_sel_bound_get(void* query_packed, int64_t n_queries, int64_t n_clips, int dim, int k, void* workspace,
                      int n_split, float* bound, vfr_stream_t stream);
int vfr_sel_bound_put(void* query_packed, int64_t n_queries, int64_t n_clips, int dim, int k, void* workspace,
                      int n_split, const float* bound, vfr_stream_t stream);
int vfr_sel_refine(const float* bank, const int32_t* vid_off, const int64_t* mom_off, int64_t n_videos,
                   int64_t n_clips, int n_max, int dim, void* query_packed, const float* queries, int64_t n_queries,
                   int k, int64_t id_base, float* out_scores, int64_t* out_ids, void* workspace, int n_split,
                   vfr_stream_t stream);
/* vfr_sel_refine for the queries [q_begin, q_begin + q_count) only (rows q of out_scores / out_ids): finished rows can
 * travel to the host while the rest is still being re-scored (what vfr_search_host does). */
int vfr_sel_refine_range(const float* bank, const int32_t* vid_off, const int64_t* mom_off, int64_t n_videos,
                         int64_t n_clips, int n_max, int dim, void* query_packed, const float* queries, int64_t n_queries,
                         int k, int64_t id_base, float* out_scores, int64_t* out_ids, void* workspace, int n_split,
                         int64_t q_begin, int64_t q_count, vfr_stream_t stream);

/* ---- K5 : integer-exact temporal IoU, ground truth, rank statistics -------------------------
 * times int32 [Q, n_annot, 2] inclusive (start, end), absent annotators = (-1, -1);
 * q_nseg int32 [Q] = clips of the query's own video; own_scores fp32 [Q, m_stride] (vfr_score_own);
 * tables uint8 [n_thr, 65, 65]: tables[t][inter][union] = (inter/union > thr_t) built on the host
 * in float64 (vfr_b200.utils.threshold_table) - replaces model/utils.py:78-82. */

/* ground-truth rule of model/evaluate.py:59-65: gt uint8 [Q, n_thr, m_stride]; per (q, t) the
 * smallest positive score tau (+inf if none), its moment index pos (-1 if none, lowest index on
 * ties), the number of positives npos, and eq_before = #{m < pos : own_scores[m] == tau}. */
int vfr_gt_select(const float* own_scores, int m_stride, const int32_t* q_nseg, const int32_t* times,
                  int n_annot, const uint8_t* tables, int n_thr, int64_t n_queries, uint8_t* gt,
                  float* tau, int32_t* pos, int32_t* npos, int32_t* eq_before, vfr_stream_t stream);

/* ranking of one video's moments (model/evaluate_single.py:52-54): order int32 [Q, m_stride] =
 * moment indices sorted by ascending (score, index); descending != 0 reverses that list, which
 * is what the reference's "[::-1]" does.  -1 padded.  m_stride <= 1024. */
int vfr_rank_order(const float* own_scores, int m_stride, const int32_t* q_nseg, int64_t n_queries,
                   int descending, int32_t* order, vfr_stream_t stream);

/* rank statistics of model/evaluate_single.py:58-73 for a ranked list: ranks int32 [Q, n_annot]
 * (1-based position of each annotated time; 0 = absent annotator, -1 = not a candidate moment),
 * integer IoU (top1_inter / top1_union [Q, n_annot]) between the top-1 moment and each annotation,
 * first_pos int32 [Q, n_thr] = position of the first positive moment in the list (m_stride if none). */
int vfr_single_metrics(const int32_t* order, int m_stride, const int32_t* q_nseg, const int32_t* times,
                       int n_annot, const uint8_t* tables, int n_thr, int64_t n_queries, int32_t* ranks,
                       int32_t* top1_inter, int32_t* top1_union, int32_t* first_pos, vfr_stream_t stream);

/* ---- K2 : visual embedding ------------------------------------------------------------------
 * replaces model/models.py:21-27,55-56 in eval mode: out = relu(x W1^T + b1) W2^T + b2, fp32.
 * x [n_rows, in_dim] is the reference's [segment | context | tef] row (model/data.py:204-213);
 * hidden is a caller workspace fp32 [n_rows, hid]. */
int vfr_linear(const float* x, int64_t n_rows, int in_dim, int ldx, const float* w, const float* bias,
               int out_dim, int relu, float* out, int ldo, vfr_stream_t stream);
int vfr_visual_embed(const float* x, int64_t n_rows, int in_dim, const float* w1, const float* b1, int hid,
                     const float* w2, const float* b2, int dim, float* hidden, float* out,
                     vfr_stream_t stream);

/* K2 on tensor cores (csrc/vfr_visual_tc.cu): split-fp16 tcgen05 GEMMs (22 significand bits per operand, power-of-two
 * operand scales computed on the device), bias / ReLU / tef columns fused into the epilogues; within 1e-5 of the fp32
 * reference at K = 8194.  `packed` = vfr_visual_pack of (W1 [hid, 2F+2], b1, W2 [dim, hid], b2), once per weight update.
 *  vfr_visual_embed_tc     x fp32 [n_rows, 2F+2] = the reference's assembled [segment | context | tef] rows -> out [n_rows, dim]
 *  vfr_visual_embed_split  the SPLIT-WEIGHT form of model/data.py:204-213 + models.py:21-27: seg fp32 [n_clips, F] and ctx
 *                          fp32 [n_videos, F] as K1 produces them, vid_off int32 [n_videos + 1] CSR clip offsets.  The
 *                          8194-wide concat is never materialised; the context product is computed once per video and
 *                          added per clip in the epilogue together with tef = (i/n, (i+1)/n) and the bias.
 * workspace: vfr_visual_embed_tc_bytes(n_rows, n_videos, F, hid, dim, split). */
size_t vfr_visual_pack_bytes(int feat_dim, int hid, int dim);
int vfr_visual_pack(const float* w1, const float* b1, const float* w2, const float* b2, int feat_dim, int hid, int dim,
                    void* packed, vfr_stream_t stream);
size_t vfr_visual_embed_tc_bytes(int64_t n_rows, int64_t n_videos, int feat_dim, int hid, int dim, int split);
int vfr_visual_embed_tc(const float* x, int64_t n_rows, int feat_dim, const void* packed, int hid, int dim, void* workspace,
                        float* out, vfr_stream_t stream);
int vfr_visual_embed_split(const float* seg, const float* ctx, const int32_t* vid_off, int64_t n_clips, int64_t n_videos,
                           int feat_dim, const void* packed, int hid, int dim, void* workspace, float* out,
                           vfr_stream_t stream);

/* ---- K3 : query embedding (GloVe gather -> BiLSTM -> Linear) -----------------------------------
 * replaces model/models.py:33-48,61-66.  Weights of one direction are re-laid-out once per weight
 * update by vfr_lstm_pack ([W_hh | W_ih] rows interleaved by gate, bias = b_ih + b_hh;
 * vfr_lstm_pack_bytes bytes).  tokens int64 [B, L]; table fp32 [vocab, emb] (row 0 = pad);
 * length_table fp32 [vocab] or NULL (the normalize_lang variant, models.py:62-64);
 * fc_w fp32 [dim, 2*hidden]; out fp32 [B, dim]; workspace vfr_text_embed_bytes bytes whose FIRST
 * int32 is set to 1 if a token id was outside [0, vocab) (the reference raises IndexError). */
size_t vfr_lstm_pack_bytes(int hidden, int emb);
int vfr_lstm_pack(const float* w_ih, const float* w_hh, const float* b_ih, const float* b_hh, int hidden,
                  int emb, float* packed, vfr_stream_t stream);
size_t vfr_text_embed_bytes(int64_t n_queries, int seq_len, int hidden, int emb);
int vfr_text_embed(const int64_t* tokens, int64_t n_queries, int seq_len, const float* table, int64_t vocab,
                   const float* length_table, int emb, const float* packed_fwd, const float* packed_bwd,
                   int hidden, const float* fc_w, const float* fc_b, int dim, void* workspace, float* out,
                   vfr_stream_t stream);

/* ---- K2 / K3, tensor-core path: split-bf16 GEMMs on tcgen05 (fp32 accuracy) ------------------------
 * Every fp32 operand is stored as a bf16 (hi, lo) pair and the product is accumulated in fp32 in TMEM
 * as Ah.Bh + Al.Bh + Ah.Bl (csrc/vfr_gemm_tc.cuh).
 * vfr_linear_tc: out = act(x W^T + b) with W packed once by vfr_tc_weight_pack; workspace
 * vfr_linear_tc_bytes(n_rows, in_dim) holds the split copy of x. */
size_t vfr_tc_weight_bytes(int out_dim, int in_dim);
int vfr_tc_weight_pack(const float* w, int out_dim, int in_dim, void* packed, vfr_stream_t stream);
size_t vfr_linear_tc_bytes(int64_t n_rows, int in_dim);
int vfr_linear_tc(const float* x, int64_t n_rows, int in_dim, int64_t ldx, const void* w_packed,
                  const float* bias, int out_dim, int relu, float* out, int64_t ldo, void* workspace,
                  vfr_stream_t stream);
/* K3 on tensor cores: same contract as vfr_text_embed; `packed` (vfr_text_pack_tc_bytes) holds both
 * LSTM directions and lang_fc, re-packed once per weight update; workspace vfr_text_embed_tc_bytes
 * whose FIRST int32 flags out-of-range token ids. */
size_t vfr_text_pack_tc_bytes(int hidden, int emb, int dim);
int vfr_text_pack_tc(const float* w_ih_f, const float* w_hh_f, const float* b_ih_f, const float* b_hh_f,
                     const float* w_ih_b, const float* w_hh_b, const float* b_ih_b, const float* b_hh_b,
                     const float* fc_w, const float* fc_b, int hidden, int emb, int dim, void* packed,
                     vfr_stream_t stream);
size_t vfr_text_embed_tc_bytes(int64_t n_queries, int seq_len, int hidden, int emb);
int vfr_text_embed_tc(const int64_t* tokens, int64_t n_queries, int seq_len, const float* table,
                      int64_t vocab, const float* length_table, int emb, const void* packed, int hidden,
                      int dim, void* workspace, float* out, vfr_stream_t stream);

/* ---- K1 : frame -> segment pooling --------------------------------------------------------------
 * replaces model/data.py:142-188.  frames fp32 [sum_v F_v, dim] (the get_rgb_features.py .npy rows
 * of all videos back to back), frame_off int64 [V+1].  mode 0 = avg, 1 = max (.npy branch,
 * data.py:163-181), 2 = the preprocessed-.h5 branch (data.py:144-161).  Outputs, L2-normalised
 * x/(|x|+1e-5): seg fp32 [V, seg_stride, dim] (rows >= n_seg[v] zeroed), ctx fp32 [V, dim],
 * n_seg int32 [V].  dim % 4 == 0.  workspace: vfr_segment_pool_bytes. */
size_t vfr_segment_pool_bytes(int64_t n_videos, int dim, int seg_stride);
int vfr_segment_pool(const float* frames, const int64_t* frame_off, int64_t n_videos, int dim, int window,
                     int mode, float* seg, int seg_stride, float* ctx, int32_t* n_seg, void* workspace,
                     vfr_stream_t stream);

/* ---- K6 : ranking loss (forward + backward) -------------------------------------------------------
 * replaces model/main.py:214-232 and its autograd graph.  posit/inter fp32 [rows_posit, dim] with
 * sample ids maskp int64 [rows_posit]; intra fp32 [rows_intra, dim] with maskn; lang fp32
 * [n_samples, dim].  loss_out: device scalar (a SUM over samples).  The same workspace
 * (vfr_ranking_loss_bytes) must be passed to bwd after fwd.  grad_out: device scalar or NULL (=1). */
size_t vfr_ranking_loss_bytes(int rows_posit, int rows_intra, int rows_inter, int n_samples);
int vfr_ranking_loss_fwd(const float* posit, const float* intra, const float* inter, const float* lang,
                         const int64_t* maskp, const int64_t* maskn, int rows_posit, int rows_intra,
                         int rows_inter, int n_samples, int dim, int normalize, float b, float lamb,
                         void* workspace, float* loss_out, vfr_stream_t stream);
int vfr_ranking_loss_bwd(const float* posit, const float* intra, const float* inter, const float* lang,
                         const int64_t* maskp, const int64_t* maskn, int rows_posit, int rows_intra,
                         int rows_inter, int n_samples, int dim, int normalize, float b, float lamb,
                         void* workspace, const float* grad_out, float* grad_posit, float* grad_intra,
                         float* grad_inter, float* grad_lang, vfr_stream_t stream);

/* ---- training step: forward-with-saved-activations, BACKWARD of both embedding branches, fused Adam ---------------
 * replaces the autograd graph / library calls behind  loss.backward(); optimizer.step()  of model/main.py:57-67 (optimiser
 * main.py:358, grad-norm logging utils.py:85-92); the ranking loss between them is K6.  fp32 on the CUDA cores: a training
 * step (R ~ 120 clip rows x 3 streams, B ~ 87 queries) is bound by launch latency and the 20 sequential recurrent steps.
 *
 * Text branch (models.py:61-66, train mode has no dropout here): w_ih / w_hh / b_ih / b_hh and their gradients are HOST
 * arrays of TWO device pointers {forward direction, reverse direction} in nn.LSTM's own layout ([4H, E], [4H, H], [4H],
 * gate order i, f, g, o).  The forward keeps every activation in `workspace` (vfr_text_train_bytes) for the backward call;
 * vfr_text_train_flag(workspace) -> DEVICE int32, 1 if a token id was outside [0, vocab).  d_length (fp32 [vocab], must be
 * zeroed by the caller, accumulated into) is the gradient of the learnable word length (normalize_lang), else NULL. */
size_t vfr_text_train_bytes(int n_queries, int seq_len, int hidden, int emb, int dim);
int vfr_text_train_fwd(const int64_t* tokens, int n_queries, int seq_len, const float* table, int64_t vocab,
                       const float* length_table, int emb, const float* const* w_ih, const float* const* w_hh,
                       const float* const* b_ih, const float* const* b_hh, int hidden, const float* fc_w, const float* fc_b,
                       int dim, void* workspace, float* out, vfr_stream_t stream);
int vfr_text_train_bwd(const int64_t* tokens, int n_queries, int seq_len, int64_t vocab, int has_length, int emb,
                       const float* const* w_ih, const float* const* w_hh, int hidden, const float* fc_w, int dim,
                       void* workspace, const float* grad_out, float* const* d_w_ih, float* const* d_w_hh,
                       float* const* d_b_ih, float* const* d_b_hh, float* d_fc_w, float* d_fc_b, float* d_length,
                       vfr_stream_t stream);
const int32_t* vfr_text_train_flag(const void* workspace);
/* Visual branch (models.py:21-27): e = relu(x W1^T + b1) W2^T + b2.  vfr_visual_train_fwd is the forward for the ~120-row
 * batches of a training step (split-K fp32 SGEMMs over all SMs; scratch vfr_visual_train_fwd_bytes) and keeps the post-ReLU
 * hidden [n, hid]; vfr_visual_train_bwd takes x [n, in_dim], that hidden and dE [n, dim]; scratch fp32 [n, hid]; d_x optional. */
size_t vfr_visual_train_fwd_bytes(int64_t n_rows, int hid, int dim);
int vfr_visual_train_fwd(const float* x, int64_t n_rows, int in_dim, const float* w1, const float* b1, int hid, const float* w2,
                         const float* b2, int dim, float* scratch, float* hidden, float* out, vfr_stream_t stream);
int vfr_visual_train_bwd(const float* x, int64_t n_rows, int in_dim, const float* hidden, int hid, const float* w1,
                         const float* w2, int dim, const float* grad_out, float* scratch, float* d_w1, float* d_b1,
                         float* d_w2, float* d_b2, float* d_x, vfr_stream_t stream);
/* torch.optim.Adam(lr, weight_decay) of main.py:358 for up to 24 tensors in ONE launch (L2 decay folded into the gradient,
 * bias-corrected moments, eps after the sqrt); pointer arrays and numel are HOST arrays, step = 1-based count of this update.
 * vfr_grad_norms: out fp32 [count] (DEVICE) = L2 norm of every gradient tensor (utils.py:85-92 logs their mean). */
int vfr_adam_step(float* const* params, const float* const* grads, float* const* exp_avg, float* const* exp_avg_sq,
                  const int64_t* numel, int count, int64_t step, float lr, float beta1, float beta2, float eps,
                  float weight_decay, vfr_stream_t stream);
int vfr_grad_norms(const float* const* grads, const int64_t* numel, int count, float* out, vfr_stream_t stream);

/* ---- training-batch construction on the device, pooled-moment features -------------------------
 * vfr_sample_negatives: the per-query draws of CustomBatchSampler.__iter__ (model/data.py:275-337) for ALL queries of an epoch
 * in one launch: times int32 [Q, n_annot, 2] ((-1,-1) padded), q_video int32 [Q], nseg int32 [V] -> out int32 [Q, 8] =
 * {video_pos, video_neg, start_t, end_t, start_tn, end_tn, status, 0}; status 1 = no intra-video candidate (the reference's
 * random.choice([]) raises IndexError there), 2 = no other video long enough, 3 = no usable annotation.  Counter-based RNG:
 * a pure function of (seed, epoch, query) - the reference's distributions, not its Mersenne-Twister stream.
 * vfr_gather_clip_rows: out fp32 [n_rows, 2F+2] = [seg[vid_off[v] + c] | ctx[v] | (c/n, (c+1)/n)] for (v, c) = (row_video[r],
 * row_clip[r]) - model/data.py:204-213 straight from the pooled features in HBM.
 * vfr_moment_pool: out fp32 [M_total, F], row mom_off[v] + moment_index = the MEAN segment feature of that moment (shared-
 * memory prefix sums per column): MCN-style pooled-moment features, an additional NON-reference scoring variant. */
int vfr_sample_negatives(const int32_t* times, int n_annot, const int32_t* q_video, const int32_t* nseg, int64_t n_queries,
                         int64_t n_videos, int same_length, uint64_t seed, uint64_t epoch, int32_t* out, vfr_stream_t stream);
int vfr_gather_clip_rows(const float* seg, const float* ctx, const int32_t* vid_off, const int32_t* row_video,
                         const int32_t* row_clip, int64_t n_rows, int feat_dim, float* out, vfr_stream_t stream);
int vfr_moment_pool(const float* seg, const int32_t* vid_off, const int64_t* mom_off, int64_t n_videos, int n_max, int feat_dim,
                    float* out, vfr_stream_t stream);

/* ---- retrieval step: K3 -> K4 behind one call --------------------------------------------------
 * The serving form of model/evaluate.py:42-80: one batch of tokenised queries against the resident
 * bank.  All pointers inside the plan are DEVICE pointers supplied by the caller (weights as for
 * vfr_text_embed, bank as for vfr_score_topk, scratch sized for max_queries):
 *   tokens_dev int64 [max_queries, seq_len]; q_emb fp32 [max_queries, dim];
 *   q_packed vfr_query_pack_bytes(max_queries, dim); text_ws vfr_text_embed_bytes(max_queries, ...);
 *   topk_ws vfr_score_topk_bytes(max_queries, n_split); out_*_dev [max_queries, k].
 * bank_packed / q_packed are only needed by engine 0. */
typedef struct vfr_search_plan {
  const float* table; int64_t vocab; const float* length_table; int emb;
  const float* lstm_fwd; const float* lstm_bwd; int hidden;
  const float* fc_w; const float* fc_b; int dim; int seq_len;
  const float* bank_packed; const int32_t* vid_off; const int64_t* mom_off;
  int64_t n_videos; int n_max; int64_t id_base;
  int64_t* tokens_dev; float* q_emb; float* q_packed; void* text_ws; void* topk_ws;
  float* out_scores_dev; int64_t* out_ids_dev; int n_split; int64_t max_queries;
  /* scoring engine: 0 = exact-fp32 CUDA-core path (vfr_score_topk); 3 = tensor-core split-bf16 path,
   * 1 = tensor-core plain-bf16 path (vfr_score_topk_tc with n_terms = engine).  Engines 1/3 need:
   * bank_tc (vfr_tc_bank_pack), bank_clips fp32 [C, dim] (exact fallback), uniform6, q_tc
   * (vfr_tc_query_bytes(max_queries)) and topk_ws sized by vfr_score_topk_tc_bytes. */
  int engine; const void* bank_tc; const float* bank_clips; int uniform6; void* q_tc;
  /* text engine: 0 = exact-fp32 CUDA-core K3 (lstm_fwd/lstm_bwd/fc_w/fc_b, text_ws =
   * vfr_text_embed_bytes); 3 = tensor-core K3 (text_tc = vfr_text_pack_tc blob, text_ws =
   * vfr_text_embed_tc_bytes). */
  int text_engine; const void* text_tc;
  /* engine 4 = filter + refine top-k (vfr_sel_topk): bank_tc = vfr_sel_bank_pack blob, q_tc =
   * vfr_sel_query_bytes(max_queries), topk_ws = vfr_sel_topk_bytes(max_queries, n_clips, n_split),
   * bank_clips fp32 [n_clips, dim]; any clip count per video <= 32.  Per-query flags: vfr_sel_flags(q_tc, n). */
  int64_t n_clips;
} vfr_search_plan;

/* the two stages separately (multi-GPU: every rank embeds its slice of the batch, the embeddings are
 * all-gathered into plan->q_emb, then every rank scores the whole batch against its bank shard) */
int vfr_search_embed_device(const vfr_search_plan* plan, const int64_t* tokens_dev, int64_t n_queries,
                            float* q_emb_out, vfr_stream_t stream);
int vfr_search_score_device(const vfr_search_plan* plan, int64_t n_queries, int k, float* out_scores_dev,
                            int64_t* out_ids_dev, vfr_stream_t stream);
/* device-resident inputs/outputs; asynchronous */
int vfr_search_device(const vfr_search_plan* plan, const int64_t* tokens_dev, int64_t n_queries, int k,
                      float* out_scores_dev, int64_t* out_ids_dev, vfr_stream_t stream);
/* HOST token ids in, HOST top-k out (copies on `stream` inside the call); returns after the
 * results have landed (synchronises the stream). */
int vfr_search_host(const vfr_search_plan* plan, const int64_t* tokens_host, int64_t n_queries, int k,
                    float* out_scores_host, int64_t* out_ids_host, vfr_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* VFR_H_ */
