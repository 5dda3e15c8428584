"""Thin torch-tensor wrappers over the libvfr C ABI: argument checking, workspace allocation
(torch owns device memory and streams - plumbing only) and pointer extraction.  Every function
requires CUDA tensors and raises otherwise: there is no CPU path in the product."""
import ctypes as C

import numpy as np
import torch

from . import _lib
from .utils import MAX_SEGMENTS, threshold_table

SEL_MAX_DIM = 1085  # largest embedding dimension of the filter + refine engine (vfr_sel_*: rows of up to 1088 fp16)


def _ptr(t):
    return None if t is None else t.data_ptr()


def _stream():
    return torch.cuda.current_stream().cuda_stream


def _need_cuda(*tensors):
    for t in tensors:
        if t is not None and not t.is_cuda:
            raise _lib.VfrError("vfr_b200 ops need CUDA tensors (no CPU fallback)")


def _f32c(t):
    return t.detach().to(torch.float32).contiguous()


class Bank:
    """Clip embeddings of V videos resident in HBM, in the layouts the kernels consume.

    ``clips`` fp32 [C, D]; ``vid_off`` int32 [V+1] CSR clip offsets (host + device copies).
    Holds the k-major packed copy used by the scoring kernels (built once, like the reference's
    ``videos`` dict of model/evaluate.py:31-35 which is also computed once and kept)."""

    def __init__(self, clips, vid_off):
        _need_cuda(clips)
        self.clips = _f32c(clips)
        vo = np.asarray(vid_off, dtype=np.int64)
        if vo.ndim != 1 or vo[0] != 0 or vo[-1] != self.clips.shape[0] or np.any(np.diff(vo) < 1):
            raise ValueError("vid_off must be increasing CSR offsets covering all clips")
        nseg = np.diff(vo)
        if nseg.max() > MAX_SEGMENTS:
            raise _lib.VfrError(f"videos with more than {MAX_SEGMENTS} clips are not supported")
        self.n_videos = len(nseg)
        self.n_max = int(nseg.max())
        self.dim = int(self.clips.shape[1])
        self.nseg_host = nseg.astype(np.int32)
        self.vid_off_host = vo
        mo = np.concatenate([[0], np.cumsum(nseg * (nseg + 1) // 2)]).astype(np.int64)
        self.mom_off_host = mo
        self.m_total = int(mo[-1])
        if self.m_total >= 2 ** 32:
            raise _lib.VfrError("more than 2^32 moments per bank shard")
        dev = self.clips.device
        self.vid_off = torch.from_numpy(vo.astype(np.int32)).to(dev)
        self.mom_off = torch.from_numpy(mo).to(dev)
        nbytes = _lib.load().vfr_bank_pack_bytes(self.n_videos, self.n_max, self.dim)
        if nbytes == 0:
            raise _lib.VfrError("vfr_bank_pack_bytes: unsupported bank shape")
        self.packed = torch.empty(nbytes // 4, dtype=torch.float32, device=dev)
        _lib.call("vfr_bank_pack", _ptr(self.clips), _ptr(self.vid_off), self.n_videos, self.n_max, self.dim,
                  _ptr(self.packed), _stream())

    @property
    def device(self):
        return self.clips.device

    def tc(self, n_terms=3):
        """Packed operand of the tensor-core scoring path (built lazily, cached per n_terms)."""
        cache = self.__dict__.setdefault("_tc", {})
        if n_terms not in cache:
            if self.n_max > 6:
                raise _lib.VfrError("the tensor-core scoring path holds videos of at most 6 clips")
            nbytes = _lib.load().vfr_tc_bank_bytes(self.n_videos)
            packed = torch.empty(nbytes, dtype=torch.uint8, device=self.device)
            _lib.call("vfr_tc_bank_pack", _ptr(self.clips), _ptr(self.vid_off), self.n_videos, self.n_max, self.dim,
                      n_terms, _ptr(packed), _stream())
            cache[n_terms] = packed
        return cache[n_terms]

    def sel(self):
        """Packed fp16 operand of the filter + refine top-k (vfr_sel_topk); built lazily, cached."""
        if "_sel" not in self.__dict__:
            if self.dim > SEL_MAX_DIM:
                raise _lib.VfrError(f"the filter + refine top-k holds embeddings of at most {SEL_MAX_DIM} dimensions")
            n_clips = int(self.clips.shape[0])
            packed = torch.empty(_lib.load().vfr_sel_bank_bytes(n_clips, self.dim), dtype=torch.uint8, device=self.device)
            _lib.call("vfr_sel_bank_pack", _ptr(self.clips), n_clips, self.dim, _ptr(packed), _stream())
            self.__dict__["_sel"] = packed
        return self.__dict__["_sel"]

    def clips_b16(self):
        """bf16 copy of the clip embeddings (the bf16 embedding path); the bank must hold bf16-representable values."""
        if "_b16" not in self.__dict__:
            b16 = self.clips.to(torch.bfloat16)
            if not torch.equal(b16.to(torch.float32), self.clips):
                raise _lib.VfrError("the bf16 path needs a bank of bf16-representable embeddings: build it from clips.bfloat16()")
            self.__dict__["_b16"] = b16.contiguous()
        return self.__dict__["_b16"]

    @property
    def uniform6(self):
        return int(bool((self.nseg_host == 6).all()))


def pack_queries(queries):
    _need_cuda(queries)
    q = _f32c(queries)
    nbytes = _lib.load().vfr_query_pack_bytes(q.shape[0], q.shape[1])
    packed = torch.empty(nbytes // 4, dtype=torch.float32, device=q.device)
    _lib.call("vfr_query_pack", _ptr(q), q.shape[0], q.shape[1], _ptr(packed), _stream())
    return packed


def score_full(bank, queries):
    """All scores in the reference's order (model/evaluate.py:49-58) -> fp32 [Q, M_total]."""
    qp = pack_queries(queries)
    Q = queries.shape[0]
    out = torch.empty((Q, bank.m_total), dtype=torch.float32, device=bank.device)
    _lib.call("vfr_score_full", _ptr(bank.packed), _ptr(bank.vid_off), _ptr(bank.mom_off), bank.n_videos,
              bank.n_max, bank.dim, _ptr(qp), Q, _ptr(out), bank.m_total, _stream())
    return out


def score_own(bank, queries, q_video):
    """Scores of each query against its own video -> fp32 [Q, m_stride] (+inf padded)."""
    _need_cuda(queries, q_video)
    q = _f32c(queries)
    qv = q_video.to(torch.int32).contiguous()
    m_stride = bank.n_max * (bank.n_max + 1) // 2
    out = torch.empty((q.shape[0], m_stride), dtype=torch.float32, device=bank.device)
    _lib.call("vfr_score_own", _ptr(bank.clips), _ptr(bank.vid_off), bank.dim, _ptr(q), q.shape[0], _ptr(qv),
              _ptr(out), m_stride, _stream())
    return out


def score_count(bank, queries, tau, q_video, n_split=0):
    """Rank counting -> (cnt_lt, cnt_eqb) int64 [Q, T]; see include/vfr.h."""
    _need_cuda(queries, tau, q_video)
    qp = pack_queries(queries)
    Q = queries.shape[0]
    tau = _f32c(tau).reshape(Q, -1)
    T = tau.shape[1]
    qv = q_video.to(torch.int32).contiguous()
    lt = torch.zeros((Q, T), dtype=torch.int32, device=bank.device)
    eqb = torch.zeros((Q, T), dtype=torch.int32, device=bank.device)
    _lib.call("vfr_score_count", _ptr(bank.packed), _ptr(bank.vid_off), bank.n_videos, bank.n_max, bank.dim,
              _ptr(qp), Q, _ptr(tau), T, _ptr(qv), _ptr(lt), _ptr(eqb), n_split, _stream())
    return lt.to(torch.int64) & 0xFFFFFFFF, eqb.to(torch.int64) & 0xFFFFFFFF


def score_topk(bank, queries, k, id_base=0, n_split=0):
    """Fused top-k -> (scores fp32 [Q, k], ids int64 [Q, k]) ascending by (score, id)."""
    _need_cuda(queries)
    qp = pack_queries(queries)
    Q = queries.shape[0]
    nbytes = _lib.load().vfr_score_topk_bytes(Q, n_split)
    ws = torch.empty(nbytes, dtype=torch.uint8, device=bank.device)
    out_s = torch.empty((Q, k), dtype=torch.float32, device=bank.device)
    out_i = torch.empty((Q, k), dtype=torch.int64, device=bank.device)
    _lib.call("vfr_score_topk", _ptr(bank.packed), _ptr(bank.vid_off), _ptr(bank.mom_off), bank.n_videos,
              bank.n_max, bank.dim, _ptr(qp), Q, k, id_base, _ptr(out_s), _ptr(out_i), _ptr(ws), n_split, _stream())
    return out_s, out_i


def pack_queries_tc(queries, n_terms=3):
    _need_cuda(queries)
    q = _f32c(queries)
    packed = torch.empty(_lib.load().vfr_tc_query_bytes(q.shape[0]), dtype=torch.uint8, device=q.device)
    _lib.call("vfr_tc_query_pack", _ptr(q), q.shape[0], q.shape[1], n_terms, _ptr(packed), _stream())
    return q, packed


def score_full_tc(bank, queries, n_terms=3):
    """Tensor-core path: all scores -> fp32 [Q, M_total] (moments of padded slots are not written)."""
    q, qp = pack_queries_tc(queries, n_terms)
    Q = q.shape[0]
    out = torch.full((Q, bank.m_total), float("nan"), dtype=torch.float32, device=bank.device)
    _lib.call("vfr_score_full_tc", _ptr(bank.tc(n_terms)), _ptr(bank.clips), _ptr(bank.vid_off), _ptr(bank.mom_off),
              bank.n_videos, bank.uniform6, bank.dim, n_terms, _ptr(qp), _ptr(q), Q, _ptr(out), bank.m_total, _stream())
    return out


def score_topk_tc(bank, queries, k, id_base=0, n_split=0, n_terms=3):
    """Tensor-core path: fused top-k -> (scores fp32 [Q, k], ids int64 [Q, k])."""
    q, qp = pack_queries_tc(queries, n_terms)
    Q = q.shape[0]
    ws = torch.empty(_lib.load().vfr_score_topk_tc_bytes(Q, bank.n_videos, n_split), dtype=torch.uint8, device=bank.device)
    out_s = torch.empty((Q, k), dtype=torch.float32, device=bank.device)
    out_i = torch.empty((Q, k), dtype=torch.int64, device=bank.device)
    _lib.call("vfr_score_topk_tc", _ptr(bank.tc(n_terms)), _ptr(bank.clips), _ptr(bank.vid_off), _ptr(bank.mom_off),
              bank.n_videos, bank.uniform6, bank.dim, n_terms, _ptr(qp), _ptr(q), Q, k, id_base, _ptr(out_s), _ptr(out_i),
              _ptr(ws), n_split, _stream())
    return out_s, out_i


def score_topk_sel(bank, queries, k, id_base=0, n_split=0, return_flags=False, bf16=False):
    """Filter + refine path (one fp16 tensor-core pass + exact fp32 re-scoring of the survivors):
    (scores fp32 [Q, k], ids int64 [Q, k]) bit-identical to ``score_topk``.  ``flags`` int32 [Q] is 0 for
    every query whose result is guaranteed exact (see include/vfr.h, vfr_sel_flags).
    ``bf16=True``: the bf16 embedding path - ``bank`` must hold bf16-representable values (``Bank(clips.bfloat16(), ...)``),
    the queries are rounded to bf16 and stage 2 reads a bf16 copy of the bank (``vfr_sel_topk_b16``)."""
    _need_cuda(queries)
    q = _f32c(queries)
    if bf16:
        q = q.to(torch.bfloat16).to(torch.float32).contiguous()
    Q = q.shape[0]
    lib = _lib.load()
    n_clips = int(bank.clips.shape[0])
    qp = torch.empty(lib.vfr_sel_query_bytes(Q, bank.dim), dtype=torch.uint8, device=bank.device)
    _lib.call("vfr_sel_query_pack", _ptr(q), Q, bank.dim, _ptr(bank.sel()), n_clips, _ptr(qp), _stream())
    ws = torch.empty(lib.vfr_sel_topk_bytes(Q, n_clips, n_split), dtype=torch.uint8, device=bank.device)
    out_s = torch.empty((Q, k), dtype=torch.float32, device=bank.device)
    out_i = torch.empty((Q, k), dtype=torch.int64, device=bank.device)
    if bf16:
        _lib.call("vfr_sel_topk_b16", _ptr(bank.sel()), _ptr(bank.clips_b16()), _ptr(bank.vid_off), _ptr(bank.mom_off),
                  bank.n_videos, n_clips, bank.n_max, bank.dim, _ptr(qp), _ptr(q), Q, k, id_base, _ptr(out_s), _ptr(out_i),
                  _ptr(ws), n_split, _stream())
    else:
        _lib.call("vfr_sel_topk", _ptr(bank.sel()), _ptr(bank.clips), _ptr(bank.vid_off), _ptr(bank.mom_off), bank.n_videos,
                  n_clips, bank.n_max, bank.dim, _ptr(qp), _ptr(q), Q, k, id_base, _ptr(out_s), _ptr(out_i), _ptr(ws),
                  n_split, _stream())
    if not return_flags:
        return out_s, out_i
    off = lib.vfr_sel_flags(_ptr(qp), Q, bank.dim) - qp.data_ptr()
    flags = qp[off:off + 4 * Q].view(torch.int32).clone()
    return out_s, out_i, flags, (qp, ws)


def topk_merge(scores, ids):
    """K7: merge [P, Q, k] per-shard lists -> [Q, k]."""
    _need_cuda(scores, ids)
    P, Q, k = scores.shape
    s = _f32c(scores)
    i = ids.to(torch.int64).contiguous()
    out_s = torch.empty((Q, k), dtype=torch.float32, device=s.device)
    out_i = torch.empty((Q, k), dtype=torch.int64, device=s.device)
    _lib.call("vfr_topk_merge", _ptr(s), _ptr(i), P, Q, k, _ptr(out_s), _ptr(out_i), _stream())
    return out_s, out_i


def pack_times(times_list, device):
    """Ragged annotator lists -> int32 [Q, A, 2] padded with (-1, -1)."""
    A = max(len(t) for t in times_list)
    arr = np.full((len(times_list), A, 2), -1, dtype=np.int32)
    for q, t in enumerate(times_list):
        if len(t):
            arr[q, :len(t)] = np.asarray(t, dtype=np.int32).reshape(-1, 2)
    return torch.from_numpy(arr).to(device)


def threshold_tables(thresholds, inclusive, device):
    tabs = np.stack([threshold_table(float(t), inclusive, 64) for t in thresholds])
    return torch.from_numpy(tabs).to(device)


def gt_select(own_scores, q_nseg, times, tables):
    """K5a -> gt uint8 [Q, T, m_stride], tau fp32 [Q, T], pos / npos / eq_before int32 [Q, T]."""
    _need_cuda(own_scores, q_nseg, times, tables)
    Q, ms = own_scores.shape
    T = tables.shape[0]
    dev = own_scores.device
    gt = torch.empty((Q, T, ms), dtype=torch.uint8, device=dev)
    tau = torch.empty((Q, T), dtype=torch.float32, device=dev)
    pos = torch.empty((Q, T), dtype=torch.int32, device=dev)
    npos = torch.empty((Q, T), dtype=torch.int32, device=dev)
    eqb = torch.empty((Q, T), dtype=torch.int32, device=dev)
    _lib.call("vfr_gt_select", _ptr(own_scores), ms, _ptr(q_nseg), _ptr(times), times.shape[1], _ptr(tables), T, Q,
              _ptr(gt), _ptr(tau), _ptr(pos), _ptr(npos), _ptr(eqb), _stream())
    return gt, tau, pos, npos, eqb


def rank_order(own_scores, q_nseg, descending):
    _need_cuda(own_scores, q_nseg)
    Q, ms = own_scores.shape
    order = torch.empty((Q, ms), dtype=torch.int32, device=own_scores.device)
    _lib.call("vfr_rank_order", _ptr(own_scores), ms, _ptr(q_nseg), Q, int(bool(descending)), _ptr(order), _stream())
    return order


def single_metrics(order, q_nseg, times, tables):
    _need_cuda(order, q_nseg, times, tables)
    Q, ms = order.shape
    A, T = times.shape[1], tables.shape[0]
    dev = order.device
    ranks = torch.empty((Q, A), dtype=torch.int32, device=dev)
    ti = torch.empty((Q, A), dtype=torch.int32, device=dev)
    tu = torch.empty((Q, A), dtype=torch.int32, device=dev)
    fp = torch.empty((Q, T), dtype=torch.int32, device=dev)
    _lib.call("vfr_single_metrics", _ptr(order.contiguous()), ms, _ptr(q_nseg), _ptr(times), A, _ptr(tables), T, Q,
              _ptr(ranks), _ptr(ti), _ptr(tu), _ptr(fp), _stream())
    return ranks, ti, tu, fp


# ---------------------------------------------------------------------------------------------
# K2 / K3 : embeddings
# ---------------------------------------------------------------------------------------------
def linear(x, weight, bias=None, relu=False):
    """fp32 ``x @ weight.T + bias`` (optionally ReLU) through vfr_linear."""
    _need_cuda(x, weight, bias)
    x = _f32c(x)
    w = _f32c(weight)
    b = None if bias is None else _f32c(bias)
    out = torch.empty((x.shape[0], w.shape[0]), dtype=torch.float32, device=x.device)
    _lib.call("vfr_linear", _ptr(x), x.shape[0], x.shape[1], x.shape[1], _ptr(w), _ptr(b), w.shape[0],
              int(bool(relu)), _ptr(out), w.shape[0], _stream())
    return out


def visual_embed(x, w1, b1, w2, b2, return_hidden=False):
    """K2: relu(x W1^T + b1) W2^T + b2 (reference model/models.py:21-26,56, eval mode)."""
    _need_cuda(x, w1, b1, w2, b2)
    x = _f32c(x)
    if x.dim() != 2 or x.shape[1] != w1.shape[1]:
        raise RuntimeError(f"mat1 and mat2 shapes cannot be multiplied ({tuple(x.shape)} and {tuple(w1.t().shape)})")
    w1, b1, w2, b2 = (_f32c(t) for t in (w1, b1, w2, b2))
    n = x.shape[0]
    hidden = torch.empty((n, w1.shape[0]), dtype=torch.float32, device=x.device)
    out = torch.empty((n, w2.shape[0]), dtype=torch.float32, device=x.device)
    if n:
        _lib.call("vfr_visual_embed", _ptr(x), n, x.shape[1], _ptr(w1), _ptr(b1), w1.shape[0], _ptr(w2), _ptr(b2),
                  w2.shape[0], _ptr(hidden), _ptr(out), _stream())
    return (out, hidden) if return_hidden else out


def visual_pack(w1, b1, w2, b2):
    """Pack the visual MLP for the tensor-core K2 (split-fp16 operands + the tef columns and biases in fp32)."""
    _need_cuda(w1, b1, w2, b2)
    w1, b1, w2, b2 = (_f32c(t) for t in (w1, b1, w2, b2))
    hid, dim = w1.shape[0], w2.shape[0]
    if (w1.shape[1] - 2) % 2 or w2.shape[1] != hid:
        raise _lib.VfrError(f"visual MLP shapes {tuple(w1.shape)} / {tuple(w2.shape)} are not [hid, 2F+2] / [dim, hid]")
    feat = (w1.shape[1] - 2) // 2
    packed = torch.empty(_lib.load().vfr_visual_pack_bytes(feat, hid, dim), dtype=torch.uint8, device=w1.device)
    _lib.call("vfr_visual_pack", _ptr(w1), _ptr(b1), _ptr(w2), _ptr(b2), feat, hid, dim, _ptr(packed), _stream())
    return packed, (feat, hid, dim)


def visual_embed_tc(x, packed, shape, rows_per_call=32768):
    """K2 on tensor cores, general form: x fp32 [N, 2F+2] (the reference's assembled rows) -> fp32 [N, dim]."""
    _need_cuda(x, packed)
    feat, hid, dim = shape
    x = _f32c(x)
    if x.dim() != 2 or x.shape[1] != 2 * feat + 2:
        raise RuntimeError(f"mat1 and mat2 shapes cannot be multiplied ({tuple(x.shape)} and ({2 * feat + 2}, {hid}))")
    out = torch.empty((x.shape[0], dim), dtype=torch.float32, device=x.device)
    lib = _lib.load()
    for r0 in range(0, x.shape[0], rows_per_call):
        n = min(rows_per_call, x.shape[0] - r0)
        ws = torch.empty(lib.vfr_visual_embed_tc_bytes(n, 0, feat, hid, dim, 0), dtype=torch.uint8, device=x.device)
        _lib.call("vfr_visual_embed_tc", _ptr(x[r0:r0 + n]), n, feat, _ptr(packed), hid, dim, _ptr(ws), _ptr(out[r0:r0 + n]),
                  _stream())
    return out


def visual_embed_split(seg, ctx, vid_off, packed, shape, videos_per_call=8192):
    """K2 on tensor cores, split-weight form: seg fp32 [C, F] (clips of all videos back to back), ctx fp32 [V, F],
    vid_off [V+1] CSR clip offsets -> fp32 [C, dim].  No [C, 2F+2] concat, context product once per video."""
    _need_cuda(seg, ctx, packed)
    feat, hid, dim = shape
    seg, ctx = _f32c(seg), _f32c(ctx)
    vo = np.asarray(vid_off, dtype=np.int64)
    V = len(vo) - 1
    if seg.shape[1] != feat or ctx.shape != (V, feat) or vo[0] != 0 or vo[-1] != seg.shape[0] or np.any(np.diff(vo) < 1):
        raise ValueError("visual_embed_split: seg [C, F], ctx [V, F] and CSR vid_off [V+1] do not fit together")
    out = torch.empty((seg.shape[0], dim), dtype=torch.float32, device=seg.device)
    lib = _lib.load()
    for v0 in range(0, V, videos_per_call):
        v1 = min(V, v0 + videos_per_call)
        c0, c1 = int(vo[v0]), int(vo[v1])
        off = torch.from_numpy((vo[v0:v1 + 1] - c0).astype(np.int32)).to(seg.device)
        ws = torch.empty(lib.vfr_visual_embed_tc_bytes(c1 - c0, v1 - v0, feat, hid, dim, 1), dtype=torch.uint8, device=seg.device)
        _lib.call("vfr_visual_embed_split", _ptr(seg[c0:c1]), _ptr(ctx[v0:v1]), _ptr(off), c1 - c0, v1 - v0, feat, _ptr(packed),
                  hid, dim, _ptr(ws), _ptr(out[c0:c1]), _stream())
    return out


def lstm_pack(w_ih, w_hh, b_ih, b_hh):
    """Re-layout of one LSTM direction for vfr_text_embed (once per weight update)."""
    _need_cuda(w_ih, w_hh, b_ih, b_hh)
    H, E = w_hh.shape[1], w_ih.shape[1]
    nbytes = _lib.load().vfr_lstm_pack_bytes(H, E)
    packed = torch.empty(nbytes // 4, dtype=torch.float32, device=w_ih.device)
    w_ih, w_hh, b_ih, b_hh = (_f32c(t) for t in (w_ih, w_hh, b_ih, b_hh))    # converted copies stay alive until the launch
    _lib.call("vfr_lstm_pack", _ptr(w_ih), _ptr(w_hh), _ptr(b_ih), _ptr(b_hh), H, E,
              _ptr(packed), _stream())
    return packed


def text_embed(tokens, table, length_table, packed_fwd, packed_bwd, hidden, fc_w, fc_b, check_tokens=True,
               max_batch=8192):
    """K3: tokens int64 [B, L] -> fp32 [B, D] (reference model/models.py:61-66)."""
    _need_cuda(tokens, table, packed_fwd, packed_bwd, fc_w, fc_b)
    tokens = tokens.to(torch.int64).contiguous()
    B, L = tokens.shape
    table = _f32c(table)
    E = table.shape[1]
    lt = None if length_table is None else _f32c(length_table).reshape(-1)
    fc_w, fc_b = _f32c(fc_w), _f32c(fc_b)
    D = fc_w.shape[0]
    out = torch.empty((B, D), dtype=torch.float32, device=tokens.device)
    for b0 in range(0, B, max_batch):
        nb = min(max_batch, B - b0)
        nbytes = _lib.load().vfr_text_embed_bytes(nb, L, hidden, E)
        ws = torch.empty(nbytes // 4, dtype=torch.float32, device=tokens.device)
        _lib.call("vfr_text_embed", _ptr(tokens[b0:b0 + nb]), nb, L, _ptr(table), table.shape[0], _ptr(lt), E,
                  _ptr(packed_fwd), _ptr(packed_bwd), hidden, _ptr(fc_w), _ptr(fc_b), D, _ptr(ws),
                  _ptr(out[b0:b0 + nb]), _stream())
        if check_tokens and int(ws[:4].view(torch.int32)[0].item()) != 0:
            raise IndexError("index out of range in self")   # what nn.Embedding raises in the reference
    return out


def pack_weight_tc(weight):
    _need_cuda(weight)
    w = _f32c(weight)
    packed = torch.empty(_lib.load().vfr_tc_weight_bytes(w.shape[0], w.shape[1]), dtype=torch.uint8, device=w.device)
    _lib.call("vfr_tc_weight_pack", _ptr(w), w.shape[0], w.shape[1], _ptr(packed), _stream())
    return packed


def linear_tc(x, w_packed, out_dim, bias=None, relu=False):
    """Tensor-core (split-bf16, fp32-accurate) ``x @ W.T + b``; ``w_packed`` from ``pack_weight_tc``."""
    _need_cuda(x, w_packed, bias)
    x = _f32c(x)
    b = None if bias is None else _f32c(bias)
    n, k = x.shape
    out = torch.empty((n, out_dim), dtype=torch.float32, device=x.device)
    if n:
        ws = torch.empty(_lib.load().vfr_linear_tc_bytes(n, k), dtype=torch.uint8, device=x.device)
        _lib.call("vfr_linear_tc", _ptr(x), n, k, k, _ptr(w_packed), _ptr(b), out_dim, int(bool(relu)), _ptr(out), out_dim,
                  _ptr(ws), _stream())
    return out


def text_pack_tc(params_fwd, params_bwd, fc_w, fc_b):
    """Pack (w_ih, w_hh, b_ih, b_hh) of both directions + lang_fc for the tensor-core K3."""
    ts = [_f32c(t) for t in (*params_fwd, *params_bwd, fc_w, fc_b)]
    _need_cuda(*ts)
    H, E, D = ts[1].shape[1], ts[0].shape[1], ts[8].shape[0]
    packed = torch.empty(_lib.load().vfr_text_pack_tc_bytes(H, E, D), dtype=torch.uint8, device=ts[0].device)
    _lib.call("vfr_text_pack_tc", *[_ptr(t) for t in ts], H, E, D, _ptr(packed), _stream())
    return packed


def text_embed_tc(tokens, table, length_table, packed, hidden, dim, check_tokens=True, max_batch=32768):
    """K3 on tensor cores: tokens int64 [B, L] -> fp32 [B, D]."""
    _need_cuda(tokens, table, packed)
    tokens = tokens.to(torch.int64).contiguous()
    B, L = tokens.shape
    table = _f32c(table)
    E = table.shape[1]
    lt = None if length_table is None else _f32c(length_table).reshape(-1)
    out = torch.empty((B, dim), dtype=torch.float32, device=tokens.device)
    for b0 in range(0, B, max_batch):
        nb = min(max_batch, B - b0)
        ws = torch.empty(_lib.load().vfr_text_embed_tc_bytes(nb, L, hidden, E), dtype=torch.uint8, device=tokens.device)
        _lib.call("vfr_text_embed_tc", _ptr(tokens[b0:b0 + nb]), nb, L, _ptr(table), table.shape[0], _ptr(lt), E,
                  _ptr(packed), hidden, dim, _ptr(ws), _ptr(out[b0:b0 + nb]), _stream())
        if check_tokens and int(ws[:4].view(torch.int32)[0].item()) != 0:
            raise IndexError("index out of range in self")
    return out


# ---------------------------------------------------------------------------------------------
# K1 : frame -> segment pooling
# ---------------------------------------------------------------------------------------------
POOL_MODES = {"avg": 0, "max": 1, "h5": 2}


def segment_pool(frames, frame_off, mode="avg", window=25, seg_stride=None):
    """K1: frames fp32 [sum F, dim] + int64 offsets [V+1] -> (seg [V, S, dim], ctx [V, dim], n_seg [V])."""
    _need_cuda(frames)
    frames = _f32c(frames)
    fo = np.asarray(frame_off, dtype=np.int64)
    V = len(fo) - 1
    dim = frames.shape[1]
    counts = np.diff(fo)
    if (counts < 1).any() or fo[-1] != frames.shape[0]:
        raise ValueError("frame_off must be increasing offsets covering all frames")
    if seg_stride is None:
        seg_stride = 6 if mode == "h5" else int(-(-counts.max() // window))
    if seg_stride > MAX_SEGMENTS:
        raise _lib.VfrError(f"videos with more than {MAX_SEGMENTS} segments are not supported")
    dev = frames.device
    fo_d = torch.from_numpy(fo).to(dev)
    seg = torch.empty((V, seg_stride, dim), dtype=torch.float32, device=dev)
    ctx = torch.empty((V, dim), dtype=torch.float32, device=dev)
    n_seg = torch.empty(V, dtype=torch.int32, device=dev)
    ws = torch.empty(_lib.load().vfr_segment_pool_bytes(V, dim, seg_stride) // 4, dtype=torch.float32, device=dev)
    _lib.call("vfr_segment_pool", _ptr(frames), _ptr(fo_d), V, dim, window, POOL_MODES[mode], _ptr(seg), seg_stride,
              _ptr(ctx), _ptr(n_seg), _ptr(ws), _stream())
    return seg, ctx, n_seg


# ---------------------------------------------------------------------------------------------
# K6 : ranking loss
# ---------------------------------------------------------------------------------------------
class _RankingLoss(torch.autograd.Function):
    @staticmethod
    def forward(ctx, posit, intra, inter, lang, maskp, maskn, n_samples, normalize, b, lamb):
        _need_cuda(posit, intra, inter, lang, maskp, maskn)
        posit, intra, inter, lang = (_f32c(t) for t in (posit, intra, inter, lang))
        maskp = maskp.to(torch.int64).contiguous()
        maskn = maskn.to(torch.int64).contiguous()
        if lang.shape[0] < n_samples:
            raise IndexError(f"index {lang.shape[0]} is out of bounds for dimension 0 with size {lang.shape[0]}")
        dim = posit.shape[1]
        nbytes = _lib.load().vfr_ranking_loss_bytes(posit.shape[0], intra.shape[0], inter.shape[0], n_samples)
        ws = torch.empty(nbytes // 4, dtype=torch.float32, device=posit.device)
        loss = torch.empty((), dtype=torch.float32, device=posit.device)
        _lib.call("vfr_ranking_loss_fwd", _ptr(posit), _ptr(intra), _ptr(inter), _ptr(lang), _ptr(maskp), _ptr(maskn),
                  posit.shape[0], intra.shape[0], inter.shape[0], n_samples, dim, int(bool(normalize)), float(b),
                  float(lamb), _ptr(ws), _ptr(loss), _stream())
        ctx.save_for_backward(posit, intra, inter, lang, maskp, maskn, ws)
        ctx.cfg = (n_samples, int(bool(normalize)), float(b), float(lamb), lang.shape[0])
        return loss

    @staticmethod
    def backward(ctx, grad_out):
        posit, intra, inter, lang, maskp, maskn, ws = ctx.saved_tensors
        n_samples, normalize, b, lamb, n_lang = ctx.cfg
        dim = posit.shape[1]
        go = _f32c(grad_out).reshape(1)
        gp, gn, gi = torch.empty_like(posit), torch.empty_like(intra), torch.empty_like(inter)
        gl = torch.zeros_like(lang)
        _lib.call("vfr_ranking_loss_bwd", _ptr(posit), _ptr(intra), _ptr(inter), _ptr(lang), _ptr(maskp), _ptr(maskn),
                  posit.shape[0], intra.shape[0], inter.shape[0], n_samples, dim, normalize, b, lamb, _ptr(ws),
                  _ptr(go), _ptr(gp), _ptr(gn), _ptr(gi), _ptr(gl), _stream())
        return gp, gn, gi, gl, None, None, None, None, None, None


def ranking_loss(posit, intra, inter, lang, maskp, maskn, n_samples, normalize=False, b=0.1, lamb=0.4):
    """K6 (reference model/main.py:214-232): differentiable scalar loss (a sum over samples)."""
    return _RankingLoss.apply(posit, intra, inter, lang, maskp, maskn, int(n_samples), normalize, b, lamb)


# ---------------------------------------------------------------------------------------------
# training step: hand-written forward-with-saved-activations / backward of both embedding branches, fused Adam
# (csrc/vfr_train.cu; reference model/main.py:57-67,358 and model/utils.py:85-92)
# ---------------------------------------------------------------------------------------------
def _ptr_array(tensors):
    return (C.c_void_p * len(tensors))(*[t.data_ptr() for t in tensors])


class _VisualTrain(torch.autograd.Function):
    """e = relu(x W1^T + b1) W2^T + b2 with a hand-written backward (vfr_visual_train_bwd)."""

    @staticmethod
    def forward(ctx, x, w1, b1, w2, b2):
        _need_cuda(x, w1, b1, w2, b2)
        x, w1, b1, w2, b2 = (_f32c(t) for t in (x, w1, b1, w2, b2))
        if x.dim() != 2 or x.shape[1] != w1.shape[1]:
            raise RuntimeError(f"mat1 and mat2 shapes cannot be multiplied ({tuple(x.shape)} and {tuple(w1.t().shape)})")
        n, hid, dim = x.shape[0], w1.shape[0], w2.shape[0]
        hidden = torch.empty((n, hid), dtype=torch.float32, device=x.device)
        out = torch.empty((n, dim), dtype=torch.float32, device=x.device)
        if n:
            scratch = torch.empty(_lib.load().vfr_visual_train_fwd_bytes(n, hid, dim) // 4, dtype=torch.float32, device=x.device)
            _lib.call("vfr_visual_train_fwd", _ptr(x), n, x.shape[1], _ptr(w1), _ptr(b1), hid, _ptr(w2), _ptr(b2), dim,
                      _ptr(scratch), _ptr(hidden), _ptr(out), _stream())
        ctx.save_for_backward(x, hidden, w1, w2)
        ctx.need_dx = x.requires_grad
        return out

    @staticmethod
    def backward(ctx, grad):
        x, hidden, w1, w2 = ctx.saved_tensors
        g = _f32c(grad)
        n, hid, dim = x.shape[0], w1.shape[0], w2.shape[0]
        dev = x.device
        d_w1, d_b1 = torch.empty_like(w1), torch.empty(hid, dtype=torch.float32, device=dev)
        d_w2, d_b2 = torch.empty_like(w2), torch.empty(dim, dtype=torch.float32, device=dev)
        d_x = torch.empty_like(x) if ctx.need_dx else None
        if n == 0:
            return (d_x, d_w1.zero_(), d_b1.zero_(), d_w2.zero_(), d_b2.zero_())
        scratch = torch.empty((n, hid), dtype=torch.float32, device=dev)
        _lib.call("vfr_visual_train_bwd", _ptr(x), n, x.shape[1], _ptr(hidden), hid, _ptr(w1), _ptr(w2), dim, _ptr(g),
                  _ptr(scratch), _ptr(d_w1), _ptr(d_b1), _ptr(d_w2), _ptr(d_b2), _ptr(d_x), _stream())
        return d_x, d_w1, d_b1, d_w2, d_b2


def visual_embed_train(x, w1, b1, w2, b2):
    """Differentiable K2 for the training step (gradients to W1, b1, W2, b2 and, if asked for, x)."""
    return _VisualTrain.apply(x, w1, b1, w2, b2)


class _TextTrain(torch.autograd.Function):
    """GloVe gather -> BiLSTM -> Linear with saved activations and a hand-written BPTT (vfr_text_train_fwd / _bwd)."""

    @staticmethod
    def forward(ctx, tokens, table, length, hidden, *params):
        # params: w_ih, w_hh, b_ih, b_hh (forward), the same four (reverse), fc_w, fc_b
        _need_cuda(tokens, table, *params)
        tokens = tokens.to(torch.int64).contiguous()
        table = _f32c(table)
        ps = [_f32c(p) for p in params]
        lt = None if length is None else _f32c(length).reshape(-1)
        B, T = tokens.shape
        E, D = table.shape[1], ps[8].shape[0]
        lib = _lib.load()
        ws = torch.empty(lib.vfr_text_train_bytes(B, T, hidden, E, D), dtype=torch.uint8, device=tokens.device)
        out = torch.empty((B, D), dtype=torch.float32, device=tokens.device)
        w_ih, w_hh, b_ih, b_hh = ([ps[i], ps[4 + i]] for i in range(4))
        _lib.call("vfr_text_train_fwd", _ptr(tokens), B, T, _ptr(table), table.shape[0], _ptr(lt), E, _ptr_array(w_ih),
                  _ptr_array(w_hh), _ptr_array(b_ih), _ptr_array(b_hh), hidden, _ptr(ps[8]), _ptr(ps[9]), D, _ptr(ws), _ptr(out),
                  _stream())
        if int(ws[:4].view(torch.int32)[0].item()) != 0:
            raise IndexError("index out of range in self")
        ctx.save_for_backward(tokens, ws, *ps)
        ctx.cfg = (hidden, E, D, table.shape[0], lt is not None)
        return out

    @staticmethod
    def backward(ctx, grad):
        tokens, ws, *ps = ctx.saved_tensors
        hidden, E, D, vocab, has_length = ctx.cfg
        B, T = tokens.shape
        g = _f32c(grad)
        grads = [torch.empty_like(p) for p in ps]
        d_len = torch.zeros(vocab, dtype=torch.float32, device=tokens.device) if has_length else None
        w_ih, w_hh = [ps[0], ps[4]], [ps[1], ps[5]]
        _lib.call("vfr_text_train_bwd", _ptr(tokens), B, T, vocab, int(has_length), E, _ptr_array(w_ih), _ptr_array(w_hh), hidden,
                  _ptr(ps[8]), D, _ptr(ws), _ptr(g), _ptr_array([grads[0], grads[4]]), _ptr_array([grads[1], grads[5]]),
                  _ptr_array([grads[2], grads[6]]), _ptr_array([grads[3], grads[7]]), _ptr(grads[8]), _ptr(grads[9]),
                  _ptr(d_len), _stream())
        return (None, None, None if d_len is None else d_len.view(-1, 1), None, *grads)


def text_embed_train(tokens, table, length, hidden, params):
    """Differentiable K3 for the training step; ``params`` = the ten LSTM / lang_fc tensors in ``CALModel._text_params``
    order, ``length`` = the learnable word-length table [vocab, 1] or None."""
    return _TextTrain.apply(tokens, table, length, hidden, *params)


def adam_step(params, grads, exp_avg, exp_avg_sq, step, lr=5e-4, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0):
    """torch.optim.Adam's update for a list of fp32 CUDA tensors, 24 tensors per launch (vfr_adam_step)."""
    _need_cuda(*params, *grads, *exp_avg, *exp_avg_sq)
    for i in range(0, len(params), 24):
        sl = slice(i, i + 24)
        numel = (C.c_int64 * len(params[sl]))(*[p.numel() for p in params[sl]])
        _lib.call("vfr_adam_step", _ptr_array(params[sl]), _ptr_array(grads[sl]), _ptr_array(exp_avg[sl]),
                  _ptr_array(exp_avg_sq[sl]), numel, len(params[sl]), int(step), float(lr), float(betas[0]), float(betas[1]),
                  float(eps), float(weight_decay), _stream())


def grad_norms(grads):
    """L2 norm of every tensor of ``grads`` -> fp32 [len(grads)] on the device (one launch per 24 tensors)."""
    _need_cuda(*grads)
    out = torch.empty(len(grads), dtype=torch.float32, device=grads[0].device)
    for i in range(0, len(grads), 24):
        sl = slice(i, i + 24)
        numel = (C.c_int64 * len(grads[sl]))(*[g.numel() for g in grads[sl]])
        _lib.call("vfr_grad_norms", _ptr_array(grads[sl]), numel, len(grads[sl]), _ptr(out[i:]), _stream())
    return out


# ---------------------------------------------------------------------------------------------
# training-batch construction on the device, pooled-moment features (csrc/vfr_sample.cu)
# ---------------------------------------------------------------------------------------------
def sample_negatives(times, q_video, nseg, same_length=True, seed=123, epoch=0):
    """The draws of the reference's ``CustomBatchSampler.__iter__`` (data.py:275-337) for all queries at once.
    ``times`` int32 [Q, A, 2], ``q_video`` int32 [Q], ``nseg`` int32 [V] (device) -> int32 [Q, 8]
    {video_pos, video_neg, start_t, end_t, start_tn, end_tn, status, 0}."""
    _need_cuda(times, q_video, nseg)
    times = times.to(torch.int32).contiguous()
    q_video = q_video.to(torch.int32).contiguous()          # (held in locals until the launch: see moment_pool)
    nseg = nseg.to(torch.int32).contiguous()
    out = torch.empty((times.shape[0], 8), dtype=torch.int32, device=times.device)
    _lib.call("vfr_sample_negatives", _ptr(times), times.shape[1], _ptr(q_video), _ptr(nseg), times.shape[0], nseg.shape[0],
              int(bool(same_length)), int(seed), int(epoch), _ptr(out), _stream())
    return out


def gather_clip_rows(seg, ctx, vid_off, row_video, row_clip):
    """``[segment | context | tef]`` rows (data.py:204-213) of the (video, clip) pairs -> fp32 [R, 2F+2]."""
    _need_cuda(seg, ctx, vid_off, row_video, row_clip)
    feat = seg.shape[1]
    out = torch.empty((row_video.shape[0], 2 * feat + 2), dtype=torch.float32, device=seg.device)
    _lib.call("vfr_gather_clip_rows", _ptr(seg), _ptr(ctx), _ptr(vid_off), _ptr(row_video), _ptr(row_clip), row_video.shape[0], feat,
              _ptr(out), _stream())
    return out


def moment_pool(seg, vid_off):
    """MCN-style pooled-moment features (a NON-reference scoring variant): the mean segment feature of every candidate
    moment of every video -> (fp32 [M_total, F], mom_off int64 [V+1]); row ``mom_off[v] + moment_index(n, s, e)``."""
    _need_cuda(seg)
    seg = _f32c(seg)
    vo = np.asarray(vid_off, dtype=np.int64)
    nseg = np.diff(vo)
    if nseg.max() > MAX_SEGMENTS or seg.shape[1] % 4:
        raise _lib.VfrError("moment_pool: at most 32 segments per video, feature dimension a multiple of 4")
    mo = np.concatenate([[0], np.cumsum(nseg * (nseg + 1) // 2)]).astype(np.int64)
    out = torch.empty((int(mo[-1]), seg.shape[1]), dtype=torch.float32, device=seg.device)
    vo_d = torch.from_numpy(vo.astype(np.int32)).to(seg.device)        # (named: a temporary would be freed - and its block
    mo_d = torch.from_numpy(mo).to(seg.device)                         #  reused by the next one - before the kernel runs)
    _lib.call("vfr_moment_pool", _ptr(seg), _ptr(vo_d), _ptr(mo_d), len(nseg), int(nseg.max()), seg.shape[1], _ptr(out), _stream())
    return out, mo
