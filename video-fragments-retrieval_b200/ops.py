"""Thin torch-tensor wrappers over the libvfr C ABI: argument checking, workspace allocation
(torch owns device memory and streams - plumbing only) and pointer extraction.  Every function
requires CUDA tensors and raises otherwise: there is no CPU path in the product."""
import numpy as np
import torch

from . import _lib
from .utils import MAX_SEGMENTS, threshold_table


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
