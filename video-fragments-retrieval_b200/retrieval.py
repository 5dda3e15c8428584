"""Corpus-scale moment retrieval (BASELINE config 5): tokenised queries -> global top-k moments.

The reference has no serving entry point - its corpus protocol is the python loop of
``model/evaluate.py:42-80`` (every query scored against every moment of every video, then sorted).
``MomentRetriever`` is that loop as a service: the bank of clip embeddings stays resident in HBM
(as the reference keeps its ``videos`` dict), each ``search`` call embeds a batch of queries (K3),
scores it against the bank with the fused top-k (K4) and returns the k best (score, moment id).

Multi-GPU: the bank is partitioned by contiguous video ranges across the ranks of
one NVSwitch box, queries are replicated, every rank computes its local top-k with GLOBAL moment ids
and the per-rank lists are exchanged with ONE ``all_gather`` over NCCL/NVLink and merged (K7).
Moment id = ``mom_off[video] + moment_index`` over the WHOLE corpus, so results do not depend on the
number of ranks.
"""
import ctypes as C

import numpy as np
import torch
import torch.distributed as dist

from . import _lib, ops


def shard_range(n_videos, rank, world):
    """Contiguous video range [v0, v1) of ``rank``."""
    return (n_videos * rank) // world, (n_videos * (rank + 1)) // world


class _DistComm:
    """The collectives of the sharded search over torch.distributed (NCCL)."""

    def __init__(self, group=None):
        self.group = group

    def all_gather(self, t):
        world = dist.get_world_size(self.group)
        out = torch.empty((world * t.shape[0],) + tuple(t.shape[1:]), dtype=t.dtype, device=t.device)   # concatenated form
        dist.all_gather_into_tensor(out, t.contiguous(), group=self.group)
        return out.view((world,) + tuple(t.shape))

    def all_reduce_sum(self, t):
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=self.group)
        return t

    def all_reduce_min(self, t):
        dist.all_reduce(t, op=dist.ReduceOp.MIN, group=self.group)
        return t


class MomentRetriever:

    ENGINES = {"exact": 0, "tc": 3, "tc_bf16": 1, "sel": 4}

    def __init__(self, model, clips, vid_off, id_base=0, max_queries=4096, k=100, n_split=0, group=None,
                 engine="auto", text_engine="auto"):
        """``model``: a ``CALModel`` (text branch used); ``clips`` fp32 [C_local, D] + ``vid_off`` =
        this rank's bank shard; ``id_base`` = global moment id of the shard's first moment.
        ``engine``: "exact" (fp32 CUDA-core scoring, bit-identical to the evaluation path), "tc"
        (tcgen05 split-bf16 scoring, fp32 scores within 1e-5), "tc_bf16" (plain bf16, 1e-2), or
        "sel" (filter + refine: one fp16 tcgen05 pass with a rigorous error band + exact fp32 re-scoring of the
        survivors; results bit-identical to "exact", any clip count per video), or "auto" = "sel" whenever
        D <= 125."""
        self.model = model
        self.bank = ops.Bank(clips, vid_off)
        self.k = int(k)
        self.max_queries = int(max_queries)
        self.group = group
        self.world = dist.get_world_size(group) if (dist.is_available() and dist.is_initialized()) else 1
        dev = self.bank.device
        self.device = dev
        self.seq_len = 20
        lib = _lib.load()
        fwd, bwd = model._packed_lstm()
        table = model.word_embedding.weight.detach().float().contiguous()
        length = model.learnable_length.weight.detach().float().reshape(-1).contiguous() if model.normalize_lang else None
        fc_w = model.lang_fc.weight.detach().float().contiguous()
        fc_b = model.lang_fc.bias.detach().float().contiguous()
        H, E, D = model.hidden_size, table.shape[1], fc_w.shape[0]
        mq = self.max_queries
        self._keep = [fwd, bwd, table, length, fc_w, fc_b]
        self.tokens_dev = torch.empty((mq, self.seq_len), dtype=torch.int64, device=dev)
        self.q_emb = torch.empty((mq, D), dtype=torch.float32, device=dev)
        if text_engine == "auto":
            text_engine = "tc" if H % 4 == 0 else "exact"
        self.text_engine = text_engine
        if text_engine == "tc":
            self.text_tc = model._packed_text_tc()
            self.text_ws = torch.empty(lib.vfr_text_embed_tc_bytes(mq, self.seq_len, H, E) // 4 + 1, dtype=torch.float32, device=dev)
        else:
            self.text_tc = None
            self.text_ws = torch.empty(lib.vfr_text_embed_bytes(mq, self.seq_len, H, E) // 4, dtype=torch.float32, device=dev)
        if engine == "auto":
            engine = "sel" if D <= 125 else "exact"
        self.engine = engine
        eng = self.ENGINES[engine]
        if eng == 0:
            self.q_packed = torch.empty(lib.vfr_query_pack_bytes(mq, D) // 4, dtype=torch.float32, device=dev)
            self.topk_ws = torch.empty(lib.vfr_score_topk_bytes(mq, n_split), dtype=torch.uint8, device=dev)
            self.q_tc = None
        elif eng == 4:
            self.n_clips = int(self.bank.clips.shape[0])
            self.q_packed = None
            self.q_tc = torch.empty(lib.vfr_sel_query_bytes(mq), dtype=torch.uint8, device=dev)
            self.topk_ws = torch.empty(lib.vfr_sel_topk_bytes(mq, self.n_clips, n_split), dtype=torch.uint8, device=dev)
        else:
            self.q_packed = None
            self.q_tc = torch.empty(lib.vfr_tc_query_bytes(mq), dtype=torch.uint8, device=dev)
            self.topk_ws = torch.empty(lib.vfr_score_topk_tc_bytes(mq, self.bank.n_videos, n_split), dtype=torch.uint8,
                                       device=dev)
        self.out_s = torch.empty((mq, self.k), dtype=torch.float32, device=dev)
        self.out_i = torch.empty((mq, self.k), dtype=torch.int64, device=dev)
        p = _lib.SearchPlan()
        p.table, p.vocab, p.length_table, p.emb = table.data_ptr(), table.shape[0], (length.data_ptr() if length is not None else None), E
        p.lstm_fwd, p.lstm_bwd, p.hidden = fwd.data_ptr(), bwd.data_ptr(), H
        p.fc_w, p.fc_b, p.dim, p.seq_len = fc_w.data_ptr(), fc_b.data_ptr(), D, self.seq_len
        p.bank_packed, p.vid_off, p.mom_off = self.bank.packed.data_ptr(), self.bank.vid_off.data_ptr(), self.bank.mom_off.data_ptr()
        p.text_engine = 3 if text_engine == "tc" else 0
        p.text_tc = self.text_tc.data_ptr() if self.text_tc is not None else None
        p.engine = eng
        if eng == 4:
            p.bank_tc, p.bank_clips = self.bank.sel().data_ptr(), self.bank.clips.data_ptr()
            p.q_tc, p.n_clips = self.q_tc.data_ptr(), self.n_clips
        elif eng:
            p.bank_tc, p.bank_clips = self.bank.tc(eng).data_ptr(), self.bank.clips.data_ptr()
            p.uniform6, p.q_tc = self.bank.uniform6, self.q_tc.data_ptr()
        p.n_videos, p.n_max, p.id_base = self.bank.n_videos, self.bank.n_max, int(id_base)
        p.tokens_dev, p.q_emb = self.tokens_dev.data_ptr(), self.q_emb.data_ptr()
        p.q_packed = self.q_packed.data_ptr() if self.q_packed is not None else None
        p.text_ws, p.topk_ws = self.text_ws.data_ptr(), self.topk_ws.data_ptr()
        p.out_scores_dev, p.out_ids_dev = self.out_s.data_ptr(), self.out_i.data_ptr()
        p.n_split, p.max_queries = int(n_split), mq
        self.plan = p
        self.sel_bound = torch.empty(mq, dtype=torch.float32, device=dev)
        self.sel_samp = torch.empty(0, dtype=torch.float32, device=dev)       # grown on first use (sharded search)
        self.sel_count = torch.empty(4 * mq, dtype=torch.int32, device=dev)
        if self.world > 1:
            self.gather_s = torch.empty((self.world, mq, self.k), dtype=torch.float32, device=dev)
            self.gather_i = torch.empty((self.world, mq, self.k), dtype=torch.int64, device=dev)
            per = (mq + self.world - 1) // self.world
            self.q_slice = torch.zeros((per, D), dtype=torch.float32, device=dev)
            self.q_gather = torch.empty((self.world * per, D), dtype=torch.float32, device=dev)
        # pinned staging for the host-buffer path
        self.host_tokens = torch.empty((mq, self.seq_len), dtype=torch.int64).pin_memory()
        self.host_s = torch.empty((mq, self.k), dtype=torch.float32).pin_memory()
        self.host_i = torch.empty((mq, self.k), dtype=torch.int64).pin_memory()
        # kernels launched per search step: gather + 20 LSTM steps + 2 fc + query pack + threshold init + score (filter) + finish (refine) (+ merge)
        if text_engine == "tc":
            # zero + len/scan/perm + gather + (join + step GEMM) x L + final join folded in L + fc
            n_text = 1 + 3 + 1 + 2 * self.seq_len + 1
        else:
            n_text = 1 + self.seq_len + 2
        if self.engine == "sel" and self.world > 1:
            # bound exchange: query pack + threshold init + sample pass + filter (first slice) + bound get / put + filter
            # (rest) + refine + merge; pooled samples (see launches_per_step): query pack + threshold init + sample pass +
            # bound put + filter + 4 counts + bound put + refine + merge
            n_k4 = 9
        elif self.engine == "sel":
            n_k4 = 5          # query pack + threshold init + sample pass + filter + refine
        else:
            n_k4 = 4 + (1 if self.world > 1 else 0)   # query pack + threshold init + score + finish (+ merge)
        self._launches_base = n_text + n_k4

    @property
    def launches_per_step(self):
        """Kernels of this library launched per search step (the claim bench.py reports as gpu_launches)."""
        pooled = any(j > 0 for j in self.__dict__.get("_sel_rank_cache", {}).values())
        return self._launches_base + (3 if pooled else 0)

    def score_only(self, n_queries):
        """K4 alone (query pack excluded) on the query embeddings left in ``q_emb`` by the previous
        search - used by bench.py to time the dominant kernel inside the steps."""
        lib_stream = torch.cuda.current_stream().cuda_stream
        b, p = self.bank, self.plan
        if p.engine == 4:
            _lib.call("vfr_sel_topk", p.bank_tc, p.bank_clips, p.vid_off, p.mom_off, b.n_videos, self.n_clips, b.n_max,
                      b.dim, p.q_tc, p.q_emb, n_queries, self.k, p.id_base, p.out_scores_dev, p.out_ids_dev, p.topk_ws,
                      p.n_split, lib_stream)
        elif p.engine:
            _lib.call("vfr_score_topk_tc", p.bank_tc, p.bank_clips, p.vid_off, p.mom_off, b.n_videos, p.uniform6, b.dim,
                      p.engine, p.q_tc, p.q_emb, n_queries, self.k, p.id_base, p.out_scores_dev, p.out_ids_dev,
                      p.topk_ws, p.n_split, lib_stream)
        else:
            _lib.call("vfr_score_topk", p.bank_packed, p.vid_off, p.mom_off, b.n_videos, b.n_max, b.dim, p.q_packed,
                      n_queries, self.k, p.id_base, p.out_scores_dev, p.out_ids_dev, p.topk_ws, p.n_split, lib_stream)

    def _sel_score_sharded(self, Q, stream, comm=None):
        """K4 of one shard when the bank is spread over several ranks.  A shard's local top-k only has to hold what
        can reach the GLOBAL top-k, so the shards agree on a threshold before the scan (include/vfr.h):

        * large shards pool their SAMPLES: every shard's 32 smallest sampled distances are all-gathered, the j-th
          smallest of the union (j from the pooled sample size: the bank holds k clips under it except with
          probability < 1e-10) becomes every shard's starting bound; after the scan ONE all-reduce(sum) of per-query
          counts verifies the guess (queries that fail are flagged and re-run through the exact engine).  Every shard
          then keeps ~k/P candidates instead of ~k: its hit / list / re-scoring work shrinks with the shard.
        * otherwise the shards exchange a certified bound half way through the scan (one all-reduce(min)).

        ``comm`` replaces the NCCL collectives (tests emulate the ranks on one GPU)."""
        p, b = self.plan, self.bank
        lib = _lib.load()
        comm = comm or _DistComm(self.group)
        n_clips = self.n_clips
        qt, ws = self.q_tc.data_ptr(), self.topk_ws.data_ptr()
        _lib.call("vfr_sel_query_pack", p.q_emb, Q, b.dim, p.bank_tc, n_clips, qt, stream)
        tiles = lib.vfr_sel_tiles(n_clips)
        rank_j = self._sel_global_rank(Q, comm)
        if rank_j > 0:
            lists = lib.vfr_sel_sample_lists(Q, n_clips, p.n_split)
            if self.sel_samp.numel() < Q * lists * 32:
                self.sel_samp = torch.empty(Q * lists * 32, dtype=torch.float32, device=self.q_emb.device)
            samp = self.sel_samp[:Q * lists * 32]
            n_s = C.c_int64(0)
            _lib.call("vfr_sel_sample", p.bank_tc, n_clips, b.dim, qt, Q, self.k, ws, p.n_split, samp.data_ptr(),
                      C.byref(n_s), stream)
            mine = samp.view(Q, lists * 32)
            if lists > 1:
                mine = torch.topk(mine, 32, dim=1, largest=False, sorted=True).values
            pooled = comm.all_gather(mine.contiguous())                                  # [P, Q, 32]
            # candidate bounds: the pooled sample values of rank j (safe a priori), j/2, j/4, j/8 (tighter guesses)
            ranks = sorted({max(1, -(-rank_j // d)) for d in (1, 2, 4, 8)}, reverse=True)
            smallest = torch.topk(pooled.permute(1, 0, 2).reshape(Q, -1), rank_j, dim=1, largest=False, sorted=True).values
            levels = smallest[:, [r - 1 for r in ranks]]
            levels = levels.t().contiguous()                                              # [L, Q], loosest first
            _lib.call("vfr_sel_bound_put", qt, Q, n_clips, b.dim, self.k, ws, p.n_split, levels[0].data_ptr(), stream)
            _lib.call("vfr_sel_filter", p.bank_tc, n_clips, b.dim, qt, Q, self.k, ws, p.n_split, 0, tiles, 1, stream)
            # how many clips are CERTAINLY within each bound, over all shards: the tightest bound that still holds k
            # clips is a certified bound of the global k-th distance - every shard re-scores only what is under it
            counts = self.sel_count[:len(ranks) * Q].view(len(ranks), Q)
            for i in range(len(ranks)):
                _lib.call("vfr_sel_count_under", qt, Q, n_clips, b.dim, self.k, ws, p.n_split, levels[i].data_ptr(),
                          counts[i].data_ptr(), stream)
            counts = comm.all_reduce_sum(counts)
            ok = counts >= self.k                                                         # [L, Q], monotone in L
            flags = self._sel_flags(Q)
            flags[(~ok[0]) & (flags == 0)] = 4                                            # the sample promised k clips that are not there
            best = ok.to(torch.int32).sum(dim=0).clamp_(min=1) - 1                        # index of the tightest bound that holds
            bound = levels.gather(0, best.view(1, Q).to(torch.int64)).view(Q).contiguous()
            _lib.call("vfr_sel_bound_put", qt, Q, n_clips, b.dim, self.k, ws, p.n_split, bound.data_ptr(), stream)
        else:
            first = min(tiles, getattr(self, "sel_first_tiles", None) or max(32, tiles // 8))
            _lib.call("vfr_sel_filter", p.bank_tc, n_clips, b.dim, qt, Q, self.k, ws, p.n_split, 0, first, 0, stream)
            if first < tiles:
                bound = self.sel_bound[:Q]
                _lib.call("vfr_sel_bound_get", qt, Q, n_clips, b.dim, self.k, ws, p.n_split, bound.data_ptr(), stream)
                bound = comm.all_reduce_min(bound)
                _lib.call("vfr_sel_bound_put", qt, Q, n_clips, b.dim, self.k, ws, p.n_split, bound.data_ptr(), stream)
                _lib.call("vfr_sel_filter", p.bank_tc, n_clips, b.dim, qt, Q, self.k, ws, p.n_split, first, tiles, 1,
                          stream)
        _lib.call("vfr_sel_refine", p.bank_clips, p.vid_off, p.mom_off, b.n_videos, n_clips, b.n_max, b.dim, qt, p.q_emb, Q,
                  self.k, p.id_base, self.out_s.data_ptr(), self.out_i.data_ptr(), ws, p.n_split, stream)

    def _sel_global_rank(self, Q, comm):
        """The rank j of the pooled-sample protocol for batches of Q queries (0: not applicable - some shard is too
        small to sample, or no rank <= 32 is safe).  Depends on the shard sizes only: one tiny all-reduce per distinct
        batch size, cached."""
        cache = self.__dict__.setdefault("_sel_rank_cache", {})
        if Q not in cache:
            lib = _lib.load()
            n_s = 0 if getattr(self, "sel_pool_samples", True) is False else \
                lib.vfr_sel_sample_clips(Q, self.n_clips, self.k, self.plan.n_split)
            t = torch.tensor([n_s, self.n_clips, 1 if n_s > 0 else 0, 1], dtype=torch.int64, device=self.q_emb.device)
            tot_s, tot_c, n_ok, n_ranks = comm.all_reduce_sum(t).tolist()
            cache[Q] = lib.vfr_sel_sample_rank(self.k, tot_s, tot_c) if n_ok == n_ranks else 0
        return cache[Q]

    def _sel_flags(self, Q):
        """int32 [Q] view of the per-query flags of the last filter + refine call (0 = guaranteed exact)."""
        off = _lib.load().vfr_sel_flags(self.q_tc.data_ptr(), Q) - self.q_tc.data_ptr()
        return self.q_tc[off:off + 4 * Q].view(torch.int32)

    def _sel_fixup(self, Q):
        """Queries the filter + refine engine could not certify (operand magnitudes outside the fp16 scales, or
        more candidates inside the error band than its lists hold - mass duplicates) are re-run through the
        exact engine.  One 4-byte read-back per step when nothing is flagged."""
        if self.plan.engine != 4:
            return
        flags = self._sel_flags(Q)
        if int((flags != 0).any().item()) == 0:
            return
        idx = torch.nonzero(flags != 0).reshape(-1)
        s, i = ops.score_topk(self.bank, self.q_emb[:Q][idx].contiguous(), self.k, id_base=int(self.plan.id_base))
        self.out_s[:Q][idx] = s
        self.out_i[:Q][idx] = i
        self.n_fixups = getattr(self, "n_fixups", 0) + int(idx.numel())

    # -- device-resident step ---------------------------------------------------------------------
    def search_device(self, tokens_dev):
        """tokens int64 [Q, 20] on the device -> (scores fp32 [Q, k], ids int64 [Q, k]) on the device.

        One GPU: K3 -> K4 in one C call.  N GPUs: every rank embeds its slice of the batch (K3 is
        data-parallel over queries), ONE all-gather replicates the embeddings, every rank scores the
        whole batch against its bank shard (K4), ONE all-gather exchanges the per-shard top-k lists,
        K7 merges them."""
        Q = tokens_dev.shape[0]
        stream = torch.cuda.current_stream().cuda_stream
        if self.world == 1:
            _lib.call("vfr_search_device", C.byref(self.plan), tokens_dev.data_ptr(), Q, self.k, self.out_s.data_ptr(),
                      self.out_i.data_ptr(), stream)
            self._sel_fixup(Q)
            return self.out_s[:Q], self.out_i[:Q]
        rank = dist.get_rank(self.group)
        per = (Q + self.world - 1) // self.world                      # queries embedded per rank
        q0, q1 = min(rank * per, Q), min((rank + 1) * per, Q)
        mine = self.q_slice[:per]
        if q1 > q0:
            _lib.call("vfr_search_embed_device", C.byref(self.plan), tokens_dev[q0:q1].data_ptr(), q1 - q0,
                      mine.data_ptr(), stream)
        gathered = self.q_gather[:self.world * per]
        dist.all_gather_into_tensor(gathered, mine, group=self.group)
        self.q_emb[:Q].copy_(gathered[:Q])
        if self.plan.engine == 4:
            self._sel_score_sharded(Q, stream)
        else:
            _lib.call("vfr_search_score_device", C.byref(self.plan), Q, self.k, self.out_s.data_ptr(),
                      self.out_i.data_ptr(), stream)
        self._sel_fixup(Q)
        gs, gi = self.gather_s[:, :Q].contiguous(), self.gather_i[:, :Q].contiguous()
        dist.all_gather_into_tensor(gs, self.out_s[:Q].contiguous(), group=self.group)
        dist.all_gather_into_tensor(gi, self.out_i[:Q].contiguous(), group=self.group)
        return ops.topk_merge(gs, gi)

    # -- host-buffer step (the call a user makes) -------------------------------------------------------
    def search(self, tokens):
        """tokens: int64 [Q, 20] HOST tensor/array -> (scores, ids) HOST tensors [Q, k]."""
        tokens = torch.as_tensor(tokens, dtype=torch.int64)
        Q = tokens.shape[0]
        if Q > self.max_queries:
            raise ValueError(f"batch of {Q} queries exceeds max_queries={self.max_queries}")
        self.host_tokens[:Q].copy_(tokens)
        stream = torch.cuda.current_stream().cuda_stream
        if self.world == 1:
            _lib.call("vfr_search_host", C.byref(self.plan), self.host_tokens.data_ptr(), Q, self.k,
                      self.host_s.data_ptr(), self.host_i.data_ptr(), stream)
            if self.plan.engine == 4 and int((self._sel_flags(Q) != 0).any().item()):
                self._sel_fixup(Q)
                self.host_s[:Q].copy_(self.out_s[:Q])
                self.host_i[:Q].copy_(self.out_i[:Q])
                torch.cuda.current_stream().synchronize()
        else:
            self.tokens_dev[:Q].copy_(self.host_tokens[:Q], non_blocking=True)
            s, i = self.search_device(self.tokens_dev[:Q])
            self.host_s[:Q].copy_(s, non_blocking=True)
            self.host_i[:Q].copy_(i, non_blocking=True)
            torch.cuda.current_stream().synchronize()
        bad = int(self.text_ws[:4].view(torch.int32)[0].item())
        if bad:
            raise IndexError("index out of range in self")
        return self.host_s[:Q], self.host_i[:Q]

    def h2d_bytes(self, Q):
        return Q * self.seq_len * 8

    def d2h_bytes(self, Q):
        return Q * self.k * (4 + 8)


# ---------------------------------------------------------------------------------------------------
# exact rank of the first positive over a SHARDED bank: the metrics of
# model/evaluate.py:67-80 need ranks far beyond any top-k, so the owner shard of a query's video computes
# tau = its smallest positive score, one all-reduce(min) hands tau to every shard, every shard counts
# score < tau (and the deterministic tie terms) over its own videos, one all-reduce(sum) adds the counts.
# ---------------------------------------------------------------------------------------------------
def _dist_reduce(group):
    def reduce(t, op):
        dist.all_reduce(t, op=dist.ReduceOp.MIN if op == "min" else dist.ReduceOp.SUM, group=group)
        return t
    return reduce


def sharded_rank_first_positive(shard, v0, q_emb, q_video, q_nseg, times_list, iou_thresholds, inclusive=False,
                                reduce=None, group=None):
    """``shard``: ``ops.Bank`` of the contiguous video range starting at global video ``v0``; ``q_video`` int [Q]
    GLOBAL index of every query's own video, ``q_nseg`` int [Q] its clip count, ``times_list`` the annotator
    times.  ``reduce(tensor, "min" | "sum")`` defaults to ``torch.distributed.all_reduce`` over ``group``.
    Returns int64 [Q, T] on the device: the 0-based rank of the first positive among ALL moments of ALL shards -
    identical to ``evaluate.rank_first_positive(...)['rank']`` on the unsharded bank."""
    if reduce is None:
        reduce = _dist_reduce(group)
    dev = shard.device
    q_video = np.asarray(q_video, dtype=np.int64)
    local = q_video - int(v0)                                    # index inside this shard (may fall outside)
    mine = (local >= 0) & (local < shard.n_videos)
    Q, T = len(q_video), len(iou_thresholds)
    tau = torch.full((Q, T), float("inf"), dtype=torch.float32, device=dev)
    own_eqb = torch.zeros((Q, T), dtype=torch.int64, device=dev)
    npos = torch.zeros((Q, T), dtype=torch.int64, device=dev)
    if mine.any():
        idx = torch.as_tensor(np.nonzero(mine)[0], device=dev)
        own = ops.score_own(shard, q_emb[idx].contiguous(), torch.as_tensor(local[mine].astype(np.int32), device=dev))
        times = ops.pack_times([times_list[i] for i in np.nonzero(mine)[0]], dev)
        tables = ops.threshold_tables(list(iou_thresholds), inclusive, dev)
        nseg = torch.as_tensor(np.asarray(q_nseg)[mine].astype(np.int32), device=dev)
        _, t_own, _, n_own, e_own = ops.gt_select(own, nseg, times, tables)
        tau[idx] = t_own
        own_eqb[idx] = e_own.to(torch.int64)
        npos[idx] = n_own.to(torch.int64)
    tau = reduce(tau, "min")
    own_eqb = reduce(own_eqb, "sum")
    npos = reduce(npos, "sum")
    # tie-break term counts videos BEFORE the query's own one: the kernel compares shard-local indices
    qv_local = torch.as_tensor(np.clip(local, -1, 2 ** 31 - 2).astype(np.int32), device=dev)
    lt, eqb = ops.score_count(shard, q_emb, tau, qv_local)
    lt = reduce(lt.contiguous(), "sum")
    eqb = reduce(eqb.contiguous(), "sum")
    return lt + eqb + own_eqb, npos
