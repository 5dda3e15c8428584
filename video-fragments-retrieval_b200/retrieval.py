"""Corpus-scale moment retrieval (BASELINE config 5): tokenised queries -> global top-k moments.

The reference has no serving entry point - its corpus protocol is the python loop of
``model/evaluate.py:42-80`` (every query scored against every moment of every video, then sorted).
``MomentRetriever`` is that loop as a service: the bank of clip embeddings stays resident in HBM
(as the reference keeps its ``videos`` dict), each ``search`` call embeds a batch of queries (K3),
scores it against the bank with the fused top-k (K4) and returns the k best (score, moment id).

Multi-GPU: the bank is partitioned by contiguous video ranges across the ranks of
one NVSwitch box; moment id = ``mom_off[video] + moment_index`` over the WHOLE corpus, so results do not
depend on the number of ranks.  One step on P ranks (``search_device``):

1. every rank embeds ITS SLICE of the query batch (K3 is data-parallel over queries), one all-gather
   replicates the embeddings;
2. every rank scores the whole batch against its bank shard (K4); with the filter + refine engine the
   shards first agree on a threshold (pooled samples, ``_sel_score_sharded``) so that a shard keeps ~k/P
   candidates;
3. the refine stage writes its lists straight into QUERY-SLICE records (ids | scores | flags) and ONE
   all-to-all hands rank r the P shard lists of slice r - 1/P of the bytes of an all-gather of ``[P, Q, k]``,
   and every rank merges only its own Q/P queries (K7);
4. rank r OWNS the results of slice r: ``search`` copies only that slice back to the host (and copied only
   that slice of the token ids to the device).  ``owned_range(Q)`` gives the rows.

The per-query flags of the filter + refine engine travel inside the records; a tiny all-gather makes every rank
see all of them (plus the out-of-range-token flag of every rank's K3), so the rare exact-engine rerun and the
``IndexError`` of a bad token are taken by all ranks together - no collective is ever guarded by a rank-local
condition, and the only host synchronisation of a step is one event wait at its very end.
"""
import contextlib
import ctypes as C

import time

import numpy as np
import torch
import torch.distributed as dist

from . import _lib, ops


def _mark(marks, name):
    if marks and name in marks:
        marks[name].record()


def shard_range(n_videos, rank, world):
    """Contiguous video range [v0, v1) of ``rank``."""
    return (n_videos * rank) // world, (n_videos * (rank + 1)) // world


def slice_rows(n_queries, world):
    """Queries per rank of the query-slice exchange: ceil(Q / P), rounded up to an even count (the records hold
    int64 ids behind fp32 scores and int32 flags: ``per * (3k + 1)`` must be even for 8-byte alignment)."""
    per = (int(n_queries) + world - 1) // world
    return per + (per & 1)


class _DistComm:
    """The collectives of the sharded search over torch.distributed (NCCL; gloo in the CPU tests)."""

    def __init__(self, group=None):
        self.group = group

    def all_gather(self, t, out=None):
        world = dist.get_world_size(self.group)
        if out is None:
            out = torch.empty((world * t.shape[0],) + tuple(t.shape[1:]), dtype=t.dtype, device=t.device)   # concatenated form
        dist.all_gather_into_tensor(out, t.contiguous(), group=self.group)
        return out.view((world,) + tuple(t.shape))

    def all_reduce_sum(self, t):
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=self.group)
        return t

    def all_reduce_min(self, t):
        dist.all_reduce(t, op=dist.ReduceOp.MIN, group=self.group)
        return t

    def all_to_all(self, send, out=None):
        """send [P, n] (row j goes to rank j) -> [P, n] (row j came from rank j)."""
        if out is None:
            out = torch.empty_like(send)
        dist.all_to_all_single(out.view(-1), send.contiguous().view(-1), group=self.group)
        return out


class MomentRetriever:

    ENGINES = {"exact": 0, "tc": 3, "tc_bf16": 1, "sel": 4}

    def __init__(self, model, clips, vid_off, id_base=0, max_queries=4096, k=100, n_split=0, group=None,
                 engine="auto", text_engine="auto", comm=None, world=None, rank=None):
        """``model``: a ``CALModel`` (text branch used); ``clips`` fp32 [C_local, D] + ``vid_off`` =
        this rank's bank shard; ``id_base`` = global moment id of the shard's first moment.
        ``engine``: "exact" (fp32 CUDA-core scoring, bit-identical to the evaluation path), "tc"
        (tcgen05 split-bf16 scoring, fp32 scores within 1e-5), "tc_bf16" (plain bf16, 1e-2), or
        "sel" (filter + refine: one fp16 tcgen05 pass with a rigorous error band + exact fp32 re-scoring of the
        survivors; results bit-identical to "exact", any clip count per video), or "auto" = "sel" whenever
        the embedding dimension fits the engine.
        ``comm`` / ``world`` / ``rank`` replace torch.distributed (the tests emulate the ranks on one GPU)."""
        self.model = model
        self.bank = ops.Bank(clips, vid_off)
        self.k = int(k)
        self.max_queries = int(max_queries)
        self.group = group
        if world is None:
            world = dist.get_world_size(group) if (dist.is_available() and dist.is_initialized()) else 1
            rank = dist.get_rank(group) if world > 1 else 0
        self.world, self.rank = int(world), int(rank or 0)
        self.comm = comm if comm is not None else (_DistComm(group) if self.world > 1 else None)
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
        per = slice_rows(mq, self.world) if self.world > 1 else mq
        self._keep = [fwd, bwd, table, length, fc_w, fc_b]
        self.tokens_dev = torch.empty((mq, self.seq_len), dtype=torch.int64, device=dev)
        # the all-gather of the query slices lands straight in q_emb: room for world * per rows
        self.q_emb = torch.zeros((max(mq, per * self.world), D), dtype=torch.float32, device=dev)
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
            engine = "sel" if D <= ops.SEL_MAX_DIM else "exact"
        self.engine = engine
        eng = self.ENGINES[engine]
        if eng == 0:
            self.q_packed = torch.empty(lib.vfr_query_pack_bytes(mq, D) // 4, dtype=torch.float32, device=dev)
            self.topk_ws = torch.empty(lib.vfr_score_topk_bytes(mq, n_split), dtype=torch.uint8, device=dev)
            self.q_tc = None
        elif eng == 4:
            self.n_clips = int(self.bank.clips.shape[0])
            self.q_packed = None
            self.q_tc = torch.empty(lib.vfr_sel_query_bytes(mq, D), dtype=torch.uint8, device=dev)
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
        self.sel_levels = torch.empty(4 * mq, dtype=torch.float32, device=dev)
        self.sel_count = torch.empty(4 * mq, dtype=torch.int32, device=dev)
        self.sel_stats = torch.zeros(8, dtype=torch.int64, device=dev)
        self.n_fixups = 0
        self.profile = False               # True: every stage of a sharded step is bracketed by CUDA events (stage_ms)
        self.trace_search = None           # a dict: ``search`` adds the host-side seconds of its phases to it (P > 1)
        self._prof = {}
        if self.world > 1:
            blk = lib.vfr_topk_block_bytes(per, self.k)
            self.send_blocks = torch.empty((self.world, blk), dtype=torch.uint8, device=dev)
            self.recv_blocks = torch.empty((self.world, blk), dtype=torch.uint8, device=dev)
            self.slice_s = torch.empty((per, self.k), dtype=torch.float32, device=dev)
            self.slice_i = torch.empty((per, self.k), dtype=torch.int64, device=dev)
            # per-query flags of my slice + one trailing word: this rank's out-of-range-token flag (K3)
            self.slice_flags = torch.zeros(per + 1, dtype=torch.int32, device=dev)
            self.flags_all = torch.zeros(self.world * (per + 1), dtype=torch.int32, device=dev)
            self.host_flags = torch.zeros(self.world * (per + 1), dtype=torch.int32).pin_memory()
            self.flags_event = torch.cuda.Event()
            self.q_slice = torch.zeros((per, D), dtype=torch.float32, device=dev)
            self.gather_s = self.gather_i = None           # (non-sel engines: allocated on first use)
        # pinned staging for the host-buffer path (a rank only ever moves its own slice)
        self.host_tokens = torch.empty((per, self.seq_len), dtype=torch.int64).pin_memory()
        self.host_s = torch.empty((per, self.k), dtype=torch.float32).pin_memory()
        self.host_i = torch.empty((per, self.k), dtype=torch.int64).pin_memory()

    # -- bookkeeping -------------------------------------------------------------------------------------
    def owned_range(self, n_queries):
        """Rows [q0, q1) of a batch of ``n_queries`` whose results this rank returns."""
        if self.world == 1:
            return 0, int(n_queries)
        per = slice_rows(n_queries, self.world)
        return min(self.rank * per, int(n_queries)), min((self.rank + 1) * per, int(n_queries))

    @contextlib.contextmanager
    def _stage(self, name):
        if not self.profile:
            yield
            return
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        yield
        b.record()
        self._prof.setdefault(name, []).append((a, b))

    def stage_ms(self, reset=True):
        """Mean milliseconds per stage of the profiled steps (``profile = True``), in first-seen order."""
        torch.cuda.synchronize()
        out = {k: float(np.mean([a.elapsed_time(b) for a, b in v])) for k, v in self._prof.items()}
        if reset:
            self._prof = {}
        return out

    def filter_stats(self, n_queries):
        """Diagnostics of the last filter + refine call (engine "sel"): candidates kept per query, compactions, flags."""
        if self.plan.engine != 4:
            return None
        p, b = self.plan, self.bank
        _lib.call("vfr_sel_stats", self.q_tc.data_ptr(), n_queries, self.n_clips, b.dim, self.k, self.topk_ws.data_ptr(),
                  p.n_split, self.sel_stats.data_ptr(), torch.cuda.current_stream().cuda_stream)
        tot, longest, flagged, lists, events, compacted = self.sel_stats[:6].tolist()
        return dict(candidates_per_query=tot / n_queries, longest_list=longest, n_flagged=flagged, lists=lists,
                    compactions_per_list=compacted / max(lists, 1), compaction_warp_events=events)

    # -- the stages of a step -----------------------------------------------------------------------------
    def embed_only(self, tokens_dev):
        """K3 (+ the all-gather of the slices on P ranks): query embeddings of the batch -> ``q_emb[:Q]``."""
        Q = tokens_dev.shape[0]
        stream = torch.cuda.current_stream().cuda_stream
        if self.world == 1:
            _lib.call("vfr_search_embed_device", C.byref(self.plan), tokens_dev.data_ptr(), Q, self.q_emb.data_ptr(), stream)
            return
        per = slice_rows(Q, self.world)
        q0, q1 = self.owned_range(Q)
        mine = self.q_slice[:per]
        with self._stage("k3_embed"):
            if q1 > q0:
                _lib.call("vfr_search_embed_device", C.byref(self.plan), tokens_dev[q0:q1].data_ptr(), q1 - q0,
                          mine.data_ptr(), stream)
        with self._stage("allgather_queries"):
            self.comm.all_gather(mine, out=self.q_emb[:self.world * per])

    def score_only(self, n_queries):
        """K4 alone on the query embeddings in ``q_emb`` (the dominant kernel; bench.py times it inside the steps).
        On P ranks this is the shard's whole K4 including the threshold protocol's collectives."""
        stream = torch.cuda.current_stream().cuda_stream
        b, p = self.bank, self.plan
        if self.world > 1 and p.engine == 4:
            self._sel_score_sharded(n_queries, stream, per=slice_rows(n_queries, self.world))
        elif p.engine == 4:
            _lib.call("vfr_sel_query_pack", p.q_emb, n_queries, b.dim, p.bank_tc, self.n_clips, p.q_tc, stream)
            _lib.call("vfr_sel_topk", p.bank_tc, p.bank_clips, p.vid_off, p.mom_off, b.n_videos, self.n_clips, b.n_max,
                      b.dim, p.q_tc, p.q_emb, n_queries, self.k, p.id_base, p.out_scores_dev, p.out_ids_dev, p.topk_ws,
                      p.n_split, stream)
        else:
            _lib.call("vfr_search_score_device", C.byref(p), n_queries, self.k, self.out_s.data_ptr(), self.out_i.data_ptr(),
                      stream)

    def _sel_score_sharded(self, Q, stream, comm=None, per=0):
        """K4 of one shard when the bank is spread over several ranks.  A shard's local top-k only has to hold what
        can reach the GLOBAL top-k, so the shards agree on a threshold before the scan (include/vfr.h):

        * large shards pool their SAMPLES: every shard's smallest sampled distances are all-gathered, the j-th
          smallest of the union (j from the pooled sample size: the bank holds k clips under it except with
          probability < 1e-10) becomes every shard's starting bound; after the scan ONE all-reduce(sum) of per-query
          counts verifies the guess (queries that fail are flagged and re-run through the exact engine) and picks the
          tightest of four candidate bounds that still holds k clips.  Every shard then keeps ~k/P candidates instead
          of ~k: its hit / list / re-scoring work shrinks with the shard.  Kernels only between the collectives
          (``vfr_sel_pool_levels`` / ``count_levels`` / ``pick_put``).
        * otherwise the shards exchange a certified bound part way through the scan (one all-reduce(min)).

        ``per`` > 0: the refine stage writes the query-slice records of the all-to-all (``send_blocks``) instead of
        the flat ``out_s`` / ``out_i``.  ``comm`` replaces the NCCL collectives (tests emulate the ranks on one GPU)."""
        p, b = self.plan, self.bank
        lib = _lib.load()
        comm = comm or self.comm
        n_clips = self.n_clips
        qt, ws = self.q_tc.data_ptr(), self.topk_ws.data_ptr()
        with self._stage("k4_query_pack"):
            _lib.call("vfr_sel_query_pack", p.q_emb, Q, b.dim, p.bank_tc, n_clips, qt, stream)
        tiles = lib.vfr_sel_tiles(n_clips)
        rank_j, n_src = self._sel_global_rank(Q, comm)
        lists = lib.vfr_sel_sample_lists(Q, n_clips, p.n_split, b.dim)
        if rank_j > 0:
            width = lists * 32
            if self.sel_samp.numel() < Q * width:
                self.sel_samp = torch.empty(Q * width, dtype=torch.float32, device=self.q_emb.device)
            samp = self.sel_samp[:Q * width].view(Q, width)
            n_s = C.c_int64(0)
            with self._stage("k4_sample_pass"):
                _lib.call("vfr_sel_sample", p.bank_tc, n_clips, b.dim, qt, Q, self.k, ws, p.n_split, samp.data_ptr(),
                          C.byref(n_s), stream)
            with self._stage("allgather_samples"):
                pooled = comm.all_gather(samp)                                               # [P, Q, width]
            # candidate bounds: the pooled sample values of rank j (safe a priori), j/2, j/4, j/8 (tighter guesses)
            ranks = sorted({max(1, -(-rank_j // d)) for d in (1, 2, 4, 8)}, reverse=True)
            L = len(ranks)
            levels, counts = self.sel_levels[:L * Q], self.sel_count[:L * Q]
            with self._stage("k4_levels"):
                _lib.call("vfr_sel_pool_levels", pooled.data_ptr(), n_src, Q, width, (C.c_int32 * L)(*ranks), L, 1,
                          levels.data_ptr(), stream)
                _lib.call("vfr_sel_bound_put", qt, Q, n_clips, b.dim, self.k, ws, p.n_split, levels.data_ptr(), stream)
            _mark(getattr(self, "_marks", None), "scan_start")
            with self._stage("k4_filter"):
                _lib.call("vfr_sel_filter", p.bank_tc, n_clips, b.dim, qt, Q, self.k, ws, p.n_split, 0, tiles, 1, stream)
            _mark(getattr(self, "_marks", None), "scan_end")
            # how many clips are CERTAINLY within each bound, over all shards: the tightest bound that still holds k
            # clips is a certified bound of the global k-th distance - every shard re-scores only what is under it
            with self._stage("k4_count"):
                _lib.call("vfr_sel_count_levels", qt, Q, n_clips, b.dim, self.k, ws, p.n_split, levels.data_ptr(), L,
                          counts.data_ptr(), stream)
            with self._stage("allreduce_counts"):
                counts = comm.all_reduce_sum(counts)
            with self._stage("k4_pick"):
                _lib.call("vfr_sel_pick_put", qt, Q, n_clips, b.dim, self.k, ws, p.n_split, levels.data_ptr(),
                          counts.data_ptr(), L, stream)
        else:
            first = min(tiles, getattr(self, "sel_first_tiles", None) or max(32, tiles // 8))
            with self._stage("k4_filter"):
                _lib.call("vfr_sel_filter", p.bank_tc, n_clips, b.dim, qt, Q, self.k, ws, p.n_split, 0, first, 0, stream)
            # EVERY rank takes part in the exchange, whatever its own tile count (a collective guarded by a rank-local
            # condition would desynchronise NCCL on uneven shards): a shard that has already scanned all of its tiles
            # contributes the certified bound of its whole bank and simply has nothing left to filter
            bound = self.sel_bound[:Q]
            _lib.call("vfr_sel_bound_get", qt, Q, n_clips, b.dim, self.k, ws, p.n_split, bound.data_ptr(), stream)
            with self._stage("allreduce_bound"):
                bound = comm.all_reduce_min(bound)
            _lib.call("vfr_sel_bound_put", qt, Q, n_clips, b.dim, self.k, ws, p.n_split, bound.data_ptr(), stream)
            if first < tiles:
                with self._stage("k4_filter"):
                    _lib.call("vfr_sel_filter", p.bank_tc, n_clips, b.dim, qt, Q, self.k, ws, p.n_split, first, tiles, 1,
                              stream)
        with self._stage("k4_refine"):
            if per > 0:
                _lib.call("vfr_sel_refine_blocks", p.bank_clips, p.vid_off, p.mom_off, b.n_videos, n_clips, b.n_max, b.dim,
                          qt, p.q_emb, Q, self.k, p.id_base, per, self.send_blocks.data_ptr(), ws, p.n_split, stream)
            else:
                _lib.call("vfr_sel_refine", p.bank_clips, p.vid_off, p.mom_off, b.n_videos, n_clips, b.n_max, b.dim, qt,
                          p.q_emb, Q, self.k, p.id_base, self.out_s.data_ptr(), self.out_i.data_ptr(), ws, p.n_split, stream)

    def _sel_global_rank(self, Q, comm):
        """(j, P): the rank j of the pooled-sample protocol for batches of Q queries (0: not applicable - some shard is
        too small to sample, more pooled values than the level kernel holds, or no rank <= 32 is safe) and the number
        of shards.  Depends on the shard sizes only: one tiny all-reduce per distinct batch size, cached."""
        cache = self.__dict__.setdefault("_sel_rank_cache", {})
        if Q not in cache:
            lib = _lib.load()
            n_s = 0 if getattr(self, "sel_pool_samples", True) is False else \
                lib.vfr_sel_sample_clips(Q, self.n_clips, self.k, self.plan.n_split, self.bank.dim)
            lists = lib.vfr_sel_sample_lists(Q, self.n_clips, self.plan.n_split, self.bank.dim)
            t = torch.tensor([n_s, self.n_clips, 1 if n_s > 0 else 0, 1, lists, lists * lists], dtype=torch.int64,
                             device=self.q_emb.device)
            tot_s, tot_c, n_ok, n_ranks, tot_lists, tot_sq = comm.all_reduce_sum(t).tolist()
            # the same list count on every shard (P sum(l^2) == (sum l)^2 holds only then), decided from reduced values
            # alone: every rank must take the same branch of the protocol, whatever its own shard looks like
            fits = n_ranks * tot_sq == tot_lists * tot_lists and tot_lists * 32 <= 1024
            cache[Q] = (lib.vfr_sel_sample_rank(self.k, tot_s, tot_c) if (n_ok == n_ranks and fits) else 0, int(n_ranks))
        return cache[Q]

    def _sel_flags(self, Q):
        """int32 [Q] view of the per-query flags of the last filter + refine call (0 = guaranteed exact)."""
        off = _lib.load().vfr_sel_flags(self.q_tc.data_ptr(), Q, self.bank.dim) - self.q_tc.data_ptr()
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
        self.n_fixups += int(idx.numel())

    # -- device-resident step ---------------------------------------------------------------------
    def search_device(self, tokens_dev, check=True, marks=None):
        """tokens int64 [Q, 20] on the device -> (scores fp32 [n, k], ids int64 [n, k]) on the device for the rows
        ``owned_range(Q)`` of the batch (one GPU: the whole batch).

        One GPU: K3 -> K4 in one C call.  P GPUs: see the module docstring.  ``check=False`` skips the end-of-step
        wait on the flags (the caller then calls ``finish_step(Q)`` before it trusts the result).  ``marks``: a dict of
        CUDA events to record inside the step - ``k3_done`` (after K3 + the all-gather of the embeddings), ``scan_start`` /
        ``scan_end`` (around the filter kernel's scan; one GPU: sample pass included) - bench.py times the stages inside
        its steps with them."""
        Q = tokens_dev.shape[0]
        stream = torch.cuda.current_stream().cuda_stream
        if self.world == 1:
            if not marks:
                _lib.call("vfr_search_device", C.byref(self.plan), tokens_dev.data_ptr(), Q, self.k, self.out_s.data_ptr(),
                          self.out_i.data_ptr(), stream)
            else:
                # the same stages through their own entry points, with the caller's events between them
                p, b = self.plan, self.bank
                _lib.call("vfr_search_embed_device", C.byref(p), tokens_dev.data_ptr(), Q, self.q_emb.data_ptr(), stream)
                _mark(marks, "k3_done")
                if p.engine == 4:
                    _lib.call("vfr_sel_query_pack", p.q_emb, Q, b.dim, p.bank_tc, self.n_clips, p.q_tc, stream)
                    _mark(marks, "scan_start")
                    _lib.call("vfr_sel_filter", p.bank_tc, self.n_clips, b.dim, p.q_tc, Q, self.k, p.topk_ws, p.n_split, 0, -1, 0,
                              stream)
                    _mark(marks, "scan_end")
                    _lib.call("vfr_sel_refine", p.bank_clips, p.vid_off, p.mom_off, b.n_videos, self.n_clips, b.n_max, b.dim,
                              p.q_tc, p.q_emb, Q, self.k, p.id_base, self.out_s.data_ptr(), self.out_i.data_ptr(), p.topk_ws,
                              p.n_split, stream)
                else:
                    _lib.call("vfr_search_score_device", C.byref(p), Q, self.k, self.out_s.data_ptr(), self.out_i.data_ptr(),
                              stream)
            self._sel_fixup(Q)
            return self.out_s[:Q], self.out_i[:Q]
        per = slice_rows(Q, self.world)
        q0, q1 = self.owned_range(Q)
        self.embed_only(tokens_dev)
        _mark(marks, "k3_done")
        self._marks = marks
        if self.plan.engine == 4:
            self._sel_score_sharded(Q, stream, per=per)
            self._marks = None
            blk = _lib.load().vfr_topk_block_bytes(per, self.k)
            # (a batch smaller than max_queries has shorter records: always the dense [P, blk] prefix of the buffers)
            send = self.send_blocks.view(-1)[:self.world * blk].view(self.world, blk)
            recv = self.recv_blocks.view(-1)[:self.world * blk].view(self.world, blk)
            with self._stage("alltoall_lists"):
                self.comm.all_to_all(send, out=recv)
            with self._stage("k7_merge"):
                _lib.call("vfr_topk_merge_blocks", recv.data_ptr(), self.world, per, q1 - q0, self.k, self.slice_s.data_ptr(),
                          self.slice_i.data_ptr(), self.slice_flags.data_ptr(), None, stream)
        else:
            # the other engines keep the flat exchange: all-gather of [P, Q, k] lists, K7 over the whole batch
            _lib.call("vfr_search_score_device", C.byref(self.plan), Q, self.k, self.out_s.data_ptr(), self.out_i.data_ptr(),
                      stream)
            gs = self.comm.all_gather(self.out_s[:Q])
            gi = self.comm.all_gather(self.out_i[:Q])
            ms, mi = ops.topk_merge(gs, gi)
            self.slice_s[:q1 - q0].copy_(ms[q0:q1])
            self.slice_i[:q1 - q0].copy_(mi[q0:q1])
            self.slice_flags[:per].zero_()
        # flags of every slice + every rank's bad-token word -> all ranks (tiny), then to pinned host memory
        with self._stage("allgather_flags"):
            if q1 > q0:
                self.slice_flags[per:per + 1].copy_(self.text_ws[:1].view(torch.int32))
            else:
                self.slice_flags[per:per + 1].zero_()            # an empty slice embedded nothing this step
            self.comm.all_gather(self.slice_flags[:per + 1], out=self.flags_all[:self.world * (per + 1)])
            self.host_flags[:self.world * (per + 1)].copy_(self.flags_all[:self.world * (per + 1)], non_blocking=True)
            self.flags_event.record()
        if check:
            self.finish_step(Q)
        return self.slice_s[:q1 - q0], self.slice_i[:q1 - q0]

    def finish_step(self, n_queries):
        """The end-of-step check of a sharded search (the step's only host wait): raises ``IndexError`` on ALL ranks
        if any rank saw an out-of-range token id, and re-runs the queries some shard flagged through the exact
        engine - every rank takes the same branch, because every rank looks at the same gathered flags."""
        if self.world == 1:
            return
        per = slice_rows(n_queries, self.world)
        self.flags_event.synchronize()
        flags = self.host_flags[:self.world * (per + 1)].view(self.world, per + 1).numpy()
        if flags[:, per].any():
            raise IndexError("index out of range in self")
        flat = flags[:, :per].reshape(-1)[:n_queries]
        if flat.any():
            self._sharded_fixup(n_queries, np.nonzero(flat)[0])

    def _sharded_fixup(self, Q, idx):
        """Exact-engine rerun of the flagged queries ``idx`` (global rows, identical on every rank) over every shard,
        flat all-gather + K7, owners patch their slice."""
        q0, q1 = self.owned_range(Q)
        idx_t = torch.as_tensor(idx, device=self.q_emb.device)
        s, i = ops.score_topk(self.bank, self.q_emb[:Q][idx_t].contiguous(), self.k, id_base=int(self.plan.id_base))
        ms, mi = ops.topk_merge(self.comm.all_gather(s), self.comm.all_gather(i))
        mine = torch.as_tensor(np.nonzero((idx >= q0) & (idx < q1))[0], device=self.q_emb.device)
        if mine.numel():
            rows = idx_t[mine] - q0
            self.slice_s[rows] = ms[mine]
            self.slice_i[rows] = mi[mine]
        self.n_fixups += int(len(idx))

    # -- host-buffer step (the call a user makes) -------------------------------------------------------
    def search(self, tokens):
        """tokens: int64 [Q, 20] HOST tensor/array (the whole batch, on every rank) -> (scores, ids) HOST tensors
        [n, k] for the rows ``owned_range(Q)`` (one GPU: all Q rows).  A rank copies only its own slice of the token ids
        to the device and only its own slice of the results back."""
        tokens = torch.as_tensor(tokens, dtype=torch.int64)
        Q = tokens.shape[0]
        if Q > self.max_queries:
            raise ValueError(f"batch of {Q} queries exceeds max_queries={self.max_queries}")
        stream = torch.cuda.current_stream().cuda_stream
        if self.world == 1:
            self.host_tokens[:Q].copy_(tokens)
            _lib.call("vfr_search_host", C.byref(self.plan), self.host_tokens.data_ptr(), Q, self.k,
                      self.host_s.data_ptr(), self.host_i.data_ptr(), stream)
            if self.plan.engine == 4 and int((self._sel_flags(Q) != 0).any().item()):
                self._sel_fixup(Q)
                self.host_s[:Q].copy_(self.out_s[:Q])
                self.host_i[:Q].copy_(self.out_i[:Q])
                torch.cuda.current_stream().synchronize()
            bad = int(self.text_ws[:4].view(torch.int32)[0].item())
            if bad:
                raise IndexError("index out of range in self")
            return self.host_s[:Q], self.host_i[:Q]
        q0, q1 = self.owned_range(Q)
        n = q1 - q0
        tr = self.trace_search                      # optional host-side timeline of the step (seconds per phase, summed)
        t0 = time.perf_counter() if tr is not None else 0.0
        self.host_tokens[:n].copy_(tokens[q0:q1])
        self.tokens_dev[q0:q1].copy_(self.host_tokens[:n], non_blocking=True)
        t1 = time.perf_counter() if tr is not None else 0.0
        s, i = self.search_device(self.tokens_dev[:Q], check=False)
        self.host_s[:n].copy_(s, non_blocking=True)
        self.host_i[:n].copy_(i, non_blocking=True)
        t2 = time.perf_counter() if tr is not None else 0.0
        torch.cuda.current_stream().synchronize()
        t3 = time.perf_counter() if tr is not None else 0.0
        fixups = self.n_fixups
        self.finish_step(Q)
        if tr is not None:
            t4 = time.perf_counter()
            for key, dt in (("stage_inputs", t1 - t0), ("issue_step", t2 - t1), ("wait_gpu", t3 - t2), ("finish", t4 - t3)):
                tr[key] = tr.get(key, 0.0) + dt
            tr["steps"] = tr.get("steps", 0) + 1
        if self.n_fixups != fixups:
            self.host_s[:n].copy_(self.slice_s[:n], non_blocking=True)
            self.host_i[:n].copy_(self.slice_i[:n], non_blocking=True)
            torch.cuda.current_stream().synchronize()
        return self.host_s[:n], self.host_i[:n]

    def h2d_bytes(self, Q):
        """Host -> device bytes of one ``search`` step, summed over all ranks (every rank copies its slice of the ids)."""
        return Q * self.seq_len * 8

    def d2h_bytes(self, Q):
        """Device -> host bytes of one ``search`` step, summed over all ranks: the result slices, and on P ranks the
        gathered flag words every rank reads back."""
        extra = 0 if self.world == 1 else self.world * self.world * (slice_rows(Q, self.world) + 1) * 4
        return Q * self.k * (4 + 8) + extra


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
