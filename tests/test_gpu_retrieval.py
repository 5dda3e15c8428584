"""GPU tests of the retrieval step (K3 -> K4 behind vfr_search_device / vfr_search_host)."""
import numpy as np
import pytest
import torch

import vfr_b200  # noqa: F401
from vfr_b200 import models, ops, synth
from vfr_b200.retrieval import MomentRetriever, shard_range
from oracle import cal_oracle as orc

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _setup(seed=21, V=300, Q=70, vocab=400):
    sd = synth.make_state_dict(seed, 8, vocab, spread=4.0)
    model = models.CALModel(visual_input_dim=18, pretrained_emb=torch.from_numpy(sd["word_embedding.weight"]))
    model.load_state_dict({k: torch.from_numpy(v) for k, v in sd.items()})
    model = model.to(DEV).eval()
    clips = synth.make_bank(seed, V, 6, 100)
    videos = synth.make_videos(seed, 4, 8)
    tokens = synth.make_queries(seed, videos, Q, vocab)["tokens"]
    return sd, model, clips, tokens


@pytest.mark.parametrize("engine", ["exact", "tc", "sel"])
def test_search_host_matches_oracle_and_device_path(engine):
    sd, model, clips, tokens = _setup()
    V = clips.shape[0] // 6
    vid_off = np.arange(V + 1) * 6
    text_engine = "tc" if engine == "sel" else engine
    retr = MomentRetriever(model, torch.from_numpy(clips).to(DEV), vid_off, max_queries=128, k=10, engine=engine,
                           text_engine=text_engine)
    assert retr.engine == engine and retr.text_engine == text_engine
    s, i = retr.search(tokens)
    sd_, id_ = retr.search_device(torch.from_numpy(tokens).to(DEV))
    assert torch.equal(s, sd_.cpu()) and torch.equal(i, id_.cpu())
    q_emb = orc.text_embed(sd, tokens)
    full = orc.score_matrix(clips, vid_off, q_emb).numpy()
    order = np.argsort(full, axis=1, kind="stable")[:, :10]
    want = np.take_along_axis(full, order, axis=1)
    np.testing.assert_allclose(s.numpy(), want, rtol=1e-5)
    # ids agree wherever the neighbouring scores are separated by more than the tolerance
    gap_ok = np.abs(np.diff(np.take_along_axis(full, np.argsort(full, axis=1)[:, :11], axis=1), axis=1)) > 1e-5 * want
    agree = (i.numpy() == order)
    assert agree[gap_ok & np.roll(gap_ok, 1, axis=1)].mean() > 0.99
    with pytest.raises(ValueError):
        retr.search(np.zeros((200, 20), dtype=np.int64))
    bad = tokens.copy()
    bad[3, 0] = 10 ** 6
    with pytest.raises(IndexError):
        retr.search(bad)


def test_search_host_large_batch_takes_the_chunked_copy_path():
    """From 4 096 queries on vfr_search_host re-scores in query chunks and copies finished rows to the host on a second
    stream: the host result equals the device-resident one, row for row (a batch that is not a multiple of the chunks)."""
    sd, model, clips, _ = _setup(V=2000)
    V = clips.shape[0] // 6
    rng = np.random.default_rng(3)
    Q = 4099
    tokens = np.zeros((Q, 20), dtype=np.int64)
    for r in range(Q):
        n = 1 + int(rng.integers(0, 12))
        tokens[r, :n] = rng.integers(1, 400, size=n)
    retr = MomentRetriever(model, torch.from_numpy(clips).to(DEV), np.arange(V + 1) * 6, max_queries=Q, k=25, engine="sel",
                           text_engine="tc")
    for _ in range(2):                      # (second call: the copy stream and its events are reused)
        s, i = retr.search(tokens)
        sd_, id_ = retr.search_device(torch.from_numpy(tokens).to(DEV))
        assert torch.equal(i, id_.cpu())
        assert torch.equal(s.view(torch.int32), sd_.cpu().view(torch.int32))


@pytest.mark.parametrize("engine", ["exact", "tc", "tc_bf16", "sel"])
def test_sharded_search_equals_single_bank_search(engine):
    sd, model, clips, tokens = _setup(seed=22, V=1000)
    V = clips.shape[0] // 6
    full = MomentRetriever(model, torch.from_numpy(clips).to(DEV), np.arange(V + 1) * 6, max_queries=128, k=100,
                           engine=engine)
    s, i = full.search_device(torch.from_numpy(tokens).to(DEV))
    parts_s, parts_i = [], []
    for r in range(4):
        v0, v1 = shard_range(V, r, 4)
        shard = MomentRetriever(model, torch.from_numpy(clips[v0 * 6:v1 * 6]).to(DEV), np.arange(v1 - v0 + 1) * 6,
                                id_base=v0 * 21, max_queries=128, k=100, engine=engine)
        ps, pi = shard.search_device(torch.from_numpy(tokens).to(DEV))
        parts_s.append(ps.clone())
        parts_i.append(pi.clone())
    ms, mi = ops.topk_merge(torch.stack(parts_s), torch.stack(parts_i))
    assert torch.equal(ms, s) and torch.equal(mi, i)


def test_sel_engine_equals_exact_engine_and_fixes_up_flagged_queries():
    """The default engine ("auto" -> filter + refine) returns the exact engine's bits; queries it cannot
    certify (here: a bank scaled so far from the queries that the fp16 operand scales do not fit) are
    re-run through the exact engine by the retriever."""
    sd, model, clips, tokens = _setup(seed=23, V=2000, Q=100)
    V = clips.shape[0] // 6
    vid_off = np.arange(V + 1) * 6
    tok = torch.from_numpy(tokens).to(DEV)
    for scale, expect_fixups in ((1.0, False), (1e33, True)):
        bank = torch.from_numpy(clips * np.float32(scale)).to(DEV)
        exact = MomentRetriever(model, bank, vid_off, max_queries=128, k=100, engine="exact", text_engine="tc")
        auto = MomentRetriever(model, bank, vid_off, max_queries=128, k=100, text_engine="tc")
        assert auto.engine == "sel"
        es, ei = exact.search_device(tok)
        gs, gi = auto.search_device(tok)
        assert torch.equal(gi, ei) and torch.equal(gs.view(torch.int32), es.view(torch.int32))
        assert (getattr(auto, "n_fixups", 0) > 0) == expect_fixups
        hs, hi = auto.search(tokens)
        assert torch.equal(hi, ei.cpu()) and torch.equal(hs.view(torch.int32), es.cpu().view(torch.int32))


def test_sharded_rank_of_first_positive_equals_single_bank(golden):
    """Exact R@k / median-rank inputs over a sharded bank: tau by all-reduce(min), counts by all-reduce(sum)
    (here the reducers combine four emulated shards on one GPU) == the unsharded evaluation core."""
    from vfr_b200 import evaluate as vev
    from vfr_b200.retrieval import sharded_rank_first_positive
    z, meta = golden("val_eval")
    clips, vid_off = z["video_emb"], z["vid_off"]
    bank = ops.Bank(torch.from_numpy(clips).to(DEV), vid_off)
    videos = synth.make_videos(meta["seed"], meta["n_videos"], meta["feat_dim"], tuple(meta["seg_choices"]), tuple(meta["seg_probs"]))
    queries = synth.make_queries(meta["seed"], videos, meta["n_queries"], meta["vocab"])
    q_emb = torch.from_numpy(z["query_emb"]).to(DEV)
    Q = q_emb.shape[0]
    q_video = queries["video_idx"][:Q]
    times = queries["times"][:Q]
    want = vev.rank_first_positive(bank, q_emb, q_video, times, [0.5, 0.7])
    V, P = bank.n_videos, 4
    shards, v0s = [], []
    for r in range(P):
        a, b = shard_range(V, r, P)
        c0, c1 = int(vid_off[a]), int(vid_off[b])
        shards.append(ops.Bank(torch.from_numpy(clips[c0:c1]).to(DEV), vid_off[a:b + 1] - c0))
        v0s.append(a)
    # run the P ranks "in lockstep": every reduce call combines the tensors the P ranks pass at the same call site
    import threading
    barrier = threading.Barrier(P)
    slots, results, lock = {}, [None] * P, threading.Lock()

    def make_reduce(rank):
        calls = [0]

        def reduce(t, op):
            key = calls[0]
            calls[0] += 1
            with lock:
                slots.setdefault(key, []).append(t.clone())
            barrier.wait()
            parts = torch.stack(slots[key])
            barrier.wait()
            return parts.min(dim=0).values if op == "min" else parts.sum(dim=0)
        return reduce

    def work(rank):
        torch.cuda.set_device(0)
        results[rank] = sharded_rank_first_positive(shards[rank], v0s[rank], q_emb, q_video, bank.nseg_host[q_video], times,
                                                    [0.5, 0.7], reduce=make_reduce(rank))
    threads = [threading.Thread(target=work, args=(r,)) for r in range(P)]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    for r in range(P):
        rank, npos = results[r]
        assert np.array_equal(rank.cpu().numpy(), want["rank"])
        assert np.array_equal(npos.cpu().numpy(), want["npos"])


class _LockstepComm:
    """The collectives of P emulated ranks (threads on one GPU): every call combines what the P ranks pass at the same
    call site."""

    def __init__(self, n_ranks):
        import threading
        self.n, self.barrier, self.lock, self.slots, self.log = n_ranks, threading.Barrier(n_ranks), threading.Lock(), {}, []

    def rank_view(self, rank):
        comm, calls = self, [0]

        class View:
            def _exchange(self, t, kind):
                key = calls[0]
                calls[0] += 1
                with comm.lock:
                    comm.slots.setdefault(key, {})[rank] = t.clone()
                comm.barrier.wait()
                parts = torch.stack([comm.slots[key][r] for r in range(comm.n)])
                comm.barrier.wait()
                if rank == 0:
                    comm.log.append(kind)
                return parts

            def all_gather(self, t, out=None):
                parts = self._exchange(t, "all_gather")
                if out is None:
                    return parts
                out.view(parts.shape).copy_(parts)
                return out.view(parts.shape)

            def all_to_all(self, send, out=None):
                parts = self._exchange(send, "all_to_all")          # [P (src), P (dst), n]
                res = parts[:, rank].contiguous()
                if out is None:
                    return res
                out.copy_(res)
                return out

            def all_reduce_sum(self, t):
                return self._exchange(t, "sum").sum(dim=0).to(t.dtype)

            def all_reduce_min(self, t):
                return self._exchange(t, "min").min(dim=0).values
        return View()


@pytest.mark.parametrize("pool,first_tiles,k,n_videos", [(False, None, 100, 9000), (False, 3, 100, 9000), (False, 1, 1, 9000),
                                                           (False, 5, 37, 9000), (True, None, 100, 120000),
                                                           (True, None, 1, 120000), (True, None, 37, 120000)])
def test_sel_sharded_protocols_equal_single_bank(pool, first_tiles, k, n_videos):
    """N-GPU K4: the shards agree on a threshold - by pooling their samples (large shards: every shard then keeps ~k/P
    candidates) or by exchanging a certified bound mid-scan.  Four emulated shards on one GPU, the collectives replaced
    by a lockstep exchange; the merged lists are bit-identical to the exact engine over the whole bank."""
    import threading
    sd, model, clips, tokens = _setup(seed=31, V=n_videos, Q=150)
    V, P = clips.shape[0] // 6, 4
    tok = torch.from_numpy(tokens).to(DEV)
    Q = tok.shape[0]
    exact = MomentRetriever(model, torch.from_numpy(clips).to(DEV), np.arange(V + 1) * 6, max_queries=256, k=k,
                            engine="exact", text_engine="tc")
    es, ei = exact.search_device(tok)
    q_emb = exact.q_emb[:Q].clone()
    shards = []
    for r in range(P):
        v0, v1 = shard_range(V, r, P)
        sh = MomentRetriever(model, torch.from_numpy(clips[v0 * 6:v1 * 6]).to(DEV), np.arange(v1 - v0 + 1) * 6,
                             id_base=v0 * 21, max_queries=256, k=k, engine="sel", text_engine="tc", n_split=1 if pool else 0)
        sh.sel_first_tiles = first_tiles
        sh.sel_pool_samples = pool
        sh.q_emb[:Q].copy_(q_emb)
        shards.append(sh)
    comm = _LockstepComm(P)
    errors = []

    def work(r):
        try:
            torch.cuda.set_device(0)
            stream = torch.cuda.current_stream().cuda_stream
            shards[r]._sel_score_sharded(Q, stream, comm=comm.rank_view(r))
            shards[r]._sel_fixup(Q)
        except Exception as e:  # noqa: BLE001
            errors.append(e)
            comm.barrier.abort()
    threads = [threading.Thread(target=work, args=(r,)) for r in range(P)]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    assert not errors, errors
    # the protocol that was meant to run did run
    assert comm.log == (["sum", "all_gather", "sum"] if pool else ["sum", "min"]), comm.log
    ms, mi = ops.topk_merge(torch.stack([sh.out_s[:Q] for sh in shards]), torch.stack([sh.out_i[:Q] for sh in shards]))
    assert torch.equal(mi, ei) and torch.equal(ms.view(torch.int32), es.view(torch.int32))
    assert all(int((sh._sel_flags(Q) != 0).sum().item()) == 0 for sh in shards)
    if pool:
        # every shard kept only what can reach the global top-k: together about k candidates (plus the error band), not P k
        kept = sum(int((sh.out_i[:Q] >= 0).sum().item()) for sh in shards)
        assert kept >= Q * k


def _run_ranks(P, fn):
    import threading
    comm = _LockstepComm(P)
    errors, results = [], [None] * P

    def work(r):
        try:
            torch.cuda.set_device(0)
            results[r] = fn(r, comm.rank_view(r))
        except Exception as e:  # noqa: BLE001
            errors.append(e)
            comm.barrier.abort()
    threads = [threading.Thread(target=work, args=(r,)) for r in range(P)]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    assert not errors, errors
    return results, comm


@pytest.mark.parametrize("n_videos,Q,k,bounds", [
    (120000, 150, 100, None),                 # pooled samples, equal shards
    (120000, 151, 37, None),                  # odd batch: the last slice is short
    (9000, 70, 10, None),                     # shards too small to sample: bound exchange
    (9000, 150, 100, [0, 300, 700, 8000, 9000]),   # UNEVEN shards: some finish their tiles before the exchange
    (120000, 3, 5, None),                     # fewer queries than ranks: empty slices
])
def test_sharded_search_step_by_query_slice_equals_single_bank(n_videos, Q, k, bounds):
    """The whole P-rank step (K3 by slice -> all-gather -> K4 with the shard threshold protocol -> all-to-all of the
    query-slice records -> K7 -> flags gather) on four emulated ranks: every rank returns its slice, and the slices
    put together are bit-identical to the exact engine over the whole bank.  No collective is skipped by any rank,
    whatever its shard size (the exchange log is the same list on every rank by construction of the lockstep comm)."""
    sd, model, clips, tokens = _setup(seed=33, V=n_videos, Q=Q)
    V, P = clips.shape[0] // 6, 4
    tok = torch.from_numpy(tokens).to(DEV)
    exact = MomentRetriever(model, torch.from_numpy(clips).to(DEV), np.arange(V + 1) * 6, max_queries=256, k=k,
                            engine="exact", text_engine="tc")
    es, ei = exact.search_device(tok)
    es, ei = es.clone(), ei.clone()
    bounds = bounds or [shard_range(V, r, P)[0] for r in range(P)] + [V]
    shards = [None] * P

    def step(r, comm):
        v0, v1 = bounds[r], bounds[r + 1]
        sh = MomentRetriever(model, torch.from_numpy(clips[v0 * 6:v1 * 6]).to(DEV), np.arange(v1 - v0 + 1) * 6,
                             id_base=v0 * 21, max_queries=256, k=k, engine="sel", text_engine="tc", comm=comm, world=P, rank=r)
        shards[r] = sh
        s, i = sh.search_device(tok)
        q0, q1 = sh.owned_range(Q)
        # the host-buffer call returns the same slice
        hs, hi = sh.search(tokens)
        assert torch.equal(hs, s.cpu()) and torch.equal(hi, i.cpu())
        return q0, q1, s.clone(), i.clone()
    results, comm = _run_ranks(P, step)
    rows = 0
    for q0, q1, s, i in results:
        assert s.shape[0] == q1 - q0
        assert torch.equal(i, ei[q0:q1]) and torch.equal(s.view(torch.int32), es[q0:q1].view(torch.int32))
        rows += q1 - q0
    assert rows == Q
    assert "all_to_all" in comm.log and sum(sh.n_fixups for sh in shards) == 0


def test_sharded_search_flagged_queries_and_bad_tokens_are_handled_by_all_ranks_together():
    """A bank the fp16 scales cannot hold flags every query on every shard: all ranks take the exact-engine rerun
    together (same collectives everywhere) and still return the exact engine's bits; an out-of-range token id seen by
    ONE rank raises IndexError on ALL ranks."""
    sd, model, clips, tokens = _setup(seed=34, V=4000, Q=60)
    V, P, k = clips.shape[0] // 6, 4, 20
    tok = torch.from_numpy(tokens).to(DEV)
    big = clips * np.float32(1e33)
    exact = MomentRetriever(model, torch.from_numpy(big).to(DEV), np.arange(V + 1) * 6, max_queries=64, k=k, engine="exact",
                            text_engine="tc")
    es, ei = exact.search_device(tok)
    es, ei = es.clone(), ei.clone()

    def step(r, comm):
        v0, v1 = shard_range(V, r, P)
        sh = MomentRetriever(model, torch.from_numpy(big[v0 * 6:v1 * 6]).to(DEV), np.arange(v1 - v0 + 1) * 6,
                             id_base=v0 * 21, max_queries=64, k=k, engine="sel", text_engine="tc", comm=comm, world=P, rank=r)
        s, i = sh.search_device(tok)
        q0, q1 = sh.owned_range(tok.shape[0])
        assert sh.n_fixups == tok.shape[0]
        assert torch.equal(i, ei[q0:q1]) and torch.equal(s.view(torch.int32), es[q0:q1].view(torch.int32))
        bad = tokens.copy()
        bad[1, 0] = 10 ** 6                      # row 1 lives in rank 0's slice only
        with pytest.raises(IndexError):
            sh.search(bad)
        return True
    results, _ = _run_ranks(P, step)
    assert all(results)
