"""GPU parity tests of the filter + refine top-k (vfr_sel_topk: one fp16 tcgen05 pass + exact fp32
re-scoring).  The bar is BIT equality with the exact-fp32 engine (vfr_score_topk / vfr_score_full),
which the other test files hold to the reference's own scores; plus a direct check that the rigorous
error bound of the fp16 pass holds with margin."""
import numpy as np
import pytest
import torch

import vfr_b200  # noqa: F401
from vfr_b200 import _lib, ops
from oracle import cal_oracle as orc

pytestmark = pytest.mark.gpu
DEV = "cuda"
CAP = 1024


def _ragged_bank(rng, n_videos, dim, seg_choices, scale=0.25, shared=1.0):
    nseg = rng.choice(np.asarray(seg_choices), size=n_videos)
    vid_off = np.concatenate([[0], np.cumsum(nseg)])
    base = np.repeat(rng.standard_normal((n_videos, dim), dtype=np.float32), nseg, axis=0)
    clips = (shared * base + 0.6 * rng.standard_normal((int(vid_off[-1]), dim), dtype=np.float32)) * scale
    return clips.astype(np.float32), vid_off


def _assert_same_as_exact(bank, q, k, n_split=0, id_base=0):
    es, ei = ops.score_topk(bank, q, k, id_base=id_base)
    gs, gi, flags, _ = ops.score_topk_sel(bank, q, k, id_base=id_base, n_split=n_split, return_flags=True)
    assert int(flags.abs().sum().item()) == 0, flags.unique()
    assert torch.equal(gi, ei), f"{(gi != ei).sum().item()} ids differ"
    assert torch.equal(gs.view(torch.int32), es.view(torch.int32))


def test_sel_matches_reference_scores(golden):
    """Top-k of the reference's own score lists (golden val_eval: ragged 5/6-clip videos)."""
    z, meta = golden("val_eval")
    bank = ops.Bank(torch.from_numpy(z["video_emb"]).to(DEV), z["vid_off"])
    n_keep = z["scores"].shape[0]
    q = torch.from_numpy(z["query_emb"][:n_keep]).to(DEV)
    k = 100
    gs, gi = ops.score_topk_sel(bank, q, k)
    ref = torch.from_numpy(z["scores"]).to(DEV)
    ws, wi = torch.sort(ref, dim=1, stable=True)
    rel = ((gs - ws[:, :k]).abs() / ws[:, :k]).max().item()
    assert rel < 1e-5, rel
    # the reference's scores of the returned moments are inside its own top-k up to the score tolerance
    ref_of_ids = torch.gather(ref, 1, gi)
    assert bool((ref_of_ids <= ws[:, k - 1:k] * (1 + 4e-5)).all())
    assert (gi == wi[:, :k]).float().mean().item() > 0.97


@pytest.mark.parametrize("k", [1, 10, 100, 128])
def test_sel_equals_exact_engine_val(golden, k):
    z, meta = golden("val_eval")
    bank = ops.Bank(torch.from_numpy(z["video_emb"]).to(DEV), z["vid_off"])
    q = torch.from_numpy(z["query_emb"]).to(DEV)           # 48 queries: ragged last query tile
    for n_split in (0, 1, 3):
        _assert_same_as_exact(bank, q, k, n_split=n_split)


@pytest.mark.parametrize("seg,n_videos,n_queries", [((6, 5), 30000, 700), ((1, 2, 3, 6), 9000, 129), ((30,), 3000, 260),
                                                     ((32, 7, 1), 2500, 64)])
def test_sel_equals_exact_engine_ragged(seg, n_videos, n_queries):
    rng = np.random.default_rng(11)
    clips, vid_off = _ragged_bank(rng, n_videos, 100, seg)
    bank = ops.Bank(torch.from_numpy(clips).to(DEV), vid_off)
    q = torch.from_numpy((rng.standard_normal((n_queries, 100), dtype=np.float32) * 0.3)).to(DEV)
    _assert_same_as_exact(bank, q, 100, id_base=12345678901)


@pytest.mark.parametrize("dim", [17, 64, 100, 125])
def test_sel_other_dims(dim):
    rng = np.random.default_rng(5)
    clips, vid_off = _ragged_bank(rng, 4000, dim, (6, 5))
    bank = ops.Bank(torch.from_numpy(clips).to(DEV), vid_off)
    q = torch.from_numpy(rng.standard_normal((200, dim), dtype=np.float32) * 0.25).to(DEV)
    _assert_same_as_exact(bank, q, 50)


def test_sel_near_duplicates_and_ties():
    """Queries that (almost) coincide with bank clips, exact duplicates inside the bank (tied scores:
    the order must be by moment id) and a bank smaller than k."""
    rng = np.random.default_rng(3)
    D = 100
    qs = rng.standard_normal((40, D), dtype=np.float32)
    near = np.repeat(qs, 6, axis=0) + 1e-3 * rng.standard_normal((240, D), dtype=np.float32)
    near[::7] = np.repeat(qs, 6, axis=0)[::7]                     # exact hits: d^2 = D eps^2
    other, _ = _ragged_bank(rng, 300, D, (6,), scale=1.0)
    dup = np.tile(other[:60], (3, 1))                              # three identical copies of 10 videos
    clips = np.concatenate([near, other, dup]).astype(np.float32)
    vid_off = np.arange(clips.shape[0] // 6 + 1) * 6
    bank = ops.Bank(torch.from_numpy(clips).to(DEV), vid_off)
    _assert_same_as_exact(bank, torch.from_numpy(qs).to(DEV), 100)
    tiny = ops.Bank(torch.from_numpy(clips[:12]).to(DEV), np.arange(3) * 6)     # 42 moments < k
    gs, gi = ops.score_topk_sel(tiny, torch.from_numpy(qs).to(DEV), 100)
    es, ei = ops.score_topk(tiny, torch.from_numpy(qs).to(DEV), 100)
    assert torch.equal(gi, ei) and torch.equal(gs.view(torch.int32), es.view(torch.int32))
    assert bool((gi[:, 42:] == -1).all()) and bool(torch.isinf(gs[:, 42:]).all())


@pytest.mark.parametrize("bank_scale,query_scale", [(1e-3, 1e-3), (30.0, 0.02), (0.01, 5.0), (1e4, 1e4)])
def test_sel_operand_scales(bank_scale, query_scale):
    """Power-of-two operand scaling keeps magnitudes far from 1 inside fp16's range."""
    rng = np.random.default_rng(8)
    clips, vid_off = _ragged_bank(rng, 6000, 100, (6, 5), scale=bank_scale)
    bank = ops.Bank(torch.from_numpy(clips).to(DEV), vid_off)
    q = torch.from_numpy(rng.standard_normal((150, 100), dtype=np.float32) * query_scale).to(DEV)
    _assert_same_as_exact(bank, q, 100)


def test_sel_offset_embeddings():
    """A large common offset (all-positive embeddings, as after a ReLU) is the worst case of the error
    bound: the dot products are large and d^2 is a small difference."""
    rng = np.random.default_rng(9)
    clips, vid_off = _ragged_bank(rng, 8000, 100, (6, 5), scale=0.05)
    clips = (clips + 1.0).astype(np.float32)
    bank = ops.Bank(torch.from_numpy(clips).to(DEV), vid_off)
    q = torch.from_numpy((rng.standard_normal((130, 100), dtype=np.float32) * 0.05 + 1.0)).to(DEV)
    _assert_same_as_exact(bank, q, 100)


@pytest.mark.parametrize("offset", [0.0, 1.0])
def test_sel_error_bound_holds(offset):
    """With fewer clips per candidate list than k nothing is ever dropped, so the stage-1 lists hold the
    approximate d^2 of EVERY (query, clip) pair: compare them with float64 and with the bound E."""
    rng = np.random.default_rng(21)
    D, C, Q, k = 100, 128, 128, 128
    clips = (rng.standard_normal((C, D)) * np.exp(rng.uniform(-3, 1, size=(C, 1))) + offset).astype(np.float32)
    qs = (rng.standard_normal((Q, D)) * np.exp(rng.uniform(-3, 1, size=(Q, 1))) + offset).astype(np.float32)
    qs[:8] = np.abs(clips[:8])          # aligned all-positive pairs: sum |q_k v_k| = |q||v|
    clips[:8] = np.abs(clips[:8])
    bank = ops.Bank(torch.from_numpy(clips).to(DEV), np.arange(C // 4 + 1) * 4)
    gs, gi, flags, (qp, ws) = ops.score_topk_sel(bank, torch.from_numpy(qs).to(DEV), k, return_flags=True)
    assert int(flags.abs().sum().item()) == 0
    n_parts, qpad = 2, 512
    cand = ws[:qpad * n_parts * CAP * 8].view(torch.int64).view(qpad, n_parts, CAP).cpu().numpy()
    cnt = ws[qpad * n_parts * CAP * 8:qpad * n_parts * (CAP * 8 + 4)].view(torch.int32).view(qpad, n_parts).cpu().numpy()
    assert cnt[:Q].sum(axis=1).tolist() == [C] * Q
    qmeta = qp[qpad * 256:qpad * 256 + qpad * 16].view(torch.float32).view(qpad, 4).cpu().numpy()
    eps = 1e-6
    d2 = (((clips[None, :, :].astype(np.float64) - qs[:, None, :].astype(np.float64) + eps) ** 2).sum(-1))
    worst = 0.0
    for qi in range(Q):
        keys = np.concatenate([cand[qi, p, :cnt[qi, p]] for p in range(n_parts)])
        ids = (keys & 0xffffffff).astype(np.int64)
        approx = (keys >> 32).astype(np.uint32).view(np.float32).astype(np.float64)
        assert sorted(ids.tolist()) == list(range(C))
        err = np.abs(approx - d2[qi, ids])
        E = qmeta[qi, 3] / 2
        worst = max(worst, float((err / E).max()))
    assert worst < 0.5, worst          # observed error stays below half the rigorous bound


def test_sel_sharded_merge_equals_single_bank():
    rng = np.random.default_rng(4)
    clips, vid_off = _ragged_bank(rng, 20000, 100, (6, 5))
    bank = ops.Bank(torch.from_numpy(clips).to(DEV), vid_off)
    q = torch.from_numpy(rng.standard_normal((300, 100), dtype=np.float32) * 0.3).to(DEV)
    want_s, want_i = ops.score_topk(bank, q, 100)
    P = 4
    parts_s, parts_i = [], []
    for r in range(P):
        v0, v1 = 20000 * r // P, 20000 * (r + 1) // P
        c0, c1 = int(vid_off[v0]), int(vid_off[v1])
        shard = ops.Bank(torch.from_numpy(clips[c0:c1]).to(DEV), vid_off[v0:v1 + 1] - c0)
        s, i = ops.score_topk_sel(shard, q, 100, id_base=int(bank.mom_off_host[v0]))
        parts_s.append(s)
        parts_i.append(i)
    ms, mi = ops.topk_merge(torch.stack(parts_s), torch.stack(parts_i))
    assert torch.equal(mi, want_i) and torch.equal(ms.view(torch.int32), want_s.view(torch.int32))


def test_sel_large_properties():
    """131 k videos x 2 k queries (beyond the oracle's reach): sortedness, id range, no duplicates, and
    the oracle's re-scoring of the winners."""
    g = torch.Generator(device=DEV).manual_seed(1)
    V, Q, k = 131072, 2048, 100
    clips = (torch.randn(V, 1, 100, device=DEV, generator=g) + 0.6 * torch.randn(V, 6, 100, device=DEV, generator=g)).reshape(-1, 100) * 0.05
    bank = ops.Bank(clips, np.arange(V + 1) * 6)
    q = torch.randn(Q, 100, device=DEV, generator=g) * 0.06
    gs, gi, flags, _ = ops.score_topk_sel(bank, q, k, return_flags=True)
    assert int(flags.abs().sum().item()) == 0
    es, ei = ops.score_topk(bank, q, k)
    assert torch.equal(gi, ei) and torch.equal(gs.view(torch.int32), es.view(torch.int32))
    assert bool((gs[:, 1:] >= gs[:, :-1]).all()) and bool((gi >= 0).all()) and bool((gi < bank.m_total).all())
    for row in gi[:8].cpu().tolist():
        assert len(set(row)) == k
    vid = (gi[:4, :5] // 21).cpu().numpy()
    mom = (gi[:4, :5] % 21).cpu().numpy()
    moments = orc.generate_moments(6)
    c = clips.cpu().numpy()
    qq = q[:4].cpu().numpy()
    for a in range(4):
        for b in range(5):
            s, e = moments[mom[a, b]]
            want = orc.moment_scores_loop(torch.from_numpy(c[vid[a, b] * 6:vid[a, b] * 6 + 6]), torch.from_numpy(qq[a:a + 1]), [(s, e)])[0]
            assert abs(want - gs[a, b].item()) / want < 1e-5


def test_sel_cluster_multicast_variant(monkeypatch):
    """VFR_SEL_CL=2: CTA pairs share every bank tile through TMA multicast (odd query-group counts get an idle
    partner CTA).  Same bits as the exact engine."""
    monkeypatch.setenv("VFR_SEL_CL", "2")
    rng = np.random.default_rng(12)
    clips, vid_off = _ragged_bank(rng, 20000, 100, (6, 5))
    bank = ops.Bank(torch.from_numpy(clips).to(DEV), vid_off)
    for n_queries in (700, 300, 129):            # 3, 2 (odd -> padded) and 1 query groups of 256
        q = torch.from_numpy((rng.standard_normal((n_queries, 100), dtype=np.float32) * 0.3)).to(DEV)
        _assert_same_as_exact(bank, q, 100)


def _sample_rank(k, sample_tiles, n_tiles):
    """The rank j the sample pass uses (sl_sample_plan in vfr_select_tc.cu): smallest j with
    P(Poisson(k * sampled clips / bank clips) >= j) < 1e-10."""
    import math
    x = k * 256.0 * sample_tiles / ((n_tiles - 1) * 256.0)
    term, cdf = math.exp(-x), 0.0
    for j in range(1, 33):
        cdf += term
        term *= x / j
        if 1.0 - cdf < 1e-10:
            return j
    return None


@pytest.mark.parametrize("n_queries", [300, 100])         # two query tiles per CTA / one (two lists per query)
def test_sel_sampled_starting_threshold(monkeypatch, n_queries):
    """Banks with >= 128 tiles per candidate list start the scan from a sampled threshold (a strided sample of the
    bank, stage 2 verifies it): same bits as the exact engine, and as the engine without the sample pass."""
    rng = np.random.default_rng(41)
    clips, vid_off = _ragged_bank(rng, 30000, 100, (6, 5))
    bank = ops.Bank(torch.from_numpy(clips).to(DEV), vid_off)
    q = torch.from_numpy((rng.standard_normal((n_queries, 100), dtype=np.float32) * 0.3)).to(DEV)
    for k in (1, 20, 100):
        _assert_same_as_exact(bank, q, k, n_split=1)
        with_sample = ops.score_topk_sel(bank, q, k, n_split=1)
        monkeypatch.setenv("VFR_SEL_SAMPLE", "0")
        without = ops.score_topk_sel(bank, q, k, n_split=1)
        monkeypatch.delenv("VFR_SEL_SAMPLE")
        assert torch.equal(with_sample[1], without[1]) and torch.equal(with_sample[0], without[0])


def test_sel_wrong_sampled_threshold_is_flagged(monkeypatch):
    """An adversarial bank: the only clips near the queries sit exactly in the tiles the sample pass reads, fewer of
    them than k.  The sampled threshold then promises k clips the bank does not have; stage 2 must notice (flag 4,
    the retriever re-runs such queries through the exact engine) - and without the sample pass the same call is
    exact."""
    rng = np.random.default_rng(42)
    V, D, k, Q = 8000, 100, 20, 300
    n_clips = V * 6
    n_tiles = (n_clips + 255) // 256
    assert n_tiles == 188
    sample_tiles, stride = 8, (n_tiles - 1) // 8          # one list per query (n_split = 1, two query tiles per CTA)
    j = _sample_rank(k, sample_tiles, n_tiles)
    assert j is not None and j < 16 < k
    centre = rng.standard_normal(D).astype(np.float32)
    clips = (centre + 3.0 * rng.standard_normal((n_clips, D))).astype(np.float32)       # far from every query
    planted = [t * stride * 256 + c * 64 + 5 for t in range(sample_tiles) for c in (0, 2)]   # 16 clips, 16 chunks
    clips[planted] = centre + 0.02 * rng.standard_normal((len(planted), D)).astype(np.float32)
    q = torch.from_numpy((centre + 0.01 * rng.standard_normal((Q, D))).astype(np.float32)).to(DEV)
    bank = ops.Bank(torch.from_numpy(clips).to(DEV), np.arange(V + 1) * 6)
    gs, gi, flags, _ = ops.score_topk_sel(bank, q, k, n_split=1, return_flags=True)
    assert bool((flags == 4).all()), flags.unique()
    monkeypatch.setenv("VFR_SEL_SAMPLE", "0")
    _assert_same_as_exact(bank, q, k, n_split=1)


# ---- joint spaces wider than one 128-column row (BASELINE configs[2]: 1024-d): the K-streaming kernel ----------------
@pytest.mark.parametrize("dim,n_videos,n_queries,k", [(126, 4000, 200, 50), (256, 6000, 300, 100), (509, 3000, 129, 100),
                                                       (1024, 30000, 300, 100), (1085, 2000, 64, 10), (1024, 3000, 20, 1)])
def test_sel_large_dims_equal_exact_engine(dim, n_videos, n_queries, k):
    """D + 3 > 128: both operands are streamed in 64-column chunks and the TMEM accumulators integrate over the chunks.
    Same bar as at D = 100: ids and score BITS of the exact-fp32 engine, no query flagged; the 30 000-video case is
    large enough for the sampled starting threshold (n_split = 1)."""
    rng = np.random.default_rng(dim)
    clips, vid_off = _ragged_bank(rng, n_videos, dim, (6, 5), scale=0.25 / np.sqrt(dim / 100.0))
    bank = ops.Bank(torch.from_numpy(clips).to(DEV), vid_off)
    q = torch.from_numpy(rng.standard_normal((n_queries, dim), dtype=np.float32) * np.float32(0.25 / np.sqrt(dim / 100.0))).to(DEV)
    _assert_same_as_exact(bank, q, k)
    _assert_same_as_exact(bank, q, k, n_split=1, id_base=987654321012)


@pytest.mark.parametrize("dim", [256, 1024])
def test_sel_large_dims_near_duplicates_offsets_and_ties(dim):
    rng = np.random.default_rng(dim + 1)
    qs = (rng.standard_normal((40, dim), dtype=np.float32) * 0.1 + 1.0).astype(np.float32)      # common offset
    near = np.repeat(qs, 6, axis=0) + 1e-3 * rng.standard_normal((240, dim), dtype=np.float32)
    near[::7] = np.repeat(qs, 6, axis=0)[::7]                                                   # exact hits
    other, _ = _ragged_bank(rng, 400, dim, (6,), scale=0.1)
    other = (other + 1.0).astype(np.float32)
    dup = np.tile(other[:60], (3, 1))
    clips = np.concatenate([near, other, dup]).astype(np.float32)
    bank = ops.Bank(torch.from_numpy(clips).to(DEV), np.arange(clips.shape[0] // 6 + 1) * 6)
    _assert_same_as_exact(bank, torch.from_numpy(qs).to(DEV), 100)


@pytest.mark.parametrize("dim", [256, 1024])
def test_sel_error_bound_holds_large_dims(dim):
    """The rigorous bound E of the fp16 pass at K = D + 3 up to 1027 (fp32 accumulation over up to 65 MMA steps):
    every approximate d^2 read back from the stage-1 lists stays below E / 2 of the float64 value."""
    rng = np.random.default_rng(21 + dim)
    C, Q, k = 128, 128, 128
    clips = (rng.standard_normal((C, dim)) * np.exp(rng.uniform(-3, 1, size=(C, 1))) / np.sqrt(dim / 100) + 0.5).astype(np.float32)
    qs = (rng.standard_normal((Q, dim)) * np.exp(rng.uniform(-3, 1, size=(Q, 1))) / np.sqrt(dim / 100) + 0.5).astype(np.float32)
    qs[:8] = np.abs(clips[:8])
    clips[:8] = np.abs(clips[:8])
    bank = ops.Bank(torch.from_numpy(clips).to(DEV), np.arange(C // 4 + 1) * 4)
    gs, gi, flags, (qp, ws) = ops.score_topk_sel(bank, torch.from_numpy(qs).to(DEV), k, return_flags=True)
    assert int(flags.abs().sum().item()) == 0
    n_parts, qpad = 1, 512                                   # two query tiles per CTA, one list per query
    pitch = (dim + 3 + 63) // 64 * 64
    cand = ws[:qpad * n_parts * CAP * 8].view(torch.int64).view(qpad, n_parts, CAP).cpu().numpy()
    cnt = ws[qpad * n_parts * CAP * 8:qpad * n_parts * (CAP * 8 + 4)].view(torch.int32).view(qpad, n_parts).cpu().numpy()
    assert cnt[:Q].sum(axis=1).tolist() == [C] * Q
    qmeta = qp[qpad * pitch * 2:qpad * pitch * 2 + qpad * 16].view(torch.float32).view(qpad, 4).cpu().numpy()
    d2 = (((clips[None, :, :].astype(np.float64) - qs[:, None, :].astype(np.float64) + 1e-6) ** 2).sum(-1))
    worst = 0.0
    for qi in range(Q):
        keys = cand[qi, 0, :cnt[qi, 0]]
        ids = (keys & 0xffffffff).astype(np.int64)
        approx = (keys >> 32).astype(np.uint32).view(np.float32).astype(np.float64)
        assert sorted(ids.tolist()) == list(range(C))
        worst = max(worst, float((np.abs(approx - d2[qi, ids]) / (qmeta[qi, 3] / 2)).max()))
    assert worst < 0.5, worst


@pytest.mark.parametrize("dim,n_videos", [(100, 20000), (1024, 6000)])
def test_bf16_embedding_path_tolerance(dim, n_videos):
    """BASELINE configs[2]: embeddings stored in bf16 (bank rows AND queries), scored by the filter + refine engine whose
    stage 2 reads the bf16 bank (vfr_sel_topk_b16).  (1) the scores are the exact engine's scores of the rounded
    embeddings, bit for bit; (2) against the fp32 embeddings the scores stay within the stated 1e-2 relative tolerance
    (observed ~1e-3 at D = 100, less at D = 1024), and the top-1 moment agrees wherever the fp32 top-2 gap exceeds it."""
    rng = np.random.default_rng(77 + dim)
    clips, vid_off = _ragged_bank(rng, n_videos, dim, (6, 5), scale=0.25 / np.sqrt(dim / 100.0))
    q = rng.standard_normal((200, dim), dtype=np.float32) * np.float32(0.25 / np.sqrt(dim / 100.0))
    clips_t, q_t = torch.from_numpy(clips).to(DEV), torch.from_numpy(q).to(DEV)
    k = 50
    fs, fi = ops.score_topk(ops.Bank(clips_t, vid_off), q_t, k)                               # fp32 embeddings
    bank16 = ops.Bank(clips_t.to(torch.bfloat16), vid_off)                                    # values rounded to bf16
    bs, bi, flags, _ = ops.score_topk_sel(bank16, q_t, k, return_flags=True, bf16=True)
    assert int(flags.abs().sum().item()) == 0
    es, ei = ops.score_topk(bank16, q_t.to(torch.bfloat16).float(), k)
    assert torch.equal(bi, ei) and torch.equal(bs.view(torch.int32), es.view(torch.int32))
    rel = ((bs - fs).abs() / fs).max().item()
    assert rel < 1e-2, rel
    gap = (fs[:, 1] - fs[:, 0]) / fs[:, 0]
    clear = gap > 2e-2
    assert bool((bi[clear, 0] == fi[clear, 0]).all())
    with pytest.raises(_lib.VfrError):                                                        # an fp32 bank is refused
        ops.score_topk_sel(ops.Bank(clips_t, vid_off), q_t, k, bf16=True)


def test_refine_by_query_range_equals_whole_refine():
    """vfr_sel_refine_range over consecutive query ranges (what vfr_search_host does to overlap the result copies with
    stage 2) writes the same rows as ONE vfr_sel_refine - ragged ranges, an empty one, the last one short."""
    rng = np.random.default_rng(11)
    clips, vid_off = _ragged_bank(rng, 3000, 100, (5, 6))
    bank = ops.Bank(torch.from_numpy(clips).to(DEV), vid_off)
    Q, k = 301, 37
    q = torch.from_numpy(rng.standard_normal((Q, 100), dtype=np.float32) * 0.25).to(DEV)
    ws_s, ws_i, flags, (qp, ws) = ops.score_topk_sel(bank, q, k, return_flags=True)      # filter + whole refine
    n_clips = int(bank.clips.shape[0])
    out_s = torch.full((Q, k), -1.0, dtype=torch.float32, device=DEV)
    out_i = torch.full((Q, k), -1, dtype=torch.int64, device=DEV)
    stream = torch.cuda.current_stream().cuda_stream
    for q0, n in ((0, 1), (1, 127), (128, 0), (128, 128), (256, 45)):
        _lib.call("vfr_sel_refine_range", bank.clips.data_ptr(), bank.vid_off.data_ptr(), bank.mom_off.data_ptr(),
                  bank.n_videos, n_clips, bank.n_max, bank.dim, qp.data_ptr(), q.data_ptr(), Q, k, 0, out_s.data_ptr(),
                  out_i.data_ptr(), ws.data_ptr(), 0, q0, n, stream)
    assert torch.equal(out_i, ws_i)
    assert torch.equal(out_s.view(torch.int32), ws_s.view(torch.int32))
    with pytest.raises(_lib.VfrError):
        _lib.call("vfr_sel_refine_range", bank.clips.data_ptr(), bank.vid_off.data_ptr(), bank.mom_off.data_ptr(),
                  bank.n_videos, n_clips, bank.n_max, bank.dim, qp.data_ptr(), q.data_ptr(), Q, k, 0, out_s.data_ptr(),
                  out_i.data_ptr(), ws.data_ptr(), 0, 300, 2, stream)


@pytest.mark.parametrize("n_src,width", [(8, 32), (3, 32), (2, 64), (1, 32), (16, 32)])
def test_pool_levels_merge_equals_counting(n_src, width):
    """The j-th smallest pooled sample values of every query: the 8-lane merge of the sorted runs (<= 8 runs), the
    binary-search ranking (sorted runs, more than 8) and the plain counting kernel (sorted_runs = 0) agree, with ties
    (quantised values) and +inf padding."""
    import ctypes as C
    rng = np.random.default_rng(n_src * 100 + width)
    Q = 333
    vals = np.round(rng.random((n_src, Q, width // 32, 32)).astype(np.float32) * 40) / 40      # many ties
    n_valid = rng.integers(0, 33, size=(n_src, Q, width // 32, 1))
    vals = np.where(np.arange(32).reshape(1, 1, 1, 32) < n_valid, vals, np.inf).astype(np.float32)
    vals = np.sort(vals, axis=-1).reshape(n_src, Q, width)
    pooled = torch.from_numpy(vals).to(DEV)
    total = n_src * width
    ranks = sorted({1, min(7, total), min(19, total), min(32, total)}, reverse=True)
    L = len(ranks)
    out = []
    for sorted_runs in (1, 0):
        lv = torch.full((L * Q,), -1.0, dtype=torch.float32, device=DEV)
        _lib.call("vfr_sel_pool_levels", pooled.data_ptr(), n_src, Q, width, (C.c_int32 * L)(*ranks), L, sorted_runs,
                  lv.data_ptr(), torch.cuda.current_stream().cuda_stream)
        out.append(lv.cpu().numpy().reshape(L, Q))
    want = np.sort(vals.transpose(1, 0, 2).reshape(Q, total), axis=1)[:, [r - 1 for r in ranks]].T
    np.testing.assert_array_equal(out[1], want)
    np.testing.assert_array_equal(out[0], want)
