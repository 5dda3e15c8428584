"""GPU parity tests of K4 (scoring / count / top-k / merge) and K5 (IoU, ranks) through the C ABI,
against (a) golden outputs of the unmodified reference and (b) the CPU oracle on seeded inputs."""
import random

import numpy as np
import pytest
import torch

import vfr_b200  # noqa: F401
from vfr_b200 import evaluate as vev
from vfr_b200 import evaluate_single as vsingle
from vfr_b200 import ops, synth
from oracle import cal_oracle as orc

pytestmark = pytest.mark.gpu
DEV = "cuda"
SCORE_RTOL = 1e-5   # north-star tolerance for fp32 scores


def _bank(z):
    return ops.Bank(torch.from_numpy(z["video_emb"]).to(DEV), z["vid_off"])


def _queries(meta):
    videos = synth.make_videos(meta["seed"], meta["n_videos"], meta["feat_dim"],
                               tuple(meta["seg_choices"]), tuple(meta["seg_probs"]))
    return videos, synth.make_queries(meta["seed"], videos, meta["n_queries"], meta["vocab"])


@pytest.mark.parametrize("case", ["tiny_eval", "long_eval", "val_eval"])
def test_score_full_matches_reference(golden, case):
    z, meta = golden(case)
    bank = _bank(z)
    n_keep = z["scores"].shape[0]
    q = torch.from_numpy(z["query_emb"][:n_keep]).to(DEV)
    got = ops.score_full(bank, q).cpu().numpy()
    assert got.shape == z["scores"].shape
    rel = np.abs(got - z["scores"]) / np.abs(z["scores"])
    assert rel.max() < SCORE_RTOL, rel.max()


@pytest.mark.parametrize("case", ["tiny_eval", "long_eval", "val_eval"])
def test_rank_of_first_positive_matches_reference(golden, case):
    z, meta = golden(case)
    _, queries = _queries(meta)
    bank = _bank(z)
    q = torch.from_numpy(z["query_emb"]).to(DEV)
    res = vev.rank_first_positive(bank, q, queries["video_idx"], queries["times"], [0.5, 0.7])
    full32 = ops.score_full(bank, q).cpu().numpy()
    n_keep = z["scores"].shape[0]
    # tie band = a few times the score error actually observed against the reference (>= 1 ulp)
    err = float((np.abs(full32[:n_keep] - z["scores"]) / z["scores"]).max())
    band = max(4 * err, 2.0 ** -22)
    assert band < SCORE_RTOL
    full = full32.astype(np.float64)
    n_tied = 0
    for ti, key in enumerate(("rank_05", "rank_07")):
        ref = z[key]
        for qi in range(len(ref)):
            tau = float(res["tau"][qi, ti])
            lo = int((full[qi] < tau * (1 - band)).sum())
            hi = int((full[qi] <= tau * (1 + band)).sum()) - 1
            assert lo <= ref[qi] <= hi, (case, key, qi, lo, ref[qi], hi)
            assert lo <= res["rank"][qi, ti] <= hi
            if hi == lo:   # no score within tolerance of the first positive: must be identical
                assert res["rank"][qi, ti] == ref[qi]
            else:
                n_tied += 1
    assert n_tied <= 0.2 * 2 * len(z["rank_05"]), f"{n_tied} tie-ambiguous queries"


def test_count_mode_is_bitwise_consistent_with_full(golden):
    z, meta = golden("val_eval")
    _, queries = _queries(meta)
    bank = _bank(z)
    q = torch.from_numpy(z["query_emb"]).to(DEV)
    res = vev.rank_first_positive(bank, q, queries["video_idx"], queries["times"], [0.5, 0.7])
    full = ops.score_full(bank, q)
    tau = torch.from_numpy(res["tau"]).to(DEV)
    for ti in range(2):
        lt = (full < tau[:, ti:ti + 1]).sum(dim=1).cpu().numpy()
        assert np.array_equal(lt, res["rank_lo"][:, ti])
        # tau itself is one of the scores, bit for bit
        assert bool(((full == tau[:, ti:ti + 1]).sum(dim=1) >= 1).all())


def test_corpus_metrics_identical_to_reference(golden):
    z, meta = golden("tiny_eval")
    _, queries = _queries(meta)
    bank = _bank(z)
    q = torch.from_numpy(z["query_emb"]).to(DEV)
    np.random.seed(123)
    m = vev.evaluate_embedded(bank, q, queries["video_idx"], queries["times"], preliminary=10 ** 9,
                              model_types=["model", "chance"])
    got = {k: {kk: float(vv) for kk, vv in v.items()} for k, v in m.items()}
    assert got == meta["metrics_corpus"]


def test_single_video_metrics_identical_to_reference(golden):
    z, meta = golden("tiny_eval")
    videos, queries = _queries(meta)
    bank = _bank(z)
    q = torch.from_numpy(z["query_emb"]).to(DEV)
    prior = synth.make_prior(sorted(set(v["num_segments"] for v in videos)))
    random.seed(123)
    m = vsingle.evaluate_embedded(bank, q, queries["video_idx"], queries["times"],
                                  ["model", "chance", "prior"], prior)
    got = {k: {kk: float(vv) for kk, vv in v.items()} for k, v in m.items()}
    assert got == meta["metrics_single"]


def test_single_video_protocol_vs_oracle_val_shape(golden):
    z, meta = golden("val_eval")
    videos, queries = _queries(meta)
    bank = _bank(z)
    q = torch.from_numpy(z["query_emb"]).to(DEV)
    prior = synth.make_prior(sorted(set(v["num_segments"] for v in videos)))
    random.seed(7)
    got = vsingle.evaluate_embedded(bank, q, queries["video_idx"], queries["times"], ["model", "chance", "prior"], prior)
    off = z["vid_off"]
    vemb = [z["video_emb"][off[i]:off[i + 1]] for i in range(len(off) - 1)]
    random.seed(7)
    want = orc.evaluate_single(vemb, z["query_emb"], queries["video_idx"], queries["times"], prior,
                               model_types=("model", "chance", "prior"), py_random=random)
    assert got == want


def _torch_topk(full, k):
    # ascending by (score, id): stable sort of the scores
    vals, idx = torch.sort(full, dim=1, stable=True)
    return vals[:, :k], idx[:, :k]


@pytest.mark.parametrize("k", [1, 10, 100, 128])
def test_topk_equals_sorted_full_scores(golden, k):
    z, meta = golden("val_eval")
    bank = _bank(z)
    q = torch.from_numpy(z["query_emb"]).to(DEV)
    full = ops.score_full(bank, q)
    ws, wi = _torch_topk(full, k)
    for n_split in (0, 1, 7):
        gs, gi = ops.score_topk(bank, q, k, n_split=n_split)
        assert torch.equal(gs, ws)
        assert torch.equal(gi, wi)


def test_topk_ragged_small_bank_and_k_larger_than_bank(golden):
    z, meta = golden("tiny_eval")
    bank = _bank(z)
    q = torch.from_numpy(z["query_emb"][:5]).to(DEV)
    full = ops.score_full(bank, q)
    gs, gi = ops.score_topk(bank, q, 100)
    ws, wi = _torch_topk(full, 100)
    assert torch.equal(gs, ws) and torch.equal(gi, wi)
    # bank with fewer moments than k: the tail is (+inf, -1)
    small = ops.Bank(torch.from_numpy(z["video_emb"][:int(z["vid_off"][2])]).to(DEV), z["vid_off"][:3])
    gs, gi = ops.score_topk(small, q, 64)
    m = small.m_total
    assert m < 64 and bool(torch.isinf(gs[:, m:]).all()) and bool((gi[:, m:] == -1).all())
    ws, wi = _torch_topk(ops.score_full(small, q), m)
    assert torch.equal(gs[:, :m], ws) and torch.equal(gi[:, :m], wi)


def test_sharded_topk_merge_equals_global():
    V, S, D, Q, k = 6000, 6, 100, 300, 100
    clips = torch.from_numpy(synth.make_bank(5, V, S, D)).to(DEV)
    q = torch.from_numpy(synth.make_query_embeddings(5, Q, D)).to(DEV)
    vid_off = np.arange(V + 1) * S
    gs, gi = ops.score_topk(ops.Bank(clips, vid_off), q, k)
    P = 4
    parts_s, parts_i = [], []
    for r in range(P):
        v0, v1 = r * V // P, (r + 1) * V // P
        shard = ops.Bank(clips[v0 * S:v1 * S], vid_off[v0:v1 + 1] - v0 * S)
        s, i = ops.score_topk(shard, q, k, id_base=v0 * 21)
        parts_s.append(s)
        parts_i.append(i)
    ms, mi = ops.topk_merge(torch.stack(parts_s), torch.stack(parts_i))
    assert torch.equal(ms, gs) and torch.equal(mi, gi)
    assert bool((ms[:, 1:] >= ms[:, :-1]).all())


@pytest.mark.parametrize("P,k,n_valid", [(8, 100, 100), (8, 100, 7), (3, 37, 20), (2, 1, 1), (5, 128, 0)])
def test_topk_merge_sorted_lists_padding_ties_and_unsorted_input(P, k, n_valid):
    """K7 takes the no-sort path when every list arrives sorted (what the shards send: ascending (score, id), (inf, -1)
    padding last) and sorts otherwise; both must equal a lexicographic (score, id) sort of all valid entries.  Scores are
    drawn from a few values so that ties are broken by id."""
    g = torch.Generator(device=DEV).manual_seed(P * 1000 + k)
    Q = 257
    scores = torch.randint(0, 12, (P, Q, k), device=DEV, generator=g).float() * 0.25
    ids = torch.randperm(P * Q * k, device=DEV, generator=g).reshape(P, Q, k).to(torch.int64)
    n_ok = torch.randint(0, n_valid + 1, (P, Q, 1), device=DEV, generator=g) if n_valid else torch.zeros((P, Q, 1), device=DEV, dtype=torch.int64)
    pad = torch.arange(k, device=DEV).view(1, 1, k) >= n_ok
    scores[pad] = float("inf")
    ids[pad] = -1
    # sort every list by (score, id): padding (inf, -1) must go last, so order by a key that maps -1 to +big
    key_id = torch.where(ids < 0, torch.full_like(ids, 2 ** 62), ids)
    order = torch.argsort(key_id, dim=2, stable=True)
    scores, ids, key_id = scores.gather(2, order), ids.gather(2, order), key_id.gather(2, order)
    order = torch.argsort(scores, dim=2, stable=True)
    scores, ids = scores.gather(2, order).contiguous(), ids.gather(2, order).contiguous()

    def reference(sc, idd):
        fs = sc.permute(1, 0, 2).reshape(Q, P * k)
        fi = idd.permute(1, 0, 2).reshape(Q, P * k)
        ki = torch.where(fi < 0, torch.full_like(fi, 2 ** 62), fi)
        o = torch.argsort(ki, dim=1, stable=True)
        fs, fi = fs.gather(1, o), fi.gather(1, o)
        o = torch.argsort(fs, dim=1, stable=True)
        return fs.gather(1, o)[:, :k], fi.gather(1, o)[:, :k]
    ws, wi = reference(scores, ids)
    ms, mi = ops.topk_merge(scores, ids)
    assert torch.equal(ms, ws) and torch.equal(mi, wi)
    perm = torch.randperm(k, device=DEV, generator=g)                      # unsorted lists: the sorting path
    ms, mi = ops.topk_merge(scores[:, :, perm].contiguous(), ids[:, :, perm].contiguous())
    assert torch.equal(ms, ws) and torch.equal(mi, wi)


def test_scores_vs_oracle_random_dims():
    # dims that are not a multiple of the 20-wide k chunk, ragged clip counts up to 32
    rng = np.random.default_rng(9)
    for D in (7, 100, 130):
        nseg = rng.integers(1, 33, size=37)
        vid_off = np.concatenate([[0], np.cumsum(nseg)])
        clips = rng.standard_normal((int(vid_off[-1]), D), dtype=np.float32)
        qs = rng.standard_normal((9, D), dtype=np.float32)
        got = ops.score_full(ops.Bank(torch.from_numpy(clips).to(DEV), vid_off), torch.from_numpy(qs).to(DEV))
        want = orc.score_matrix(clips, vid_off, qs).numpy()
        np.testing.assert_allclose(got.cpu().numpy(), want, rtol=SCORE_RTOL)


def test_near_duplicate_embeddings_keep_fp32_accuracy():
    # the GEMM expansion cancels catastrophically here; the direct form must not
    rng = np.random.default_rng(3)
    D = 100
    qs = rng.standard_normal((4, D), dtype=np.float32)
    clips = np.repeat(qs, 6, axis=0) + 1e-3 * rng.standard_normal((24, D), dtype=np.float32)
    vid_off = np.arange(5) * 6
    got = ops.score_full(ops.Bank(torch.from_numpy(clips).to(DEV), vid_off), torch.from_numpy(qs).to(DEV))
    want = orc.score_matrix(clips, vid_off, qs).numpy()
    np.testing.assert_allclose(got.cpu().numpy(), want, rtol=SCORE_RTOL)


def test_corpus_scale_properties():
    # size-independent properties at a shape the oracle cannot reach: 128k videos x 2k queries
    V, S, D, Q, k = 131072, 6, 100, 2048, 100
    clips = torch.from_numpy(synth.make_bank(1, V, S, D)).to(DEV)
    q = torch.from_numpy(synth.make_query_embeddings(1, Q, D)).to(DEV)
    bank = ops.Bank(clips, np.arange(V + 1) * S)
    s, i = ops.score_topk(bank, q, k)
    assert bool((s[:, 1:] >= s[:, :-1]).all())                      # sorted
    assert bool((i >= 0).all()) and bool((i < V * 21).all())
    assert all(len(set(row.tolist())) == k for row in i[:64].cpu())  # no duplicates
    # count(score < kth) is exactly k-1 minus ties below; count(score < best) == 0
    lt, _ = ops.score_count(bank, q, torch.stack([s[:, 0], s[:, k - 1]], dim=1), torch.zeros(Q, dtype=torch.int32, device=DEV))
    assert bool((lt[:, 0] == 0).all())
    assert bool((lt[:, 1] <= k - 1).all()) and bool((lt[:, 1] >= k - 8).all())
    # recompute the winners' scores with the oracle on a subsample
    ii = i[:8, :5].cpu().numpy()
    for qi in range(8):
        for j in range(5):
            v, m = divmod(int(ii[qi, j]), 21)
            se = orc.generate_moments(S)[m]
            want = orc.moment_scores(clips[v * S:(v + 1) * S].cpu().numpy(), q[qi].cpu().numpy(), [se])[0].item()
            assert abs(want - s[qi, j].item()) <= SCORE_RTOL * want


@pytest.mark.parametrize("D", [100, 1024])
def test_bf16_embeddings_stay_within_stated_tolerance(D):
    """BASELINE config 3: embeddings stored in bf16 (joint space of 100 or 1024 dimensions) change the fp32
    scores by less than the stated 1e-2 relative tolerance (observed ~1e-3), and the fp32 path itself matches
    the oracle at 1e-5 at D = 1024 too.  The retriever falls back to the exact engine above 125 dimensions."""
    rng = np.random.default_rng(17)
    nseg = rng.choice([5, 6], size=400)
    vid_off = np.concatenate([[0], np.cumsum(nseg)])
    clips = (rng.standard_normal((int(vid_off[-1]), D)) / np.sqrt(D)).astype(np.float32)
    qs = (rng.standard_normal((64, D)) / np.sqrt(D)).astype(np.float32)
    full = ops.score_full(ops.Bank(torch.from_numpy(clips).to(DEV), vid_off), torch.from_numpy(qs).to(DEV))
    want = orc.score_matrix(clips, vid_off, qs).numpy()
    np.testing.assert_allclose(full.cpu().numpy(), want, rtol=SCORE_RTOL)
    cb = torch.from_numpy(clips).to(DEV).bfloat16().float()
    qb = torch.from_numpy(qs).to(DEV).bfloat16().float()
    low = ops.score_full(ops.Bank(cb, vid_off), qb)
    rel = ((low - full).abs() / full).max().item()
    assert rel < 1e-2, rel
    # rankings: the bf16 top-10 of every query lies inside the fp32 top-30
    top_low = torch.topk(low, 10, dim=1, largest=False).indices
    top_ref = torch.topk(full, 30, dim=1, largest=False).indices
    inside = (top_low.unsqueeze(2) == top_ref.unsqueeze(1)).any(dim=2).float().mean().item()
    assert inside > 0.9, inside
