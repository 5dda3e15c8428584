"""GPU parity tests of the tensor-core scoring path (tcgen05 split-bf16 GEMM + fused epilogue)
against the reference's own scores (goldens), the CPU oracle and the exact-fp32 CUDA path."""
import numpy as np
import pytest
import torch

import vfr_b200  # noqa: F401
from vfr_b200 import ops, synth
from oracle import cal_oracle as orc

pytestmark = pytest.mark.gpu
DEV = "cuda"
SCORE_RTOL = 1e-5


def _bank(z):
    return ops.Bank(torch.from_numpy(z["video_emb"]).to(DEV), z["vid_off"])


@pytest.mark.parametrize("case", ["tiny_eval", "val_eval"])
def test_tc_scores_match_reference(golden, case):
    z, meta = golden(case)
    bank = _bank(z)                     # ragged 5/6-clip videos: padded slots are masked
    n_keep = z["scores"].shape[0]
    q = torch.from_numpy(z["query_emb"][:n_keep]).to(DEV)
    got = ops.score_full_tc(bank, q).cpu().numpy()
    assert not np.isnan(got).any()
    rel = np.abs(got - z["scores"]) / z["scores"]
    assert rel.max() < SCORE_RTOL, rel.max()


def test_tc_scores_vs_exact_path_many_queries(golden):
    z, meta = golden("val_eval")
    bank = _bank(z)
    q = torch.from_numpy(z["query_emb"]).to(DEV)       # 48 queries: a ragged last query tile
    exact = ops.score_full(bank, q)
    got = ops.score_full_tc(bank, q)
    rel = ((got - exact).abs() / exact).max().item()
    assert rel < SCORE_RTOL, rel


def test_tc_near_duplicates_use_exact_fallback():
    rng = np.random.default_rng(3)
    D = 100
    qs = rng.standard_normal((4, D), dtype=np.float32)
    clips = np.repeat(qs, 6, axis=0) + 1e-3 * rng.standard_normal((24, D), dtype=np.float32)
    vid_off = np.arange(5) * 6
    got = ops.score_full_tc(ops.Bank(torch.from_numpy(clips).to(DEV), vid_off), torch.from_numpy(qs).to(DEV))
    want = orc.score_matrix(clips, vid_off, qs).numpy()
    np.testing.assert_allclose(got.cpu().numpy(), want, rtol=SCORE_RTOL)


def _check_topk(bank, q, k, n_split=0, n_terms=3, rtol=SCORE_RTOL):
    exact = ops.score_full(bank, q)
    ws, wi = torch.sort(exact, dim=1, stable=True)
    gs, gi = ops.score_topk_tc(bank, q, k, n_split=n_split, n_terms=n_terms)
    kk = min(k, bank.m_total)
    assert bool((gs[:, 1:kk] >= gs[:, :kk - 1]).all())
    # sorted score lists agree to the score tolerance
    assert ((gs[:, :kk] - ws[:, :kk]).abs() / ws[:, :kk]).max().item() < 2 * rtol
    # every returned moment is a legitimate member of the top-k up to the tolerance, no duplicates
    exact_of_ids = torch.gather(exact, 1, gi[:, :kk].clamp_min(0))
    assert bool((gi[:, :kk] >= 0).all())
    assert bool((exact_of_ids <= ws[:, kk - 1:kk] * (1 + 4 * rtol)).all())
    assert ((exact_of_ids - gs[:, :kk]).abs() / exact_of_ids).max().item() < 2 * rtol
    for row in gi[:16, :kk].cpu().tolist():
        assert len(set(row)) == kk
    if kk < k:
        assert bool(torch.isinf(gs[:, kk:]).all()) and bool((gi[:, kk:] == -1).all())
    agree = (gi[:, :kk] == wi[:, :kk]).float().mean().item()
    return agree


@pytest.mark.parametrize("k", [1, 10, 100])
def test_tc_topk_vs_exact(golden, k):
    z, meta = golden("val_eval")
    bank = _bank(z)
    q = torch.from_numpy(z["query_emb"]).to(DEV)
    for n_split in (0, 1, 5):
        assert _check_topk(bank, q, k, n_split) > 0.97


def test_tc_topk_small_bank(golden):
    z, meta = golden("tiny_eval")
    bank = _bank(z)
    q = torch.from_numpy(z["query_emb"]).to(DEV)
    _check_topk(bank, q, 100)
    _check_topk(bank, q, 128)


def test_tc_plain_bf16_mode_tolerance(golden):
    # BASELINE config 3: bf16 embeddings within a stated 1e-2 tolerance of the fp32 scores
    z, meta = golden("val_eval")
    bank = _bank(z)
    q = torch.from_numpy(z["query_emb"][:8]).to(DEV)
    exact = ops.score_full(bank, q)
    got = ops.score_full_tc(bank, q, n_terms=1)
    assert ((got - exact).abs() / exact).max().item() < 1e-2
    _check_topk(bank, q, 10, n_terms=1, rtol=1e-2)


def test_tc_corpus_scale_properties_and_sharding():
    V, S, D, Q, k = 131072, 6, 100, 1024, 100
    clips = torch.from_numpy(synth.make_bank(1, V, S, D)).to(DEV)
    q = torch.from_numpy(synth.make_query_embeddings(1, Q, D)).to(DEV)
    vid_off = np.arange(V + 1) * S
    bank = ops.Bank(clips, vid_off)
    s, i = ops.score_topk_tc(bank, q, k)
    es, ei = ops.score_topk(bank, q, k)                  # exact CUDA-core path
    assert ((s - es).abs() / es).max().item() < 2 * SCORE_RTOL
    assert (i == ei).float().mean().item() > 0.97
    parts_s, parts_i = [], []
    for r in range(2):
        v0, v1 = r * V // 2, (r + 1) * V // 2
        shard = ops.Bank(clips[v0 * S:v1 * S], vid_off[v0:v1 + 1] - v0 * S)
        ps, pi = ops.score_topk_tc(shard, q, k, id_base=v0 * 21)
        parts_s.append(ps)
        parts_i.append(pi)
    ms, mi = ops.topk_merge(torch.stack(parts_s), torch.stack(parts_i))
    assert torch.equal(ms, s) and torch.equal(mi, i)    # same arithmetic per pair, whatever the sharding
