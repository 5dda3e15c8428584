"""GPU tests of the split-bf16 tcgen05 GEMM and of the tensor-core K3 built on it."""
import numpy as np
import pytest
import torch

import vfr_b200  # noqa: F401
from vfr_b200 import models, ops, synth

pytestmark = pytest.mark.gpu
DEV = "cuda"
# stated tolerance of the tensor-core (split-bf16) query-embedding engine against the fp32 reference,
# relative to the largest embedding component: 20 chained recurrent GEMMs, each ~5e-6 (see below).
# (The north star allows 1e-2 for bf16 embeddings; the exact-fp32 engine stays at 1e-5.)
TC_TEXT_TOL = 3e-5


@pytest.mark.parametrize("m,n,k", [(1, 5, 3), (257, 257, 33), (300, 501, 777), (1000, 100, 2000), (513, 500, 8194)])
def test_linear_tc_fp32_accuracy(m, n, k):
    g = torch.Generator(device=DEV).manual_seed(m * 7 + n)
    x = torch.randn(m, k, device=DEV, generator=g)
    w = torch.randn(n, k, device=DEV, generator=g) * 0.05
    b = torch.randn(n, device=DEV, generator=g)
    want = x.double() @ w.double().t() + b.double()
    wp = ops.pack_weight_tc(w)
    got = ops.linear_tc(x, wp, n, b)
    scale = want.abs().max().item()
    err = (got.double() - want).abs().max().item()
    # split-bf16 keeps 16 mantissa bits per operand (~2e-6); on top of that the TMEM accumulator truncates
    # instead of rounding, which adds a bias that grows with the number of accumulated k-steps (3K/16)
    tol = 1e-5 if k <= 2048 else 5e-5
    assert err <= tol * scale, (err, scale)
    # the exact-fp32 CUDA-core GEMM agrees with it to the same tolerance
    ref32 = ops.linear(x, w, b)
    assert (ref32 - got).abs().max().item() <= tol * scale
    got_relu = ops.linear_tc(x, wp, n, b, relu=True)
    assert torch.equal(got_relu, got.clamp_min(0))


def _model(sd, feat_dim, normalize_lang=False):
    m = models.CALModel(visual_input_dim=2 * feat_dim + 2, pretrained_emb=torch.from_numpy(sd["word_embedding.weight"]),
                        normalize_lang=normalize_lang)
    m.load_state_dict({k: torch.from_numpy(v) for k, v in sd.items()})
    return m.to(DEV).eval()


def test_text_embed_tc_matches_reference(golden):
    z, meta = golden("tiny_eval")
    videos = synth.make_videos(meta["seed"], meta["n_videos"], meta["feat_dim"], tuple(meta["seg_choices"]), tuple(meta["seg_probs"]))
    queries = synth.make_queries(meta["seed"], videos, meta["n_queries"], meta["vocab"])
    sd = synth.make_state_dict(meta["seed"], meta["feat_dim"], meta["vocab"], spread=meta["spread"])
    model = _model(sd, meta["feat_dim"])
    tok = torch.from_numpy(queries["tokens"]).to(DEV)
    with torch.no_grad():
        exact = model(tok, False, DEV)
        model.engine = "tc"
        got = model(tok, False, DEV)
    scale = np.abs(z["query_emb"]).max()
    assert np.abs(got.cpu().numpy() - z["query_emb"]).max() <= TC_TEXT_TOL * scale
    assert (got - exact).abs().max().item() <= TC_TEXT_TOL * scale


@pytest.mark.parametrize("nl", [0, 1])
def test_text_embed_tc_normalize_lang_and_bad_tokens(golden, nl):
    z, meta = golden(f"text_nl{nl}")
    sd = synth.make_state_dict(meta["seed"], meta["feat_dim"], meta["vocab"], normalize_lang=bool(nl))
    q = synth.make_queries(meta["seed"], synth.make_videos(meta["seed"], 4, meta["feat_dim"]), meta["n_queries"], meta["vocab"])
    model = _model(sd, meta["feat_dim"], bool(nl))
    model.engine = "tc"
    with torch.no_grad():
        emb = model(torch.from_numpy(q["tokens"]).to(DEV), False, DEV)
    scale = np.abs(z["emb_batch1"]).max()
    assert np.abs(emb.cpu().numpy() - z["emb_batch1"]).max() <= TC_TEXT_TOL * scale
    bad = torch.from_numpy(q["tokens"]).clone()
    bad[2, 1] = -3
    with pytest.raises(IndexError):
        model(bad.to(DEV), False, DEV)


def test_text_embed_tc_large_batch_vs_exact():
    sd = synth.make_state_dict(5, 8, 3000)
    model = _model(sd, 8)
    rng = np.random.default_rng(0)
    tok = torch.from_numpy(rng.integers(1, 3000, size=(700, 20))).to(DEV)    # not a multiple of the 256-row tile
    with torch.no_grad():
        exact = model(tok, False, DEV)
        model.engine = "tc"
        got = model(tok, False, DEV)
    assert (got - exact).abs().max().item() <= TC_TEXT_TOL * exact.abs().max().item()


@pytest.mark.parametrize("nl", [False, True])
def test_text_embed_tc_padding_aware_rows(nl):
    """The tensor-core K3 orders rows by length and lets a query join the backward recurrence at its first
    real token with the state an all-padding row has reached (reference: padding is fed through the LSTM,
    models.py:65).  Every length 0..20, zeros INSIDE a query, a non-zero embedding for the padding id and
    batch composition must not change a query's embedding."""
    sd = synth.make_state_dict(7, 8, 600, normalize_lang=nl)
    sd["word_embedding.weight"][0] = 0.3 * np.random.default_rng(1).standard_normal(100).astype(np.float32)
    model = _model(sd, 8, nl)
    rng = np.random.default_rng(2)
    tok = np.zeros((300, 20), dtype=np.int64)
    for i in range(300):
        n = i % 21                                       # every length, including 0 (all padding) and 20
        tok[i, :n] = rng.integers(1, 600, size=n)
    tok[40:60, 2] = 0                                    # padding id in the middle of a query
    t = torch.from_numpy(tok).to(DEV)
    with torch.no_grad():
        exact = model(t, False, DEV)
        model.engine = "tc"
        got = model(t, False, DEV)
        alone = torch.cat([model(t[i:i + 1], False, DEV) for i in (0, 1, 20, 21, 45, 299)])
        shuffled = model(t.flip(0), False, DEV).flip(0)
    assert (got - exact).abs().max().item() <= TC_TEXT_TOL * exact.abs().max().item()
    assert torch.equal(alone, got[[0, 1, 20, 21, 45, 299]])
    assert torch.equal(shuffled, got)


@pytest.mark.parametrize("n_queries", [1, 257, 700, 4737])
def test_text_embed_tc_overlapped_step_launches_equal_serial_ones(n_queries, monkeypatch):
    """The 20 step GEMMs of a batch of up to 20 000 rows overlap through programmatic dependent launch; a tile waits for
    the row blocks of the previous step it reads through per-row-block counters (vfr_gemm_tc2.cuh, G2Deps).  The
    embeddings must carry the same bits as with serial launches (VFR_K3_OVERLAP=0), every repetition."""
    sd = synth.make_state_dict(11, 8, 900)
    model = _model(sd, 8)
    model.engine = "tc"
    rng = np.random.default_rng(n_queries)
    tok = np.zeros((n_queries, 20), dtype=np.int64)
    lens = rng.integers(0, 21, size=n_queries)
    for r in range(n_queries):
        tok[r, :lens[r]] = rng.integers(1, 900, size=lens[r])
    t = torch.from_numpy(tok).to(DEV)
    with torch.no_grad():
        monkeypatch.setenv("VFR_K3_OVERLAP", "0")
        serial = model(t, False, DEV).clone()
        monkeypatch.setenv("VFR_K3_OVERLAP", "1")
        for _ in range(3):
            assert torch.equal(model(t, False, DEV).view(torch.int32), serial.view(torch.int32))


def test_visual_embed_tc_matches_reference_and_exact_path(golden):
    """K2 on tensor cores (engine "tc"): the reference's own visual embeddings within 1e-5 of their scale."""
    z, meta = golden("tiny_eval")
    sd = synth.make_state_dict(meta["seed"], meta["feat_dim"], meta["vocab"], spread=meta["spread"])
    model = _model(sd, meta["feat_dim"])
    rng = np.random.default_rng(0)
    x = torch.from_numpy(rng.random((777, 2 * meta["feat_dim"] + 2), dtype=np.float32)).to(DEV)
    with torch.no_grad():
        assert model.visual_engine == "tc"                     # the default K2 is the tcgen05 split-fp16 path
        got = model(x)
        model.visual_engine = "exact"
        exact = model(x)
        model.visual_engine = "tc_bf16x3"
        got3 = model(x)
    assert (got - exact).abs().max().item() <= 1e-5 * exact.abs().max().item()
    assert (got3 - exact).abs().max().item() <= 1e-5 * exact.abs().max().item()
    # the reference's full-size rows: [L2-normalised segment | L2-normalised context | tef] (data.py:174-213)
    big = CALModelBig()
    seg = rng.random((1000, 4096), dtype=np.float32)
    ctx = rng.random((1000, 4096), dtype=np.float32)
    seg /= np.linalg.norm(seg, axis=1, keepdims=True) + 1e-5
    ctx /= np.linalg.norm(ctx, axis=1, keepdims=True) + 1e-5
    xb = torch.from_numpy(np.concatenate([seg, ctx, rng.random((1000, 2), dtype=np.float32)], axis=1)).to(DEV)
    with torch.no_grad():
        g2 = big(xb)                                            # default: tcgen05 split-fp16
        big.visual_engine = "exact"
        e2 = big(xb)
        big.visual_engine = "tc_bf16x3"
        g3 = big(xb)
        lin1, lin2 = big.visual_fc[0], big.visual_fc[2]
        want = torch.relu(xb.double() @ lin1.weight.double().t() + lin1.bias.double()) @ lin2.weight.double().t() + lin2.bias.double()
        # the split-weight form on the same data: seg / ctx / CSR offsets in, no 8194-wide rows
        vid_off = np.arange(0, 1001, 5)
        ctx_v = ctx[::5]
        xs = np.concatenate([seg, np.repeat(ctx_v, 5, axis=0), np.tile(np.stack([np.arange(5) / np.float32(5), (np.arange(5) + 1) / np.float32(5)], 1).astype(np.float32), (200, 1))], axis=1)
        big.visual_engine = "tc"
        g_split = big.embed_clips(torch.from_numpy(seg).to(DEV), torch.from_numpy(ctx_v).to(DEV), vid_off)
        xs_t = torch.from_numpy(xs).to(DEV)
        want_split = torch.relu(xs_t.double() @ lin1.weight.double().t() + lin1.bias.double()) @ lin2.weight.double().t() + lin2.bias.double()
    scale = want.abs().max().item()
    # fp32 bar for the exact path AND for the default tensor-core path (22-bit split-fp16 operands); the round-1 split-bf16
    # variant adds its 2^-16 errors coherently over 8194 all-positive products and is held to 5e-5
    assert (e2.double() - want).abs().max().item() <= 1e-5 * scale
    assert (g2.double() - want).abs().max().item() <= 1e-5 * scale
    assert (g3.double() - want).abs().max().item() <= 5e-5 * scale
    assert (g_split.double() - want_split).abs().max().item() <= 1e-5 * want_split.abs().max().item()


def CALModelBig():
    torch.manual_seed(123)
    table = torch.randn(50, 100) * 0.4
    return models.CALModel(visual_input_dim=8194, pretrained_emb=table).to(DEV).eval()


def test_visual_embed_tc_row_scales_and_unaligned_rows():
    """K2's operand exponent is per ROW (found in the split pass): rows whose magnitudes differ by 1e6, an all-zero row
    and a row count that is not a multiple of the 256-row tile each keep fp32 accuracy relative to THEIR OWN output;
    the assembled [N, 8194] rows are only 8-byte aligned (8194 floats per row: the scalar branch of the split kernel),
    the split-weight inputs 16-byte aligned (the vector branch)."""
    big = CALModelBig()
    rng = np.random.default_rng(5)
    n_vid, n = 77, 5
    seg = rng.random((n_vid * n, 4096), dtype=np.float32)
    ctx = rng.random((n_vid, 4096), dtype=np.float32)
    row_scale = (10.0 ** rng.uniform(-3, 3, size=(n_vid * n, 1))).astype(np.float32)
    seg *= row_scale
    seg[7] = 0.0
    ctx *= (10.0 ** rng.uniform(-3, 3, size=(n_vid, 1))).astype(np.float32)
    ctx[3] = 0.0
    vid_off = np.arange(n_vid + 1) * n
    tef = np.tile(np.stack([np.arange(n) / np.float32(n), (np.arange(n) + 1) / np.float32(n)], 1).astype(np.float32), (n_vid, 1))
    x = torch.from_numpy(np.concatenate([seg, np.repeat(ctx, n, axis=0), tef], axis=1)).to(DEV)
    lin1, lin2 = big.visual_fc[0], big.visual_fc[2]
    with torch.no_grad():
        hid = torch.relu(x.double() @ lin1.weight.double().t() + lin1.bias.double())
        want = hid @ lin2.weight.double().t() + lin2.bias.double()
        got_rows = big(x)
        got_split = big.embed_clips(torch.from_numpy(seg).to(DEV), torch.from_numpy(ctx).to(DEV), vid_off)
    # per-row bar: 1e-5 of what the row's own hidden activations can contribute to an output
    row_bar = 1e-5 * (hid.abs() @ lin2.weight.double().abs().t()).max(dim=1, keepdim=True).values.clamp_min(1e-30)
    assert bool(((got_rows.double() - want).abs() <= row_bar).all())
    assert bool(((got_split.double() - want).abs() <= row_bar).all())
