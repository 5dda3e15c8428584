"""GPU parity tests of K1 (pooling), K2 (visual embedding), K3 (query embedding), K6 (ranking loss)
and of the end-to-end drop-in entry points, against golden outputs of the unmodified reference."""
import random

import numpy as np
import pytest
import torch
from torch.utils.data import DataLoader

import vfr_b200  # noqa: F401
from vfr_b200 import data as vdata
from vfr_b200 import evaluate as vev
from vfr_b200 import evaluate_single as vsingle
from vfr_b200 import main as vmain
from vfr_b200 import models, ops, synth
from oracle import cal_oracle as orc

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _inputs(meta):
    videos = synth.make_videos(meta["seed"], meta["n_videos"], meta["feat_dim"],
                               tuple(meta["seg_choices"]), tuple(meta["seg_probs"]))
    queries = synth.make_queries(meta["seed"], videos, meta["n_queries"], meta["vocab"])
    sd = synth.make_state_dict(meta["seed"], meta["feat_dim"], meta["vocab"], spread=meta["spread"])
    return videos, queries, sd


def _model(sd, feat_dim, normalize_lang=False):
    m = models.CALModel(visual_input_dim=2 * feat_dim + 2, pretrained_emb=torch.from_numpy(sd["word_embedding.weight"]),
                        normalize_lang=normalize_lang)
    m.load_state_dict({k: torch.from_numpy(v) for k, v in sd.items()})
    return m.to(DEV).eval()


def _close(got, want, tol):
    got, want = np.asarray(got, dtype=np.float64), np.asarray(want, dtype=np.float64)
    scale = np.abs(want).max()
    assert np.abs(got - want).max() <= tol * scale, (np.abs(got - want).max(), scale)


@pytest.mark.parametrize("case", ["tiny_eval", "long_eval", "val_eval"])
def test_visual_and_text_embeddings_match_reference(golden, case):
    z, meta = golden(case)
    videos, queries, sd = _inputs(meta)
    model = _model(sd, meta["feat_dim"])
    feats = torch.from_numpy(np.concatenate([synth.clip_features(v) for v in videos])).to(DEV)
    with torch.no_grad():
        vemb = model(feats)
        qemb = model(torch.from_numpy(queries["tokens"]).to(DEV), False, DEV)
    _close(vemb.cpu().numpy(), z["video_emb"], 1e-5)
    _close(qemb.cpu().numpy(), z["query_emb"], 1e-5)
    # and the scores computed from OUR embeddings stay within the north-star tolerance
    n_keep = z["scores"].shape[0]
    got = ops.score_full(ops.Bank(vemb, z["vid_off"]), qemb[:n_keep]).cpu().numpy()
    assert (np.abs(got - z["scores"]) / z["scores"]).max() < 1e-5


@pytest.mark.parametrize("nl", [0, 1])
def test_text_embed_normalize_lang_variants(golden, nl):
    z, meta = golden(f"text_nl{nl}")
    sd = synth.make_state_dict(meta["seed"], meta["feat_dim"], meta["vocab"], normalize_lang=bool(nl))
    q = synth.make_queries(meta["seed"], synth.make_videos(meta["seed"], 4, meta["feat_dim"]), meta["n_queries"], meta["vocab"])
    model = _model(sd, meta["feat_dim"], bool(nl))
    with torch.no_grad():
        emb = model(torch.from_numpy(q["tokens"]).to(DEV), False, DEV)
        one = model(torch.from_numpy(q["tokens"][:1]).to(DEV), False, DEV)   # the reference's batch-1 call
    _close(emb.cpu().numpy(), z["emb_batch1"], 1e-5)
    _close(one.cpu().numpy(), z["emb_batch1"][:1], 1e-5)
    with pytest.raises(IndexError):
        bad = torch.from_numpy(q["tokens"]).clone()
        bad[0, 0] = meta["vocab"] + 5
        model(bad.to(DEV), False, DEV)


def test_segment_pooling_matches_reference(golden):
    z, meta = golden("pool")
    frames = [synth.make_frames(s, nf, meta["feat_dim"]) for s, nf in zip(meta["frame_seeds"], meta["n_frames"])]
    for pooling in ("avg", "max"):
        pooled = vdata.pool_videos(frames, pooling, device=DEV)
        for i, feats in enumerate(pooled):
            want_seg, want_ctx = z[f"{pooling}_v{i}_seg"], z[f"{pooling}_v{i}_ctx"]
            assert feats["num_segments"] == want_seg.shape[0]
            assert feats["segment_features"].dtype == np.float64 and feats["context_features"].dtype == np.float32
            np.testing.assert_allclose(feats["segment_features"], want_seg, rtol=2e-6, atol=1e-9)
            np.testing.assert_allclose(feats["context_features"], want_ctx, rtol=2e-6, atol=1e-9)
    h5 = [synth.make_frames(s, nf, meta["feat_dim"]) for s, nf in zip(meta["h5_seeds"], meta["h5_n_frames"])]
    pooled = vdata.pool_videos(h5, "avg", preprocessed=True, device=DEV)
    for i, feats in enumerate(pooled):
        np.testing.assert_allclose(feats["segment_features"], z[f"h5_h{i}_seg"], rtol=2e-6, atol=1e-9)
        np.testing.assert_allclose(feats["context_features"], z[f"h5_h{i}_ctx"], rtol=2e-6, atol=1e-9)


def test_segment_pooling_full_size_vs_oracle():
    # DiDeMo-sized video: 150 frames x 4096 fc7 features, plus a ragged one and a 5-segment h5 clip
    frames = [synth.make_frames(1, 150, 4096), synth.make_frames(2, 137, 4096), synth.make_frames(3, 120, 4096)]
    for pooling in ("avg", "max"):
        for feats, fr in zip(vdata.pool_videos(frames, pooling, device=DEV), frames):
            seg, ctx, n = orc.segment_pool(fr, pooling)
            assert feats["num_segments"] == n
            np.testing.assert_allclose(feats["segment_features"], seg, rtol=2e-6, atol=1e-9)
            np.testing.assert_allclose(feats["context_features"], ctx, rtol=2e-6, atol=1e-9)
    for feats, fr in zip(vdata.pool_videos(frames, preprocessed=True, device=DEV), frames):
        seg, ctx, n = orc.segment_pool_h5(fr)
        assert feats["num_segments"] == n == (5 if len(fr) <= 125 else 6)
        np.testing.assert_allclose(feats["segment_features"], seg, rtol=2e-6, atol=1e-9)
        np.testing.assert_allclose(feats["context_features"], ctx, rtol=2e-6, atol=1e-9)


@pytest.mark.parametrize("norm", [0, 1])
def test_ranking_loss_forward_backward_match_reference(golden, norm):
    z, meta = golden("train_step")
    embs = [torch.from_numpy(z[f"emb_{k}"]).to(DEV).requires_grad_(True) for k in ("posit", "intra", "inter", "lang")]
    tr = vmain.Trainer(device=DEV, normalize_loss=bool(norm), b=meta["b"], lamb=meta["lamb"])
    loss, n = tr.ranking_loss(*embs, torch.from_numpy(z["maskp"]).to(DEV), torch.from_numpy(z["maskn"]).to(DEV))
    assert n == int(z[f"norm{norm}_n"])
    np.testing.assert_allclose(loss.item(), float(z[f"norm{norm}_loss"]), rtol=2e-6)
    (2.0 * loss).backward()
    for k, t in zip(("posit", "intra", "inter", "lang"), embs):
        want = 2.0 * z[f"norm{norm}_grad_{k}"]
        _close(t.grad.cpu().numpy(), want, 2e-5)


def _dataset(videos, queries, validate=True):
    ds = vdata.CustomDataset.__new__(vdata.CustomDataset)
    ds.validate = validate
    ds.video_features = {v["name"]: v for v in videos}
    ds.num_segments_info = {v["name"]: v["num_segments"] for v in videos}
    ds.lang_features = {a: torch.from_numpy(queries["tokens"][i:i + 1]) for i, a in enumerate(queries["annot_id"])}
    annotations = {a: dict(video=videos[int(queries["video_idx"][i])]["name"], description="", times=queries["times"][i])
                   for i, a in enumerate(queries["annot_id"])}
    return ds, annotations


def _iters(ds, videos, annotations, max_seg=6):
    vit = DataLoader(ds, collate_fn=vdata.validate_collate,
                     batch_sampler=vdata.VideoBatchSampler([v["name"] for v in videos], ds.num_segments_info))
    lit = DataLoader(ds, collate_fn=vdata.validate_collate,
                     batch_sampler=vdata.LanguageBatchSampler(annotations, ds.num_segments_info, max_seg))
    return vit, lit


def test_evaluate_drop_in_end_to_end(golden, capsys):
    """evaluate.evaluate / evaluate_single.evaluate with the reference's signatures, iterators in,
    metric dicts out - identical to what the unmodified reference returned on the same inputs."""
    z, meta = golden("tiny_eval")
    videos, queries, sd = _inputs(meta)
    model = _model(sd, meta["feat_dim"])
    ds, annotations = _dataset(videos, queries)
    vit, lit = _iters(ds, videos, annotations)
    torch.random.manual_seed(123); random.seed(123); np.random.seed(123)
    m = vev.evaluate(model, vit, lit, annotations, DEV, preliminary=10, model_types=["model", "chance"])
    assert {k: {kk: float(vv) for kk, vv in v.items()} for k, v in m.items()} == meta["metrics_corpus"]
    assert "model, IoU=0.5:" in capsys.readouterr().out
    prior = synth.make_prior(sorted(set(v["num_segments"] for v in videos)))
    random.seed(123)
    m = vsingle.evaluate(model, vit, lit, annotations, DEV, ["model", "chance", "prior"], prior)
    assert {k: {kk: float(vv) for kk, vv in v.items()} for k, v in m.items()} == meta["metrics_single"]
    with pytest.raises(IndexError):   # reference quirk: prior[] is indexed unconditionally
        vsingle.evaluate(model, vit, lit, annotations, DEV)


def test_long_video_drop_in(golden):
    z, meta = golden("long_eval")
    videos, queries, sd = _inputs(meta)
    model = _model(sd, meta["feat_dim"])
    ds, annotations = _dataset(videos, queries)
    vit, lit = _iters(ds, videos, annotations, max_seg=30)
    m = vev.evaluate(model, vit, lit, annotations, DEV, preliminary=0)
    want = {k: v for k, v in meta["metrics_corpus"].items() if k.startswith("model")}
    got = {k: {kk: float(vv) for kk, vv in v.items()} for k, v in m.items()}
    for k in want:   # 465 overlapping moments per video: allow the tie range on MR only
        for kk in ("R@1", "R@10", "R@100"):
            assert abs(got[k][kk] - want[k][kk]) <= 100.0 / meta["n_queries"] + 1e-9
        assert abs(got[k]["MR"] - want[k]["MR"]) <= 2


def test_training_step_gradients_match_torch_reference(golden):
    z, meta = golden("train_step")
    sd = synth.make_state_dict(meta["seed"], meta["feat_dim"], meta["vocab"], spread=meta["spread"])
    model = _model(sd, meta["feat_dim"])            # eval(): dropout off, as the golden was produced
    for p in model.parameters():
        p.grad = None
    tr = vmain.Trainer(device=DEV)
    batch = {k: torch.from_numpy(z[k]) for k in ("posit", "intra", "inter", "lang", "maskp", "maskn")}
    loss, n = tr.ranking_loss(*tr._embed_batch(model, batch))
    np.testing.assert_allclose(loss.item(), float(z["norm0_loss"]), rtol=1e-5)
    loss.backward()
    # oracle: the same step in plain torch fp32 on the CPU
    cpu = {k: torch.from_numpy(v.copy()).requires_grad_(k != "word_embedding.weight") for k, v in sd.items()}
    embs = [orc.visual_embed(cpu, batch[k]) for k in ("posit", "intra", "inter")]
    lang = orc.text_embed(cpu, batch["lang"])
    ref_loss, _ = orc.ranking_loss(*embs, lang, batch["maskp"], batch["maskn"])
    ref_loss.backward()
    for name, p in model.named_parameters():
        if not p.requires_grad:
            continue
        assert p.grad is not None, name
        _close(p.grad.cpu().numpy(), cpu[name].grad.numpy(), 2e-4)


def test_train_and_test_epoch_run(golden):
    z, meta = golden("train_step")
    sd = synth.make_state_dict(meta["seed"], meta["feat_dim"], meta["vocab"], spread=meta["spread"])
    model = _model(sd, meta["feat_dim"])
    batch = {k: torch.from_numpy(z[k]) for k in ("posit", "intra", "inter", "lang", "maskp", "maskn")}
    tr = vmain.Trainer(device=DEV, compute_grads=True)
    opt = torch.optim.Adam(model.parameters(), lr=5e-4, weight_decay=5e-3)
    before = tr.test_epoch(model, [batch])
    np.testing.assert_allclose(before, float(z["norm0_loss"]) / int(z["norm0_n"]), rtol=1e-5)
    torch.manual_seed(0)
    for _ in range(3):
        tr.train_epoch(model, [batch], opt)
    after = tr.test_epoch(model, [batch])
    assert after < before and tr.global_step == 3


def test_validate_epoch_matches_oracle(golden):
    z, meta = golden("tiny_eval")
    videos, queries, sd = _inputs(meta)
    model = _model(sd, meta["feat_dim"])
    ds, annotations = _dataset(videos, queries)
    vit, lit = _iters(ds, videos, annotations)
    tr = vmain.Trainer(device=DEV)
    pr = tr.validate_epoch(model, vit, lit, annotations, size=-1)
    # oracle: full scores + argsort with the '>=' rule (model/main.py:148-183)
    off = z["vid_off"]
    full = orc.score_matrix(z["video_emb"], off, z["query_emb"]).numpy()
    mom_off = np.concatenate([[0], np.cumsum([len(orc.generate_moments(off[i + 1] - off[i])) for i in range(len(off) - 1)])])
    thr_range = [i / 10 for i in range(11)]
    tp = {t: {k: 0 for k in (1, 10, 100)} for t in thr_range}
    rel = {t: 0 for t in thr_range}
    ranks = {0.5: [], 0.7: []}
    for q in range(meta["n_queries"]):
        v = int(queries["video_idx"][q])
        order = np.argsort(full[q], kind="stable")
        for t in thr_range:
            gt = np.zeros(full.shape[1], dtype=int)
            gt[mom_off[v]:mom_off[v + 1]] = orc.gt_bits(queries["times"][q], orc.generate_moments(off[v + 1] - off[v]), t, True)
            pred = gt[order]
            rel[t] += pred.sum()
            if t in ranks:
                ranks[t].append(int(np.where(pred == 1)[0][0]) + 1)
            for k in tp[t]:
                tp[t][k] += pred[:k].sum()
    for k in (1, 10, 100):
        np.testing.assert_allclose(pr["precision"][k], [tp[t][k] / (k * meta["n_queries"]) for t in thr_range])
        np.testing.assert_allclose(pr["recall"][k], [tp[t][k] / rel[t] for t in thr_range])
    assert tr.last_validation["median_rank"][0.5] == ranks[0.5]
    assert tr.last_validation["median_rank"][0.7] == ranks[0.7]


# ---- a TRAINED reference model (informative R@k): the drop-in protocols against what the reference returned ----------
def _trained(golden, case):
    z, meta = golden(case)
    videos = synth.make_videos(meta["seed"], meta["n_videos"], meta["feat_dim"], tuple(meta["seg_choices"]),
                               tuple(meta["seg_probs"]))
    queries = synth.make_queries(meta.get("query_seed", meta["seed"]), videos, meta["n_queries"], meta["vocab"])
    return z, meta, videos, queries


def _trained_model(z, meta):
    sd = {k[2:]: torch.from_numpy(z[k]) for k in z.files if k.startswith("w:")}
    m = models.CALModel(visual_input_dim=2 * meta["feat_dim"] + 2, pretrained_emb=sd["word_embedding.weight"],
                        hidden_size=meta["hidden"])
    m.load_state_dict(sd)
    return m.to(DEV).eval()


def _plain(m):
    return {k: {kk: float(vv) for kk, vv in v.items()} for k, v in m.items()}


def test_trained_model_evaluate_drop_in_matches_reference(golden):
    """evaluate.evaluate / evaluate_single.evaluate on a briefly trained model (R@1 = 21.9, R@10 = 60.6, R@100 = 91.9 in
    the reference's own run): iterators in, metric dicts out, equal to the reference's dicts."""
    z, meta, videos, queries = _trained(golden, "mid_trained")
    assert meta["metrics_corpus"]["model, IoU=0.5"]["R@1"] > 0
    model = _trained_model(z, meta)
    ds, annotations = _dataset(videos, queries)
    vit, lit = _iters(ds, videos, annotations)
    torch.random.manual_seed(123); random.seed(123); np.random.seed(123)
    m = vev.evaluate(model, vit, lit, annotations, DEV, preliminary=0, model_types=["model", "chance"])
    assert _plain(m) == meta["metrics_corpus"]
    prior = synth.make_prior(sorted(set(v["num_segments"] for v in videos)))
    random.seed(123)
    m = vsingle.evaluate(model, vit, lit, annotations, DEV, ["model", "chance", "prior"], prior)
    assert _plain(m) == meta["metrics_single"]
    # stage by stage: embeddings and the rank of the first positive of every query
    feats = torch.from_numpy(np.concatenate([synth.clip_features(v) for v in videos])).to(DEV)
    with torch.no_grad():
        vemb = model(feats)
        qemb = model(torch.from_numpy(queries["tokens"]).to(DEV), False, DEV)
    _close(vemb.cpu().numpy(), z["video_emb"], 1e-5)
    _close(qemb.cpu().numpy(), z["query_emb"], 1e-5)


@pytest.mark.parametrize("tag,size", [("size250", 250), ("size100", 100), ("all", -1)])
def test_validate_epoch_matches_reference_golden(golden, tag, size):
    """Trainer.validate_epoch against the reference's own validate_epoch (model/main.py:121-212) run by
    oracle/gen_golden.py with recording writers: the CustomRecall / MedianRank / MeanReciprocalRank scalar groups and
    the returned precision / recall curves."""
    from oracle.ref_harness import NullWriter
    z, meta, videos, queries = _trained(golden, "mid_trained")
    model = _trained_model(z, meta)
    ds, annotations = _dataset(videos, queries)
    vit, lit = _iters(ds, videos, annotations)
    w = NullWriter()
    tr = vmain.Trainer(val_writer=w, device=DEV)
    tr.global_step = 7
    pr = tr.validate_epoch(model, vit, lit, annotations, size=size)
    ref = meta["validate"][tag]
    assert [[n, v, s] for n, v, s in w.scalar_groups] == ref["scalars"]
    assert {k: {str(kk): [float(x) for x in vv] for kk, vv in v.items()} for k, v in pr.items()} == ref["pr_curve"]


def test_val_shape_trained_embeddings_metrics_match_reference(golden):
    """The val shape (1,094 videos, 21.9 k moments, 192 queries) with the embeddings of a trained reference model:
    R@1 / R@10 / R@100 / MR dicts equal to the reference's, per-query ranks equal off ties."""
    z, meta, videos, queries = _trained(golden, "val_trained")
    want = meta["metrics_corpus"]
    assert want["model, IoU=0.5"]["R@1"] > 0 and want["model, IoU=0.5"]["R@100"] > 30
    bank = ops.Bank(torch.from_numpy(z["video_emb"]).to(DEV), z["vid_off"])
    q_emb = torch.from_numpy(z["query_emb"]).to(DEV)
    np.random.seed(123)
    m = vev.evaluate_embedded(bank, q_emb, queries["video_idx"], queries["times"], preliminary=0,
                              model_types=("model", "chance"), verbose=False)
    assert _plain(m) == want
    res = vev.rank_first_positive(bank, q_emb, queries["video_idx"], queries["times"], [0.5, 0.7])
    for ti, key in enumerate(("rank_05", "rank_07")):
        ours, ref = res["rank"][:, ti], z[key]
        assert (ours == ref).mean() > 0.97 and np.abs(ours - ref).max() <= 3      # the reference's argsort is unstable on ties
    prior = synth.make_prior(sorted(set(v["num_segments"] for v in videos)))
    random.seed(123)
    m = vsingle.evaluate_embedded(bank, q_emb, queries["video_idx"], queries["times"], ("model", "chance", "prior"), prior)
    assert _plain(m) == meta["metrics_single"]


def test_feature_files_stream_through_pinned_staging_into_k1(golden, tmp_path, monkeypatch):
    """Feature ingestion: the dataset's real constructor path on real files -
    ``features_vgg19/vgg19_ft_<video>.npy`` as get_rgb_features.py writes them (data.py:164-166) and MCN's
    ``fc7_subsample5_fps25_<video>.h5`` (data.py:145-148, through a stand-in h5py module: h5py is not installed) -
    memory-mapped, staged in pinned buffers small enough to force several chunks and a video larger than a chunk, pooled
    by K1; the result equals what the reference's own load_video_features produced from the same files (golden)."""
    import sys
    import types
    z, meta = golden("pool")
    monkeypatch.setitem(vdata.FEATURE_DIM, "vgg19", meta["feat_dim"])
    (tmp_path / "features_vgg19").mkdir()
    names = []
    for i, (seed, nf) in enumerate(zip(meta["frame_seeds"], meta["n_frames"])):
        np.save(tmp_path / "features_vgg19" / f"vgg19_ft_v{i}.npy", synth.make_frames(seed, nf, meta["feat_dim"]))
        names.append(f"v{i}")
    real_pool = vdata.pool_video_files
    monkeypatch.setattr(vdata, "pool_video_files", lambda *a, **k: real_pool(*a, chunk_frames=140, **k))   # 151 > 140 frames
    for pooling in ("avg", "max"):
        ds = vdata.CustomDataset(names, {}, str(tmp_path), "vgg19", pooling=pooling, device=DEV)
        for i, n in enumerate(names):
            got = ds.video_features[n]
            want_seg, want_ctx = z[f"{pooling}_v{i}_seg"], z[f"{pooling}_v{i}_ctx"]
            assert got["num_segments"] == want_seg.shape[0] == ds.num_segments_info[n]
            assert got["segment_features"].dtype == np.float64
            _close(got["segment_features"], want_seg, 2e-6)
            _close(got["context_features"], want_ctx, 2e-6)
    store = {}
    for i, (seed, nf) in enumerate(zip(meta["h5_seeds"], meta["h5_n_frames"])):
        store[f"fc7_subsample5_fps25_h{i}.h5"] = synth.make_frames(seed, nf, meta["feat_dim"])

    class FakeFile:
        def __init__(self, path):
            self.key = str(path).split("/")[-1]
        def __enter__(self):
            return self
        def __exit__(self, *a):
            return False
        def __getitem__(self, k):
            return store[self.key]
    fake = types.ModuleType("h5py")
    fake.File = FakeFile
    monkeypatch.setitem(sys.modules, "h5py", fake)
    h5names = [f"h{i}" for i in range(len(meta["h5_seeds"]))]
    ds = vdata.CustomDataset(h5names, {}, str(tmp_path), "vgg19", prep=True, device=DEV)
    for i, n in enumerate(h5names):
        _close(ds.video_features[n]["segment_features"], z[f"h5_h{i}_seg"], 2e-6)
        _close(ds.video_features[n]["context_features"], z[f"h5_h{i}_ctx"], 2e-6)


# ---- whole training steps on the hand-written kernels (forward with saved activations, backward, fused Adam) ------------
@pytest.mark.parametrize("norm", [0, 1])
def test_training_steps_match_reference_train_epoch(golden, norm):
    """Trainer.train_epoch with our CALModel (dropout_rate = 0, as the golden was produced) and FusedAdam(lr 5e-4,
    wd 5e-3) on the batches the reference's sampler produced: the logged loss / n and mean gradient norm of every step
    and the weights after the first and the last step equal what the reference's Trainer + torch.optim.Adam produced."""
    from oracle.ref_harness import NullWriter
    z, meta = golden("train_full_step")
    sd = synth.make_state_dict(meta["seed"], meta["feat_dim"], meta["vocab"], hidden=meta["hidden"], spread=meta["spread"])
    model = models.CALModel(visual_input_dim=2 * meta["feat_dim"] + 2, pretrained_emb=torch.from_numpy(sd["word_embedding.weight"]),
                            hidden_size=meta["hidden"], dropout_rate=0.0)
    model.load_state_dict({k: torch.from_numpy(v) for k, v in sd.items()})
    model = model.to(DEV)
    w = NullWriter()
    tr = vmain.Trainer(train_writer=w, device=DEV, compute_grads=True, normalize_loss=bool(norm))
    opt = vmain.FusedAdam(model.parameters(), lr=meta["lr"], weight_decay=meta["weight_decay"])
    logged = z[f"norm{norm}_logged"].reshape(-1, 2)
    for step in range(meta["steps"]):
        batch = {k: torch.from_numpy(z[f"b{step}_{k}"]) for k in ("posit", "intra", "inter", "lang", "maskp", "maskn")}
        epoch_loss = tr.train_epoch(model, [batch], opt)
        np.testing.assert_allclose(epoch_loss, float(z[f"norm{norm}_s{step}_epoch_loss"]), rtol=2e-5)
        got = dict((n, v) for n, v, s in w.scalars if s == step)
        np.testing.assert_allclose(got["loss"], logged[step, 0], rtol=2e-5)
        np.testing.assert_allclose(got["grad_norm"], logged[step, 1], rtol=2e-4)
        state = model.state_dict()
        for k in z.files:
            pre = f"norm{norm}_s{step}_w:"
            if k.startswith(pre):
                # Adam's first steps move every weight by ~lr whatever the size of its gradient (m / sqrt(v) ~ +-1), so a
                # gradient of ~1e-9 that differs in its last bits moves a weight visibly: the bar is a tenth of one
                # update for the worst element and a thousandth on average
                err = np.abs(state[k[len(pre):]].cpu().numpy().astype(np.float64) - z[k])
                assert err.max() <= 0.1 * meta["lr"] and err.mean() <= 1e-3 * meta["lr"], (k, err.max(), err.mean())
    assert tr.global_step == meta["steps"]


def test_fused_adam_equals_torch_adam():
    torch.manual_seed(5)
    shapes = [(500, 514), (500,), (100, 500), (100,), (512, 100), (512, 128), (3,)] * 5           # 35 tensors: two launches
    ours = [torch.nn.Parameter(torch.randn(s, device=DEV)) for s in shapes]
    ref = [torch.nn.Parameter(p.detach().clone()) for p in ours]
    a = vmain.FusedAdam(ours, lr=5e-4, weight_decay=5e-3)
    b = torch.optim.Adam(ref, lr=5e-4, weight_decay=5e-3)
    for it in range(4):
        for p, q in zip(ours, ref):
            g = torch.randn_like(p) * (10.0 ** (it - 2))
            p.grad, q.grad = g.clone(), g.clone()
        a.step()
        b.step()
    for p, q in zip(ours, ref):
        assert (p - q).abs().max().item() <= 2e-6 * max(1.0, q.abs().max().item())
    assert a.state[ours[0]]["step"] == 4


def test_text_backward_with_learnable_length_matches_torch_autograd(golden):
    """normalize_lang=True: gradients of the LSTM, lang_fc AND the learnable word length (models.py:36-38,62-64) from the
    hand-written BPTT against torch autograd over the CPU restatement."""
    _, meta = golden("text_nl1")
    sd = synth.make_state_dict(meta["seed"], meta["feat_dim"], meta["vocab"], normalize_lang=True)
    model = _model(sd, meta["feat_dim"], normalize_lang=True).train()
    q = synth.make_queries(meta["seed"], synth.make_videos(meta["seed"], 4, meta["feat_dim"]), meta["n_queries"], meta["vocab"])
    tok = torch.from_numpy(q["tokens"])
    g = torch.Generator().manual_seed(3)
    up = torch.randn(tok.shape[0], 100, generator=g)
    out = model(tok.to(DEV), False, DEV)
    out.backward(up.to(DEV))
    cpu = {k: torch.from_numpy(v.copy()).requires_grad_(k != "word_embedding.weight") for k, v in sd.items()}
    ref = orc.text_embed(cpu, q["tokens"], normalize_lang=True)
    ref.backward(up)
    _close(out.detach().cpu().numpy(), ref.detach().numpy(), 1e-5)
    for name, p in model.named_parameters():
        if name.startswith(("lstm", "lang_fc", "learnable_length")):
            assert p.grad is not None, name
            _close(p.grad.cpu().numpy(), cpu[name].grad.numpy(), 2e-4)
    assert float(model.learnable_length.weight.grad[0].abs().sum()) == 0.0        # padding_idx row


# ---- training batches built on the device (negative sampling + row gather), pooled-moment features ---------------------
@pytest.mark.parametrize("same_length", [True, False])
def test_device_batch_sampler_draws_what_the_reference_sampler_may_draw(golden, same_length):
    """DeviceBatchSampler against the reference's CustomBatchSampler semantics (data.py:275-337): every draw satisfies the
    reference's constraints, the batches are cut by its rule, the gathered rows equal make_visual_features bit for bit,
    the draws cover the admissible choices about uniformly, and a training epoch consumes the stream."""
    _, meta = golden("train_step")
    videos = synth.make_videos(meta["seed"], meta["n_videos"], meta["feat_dim"])
    queries = synth.make_queries(meta["seed"], videos, meta["n_queries"], meta["vocab"])
    sd = synth.make_state_dict(meta["seed"], meta["feat_dim"], meta["vocab"], spread=meta["spread"])
    ds, annotations = _dataset(videos, queries, validate=False)
    if same_length:      # a positive that fills its video has no same-length negative: the reference raises there (H6)
        annotations = {a: v for a, v in annotations.items()
                       if all(t[1] - t[0] + 1 < ds.num_segments_info[v["video"]] for t in v["times"])}
    samp = vdata.DeviceBatchSampler(120, annotations, ds.num_segments_info, ds, same_length=same_length, seed=7, device=DEV)
    names = samp.videos
    seen_q, n_batches = 0, 0
    intra_hist = {}
    for epoch in range(3):
        for batch in samp:
            n_batches += 1
            s = batch["samples"]
            B = s.shape[0]
            assert batch["lang"].shape == (B, 20) and int(batch["maskp"].max()) == B - 1
            rows_p = int((s[:, 3] - s[:, 2] + 1).sum())
            rows_n = int((s[:, 5] - s[:, 4] + 1).sum())
            assert batch["posit"].shape == (rows_p, 2 * meta["feat_dim"] + 2) and batch["intra"].shape[0] == rows_n
            assert batch["inter"].shape[0] == rows_p and batch["maskn"].shape[0] == rows_n
            # the reference's cut rule: the batch closed at the first query that reached batch_size (or is the last one)
            cum_p, cum_n = np.cumsum(s[:, 3] - s[:, 2] + 1), np.cumsum(s[:, 5] - s[:, 4] + 1)
            assert (np.maximum(cum_p, cum_n)[:-1] < 120).all()
            r = 0
            for i in range(B):
                vp, vn, st, en, sn, enn, status = (int(x) for x in s[i, :7])
                info = annotations[samp.annot_ids[int(np.nonzero((samp.lang == batch["lang"][i]).all(dim=1).cpu().numpy())[0][0])]]
                n = ds.num_segments_info[names[vp]]
                assert status == 0 and [st, en] in [list(t) for t in info["times"]]
                assert 0 <= sn <= enn < n and (sn, enn) != (st, en)
                if same_length:
                    assert enn - sn == en - st
                else:
                    assert [sn, enn] not in [list(t) for t in info["times"]]
                assert vn != vp and ds.num_segments_info[names[vn]] >= en + 1
                intra_hist.setdefault((n, st, en), {}).setdefault((sn, enn), 0)
                intra_hist[(n, st, en)][(sn, enn)] += 1
                want = ds.make_visual_features(names[vp], st, en)
                assert torch.equal(batch["posit"][r:r + en - st + 1].cpu(), want)
                assert torch.equal(batch["inter"][r:r + en - st + 1].cpu(), ds.make_visual_features(names[vn], st, en))
                r += en - st + 1
            seen_q += B
    assert seen_q == 3 * len(annotations) and n_batches >= 3
    # every admissible same-length negative of a frequent (video length, positive) class was drawn
    if same_length:
        (n, st, en), hist = max(intra_hist.items(), key=lambda kv: sum(kv[1].values()))
        assert len(hist) == (n - (en - st)) - 1 and min(hist.values()) > 0
    model = _model(sd, meta["feat_dim"]).train()
    tr = vmain.Trainer(device=DEV, compute_grads=False)
    loss = tr.train_epoch(model, samp, vmain.FusedAdam(model.parameters(), lr=5e-4, weight_decay=5e-3))
    assert np.isfinite(loss) and tr.global_step >= 1


def test_device_batch_sampler_raises_like_the_reference_on_a_full_video_positive(golden):
    _, meta = golden("train_step")
    videos = synth.make_videos(meta["seed"], 8, meta["feat_dim"])
    queries = synth.make_queries(meta["seed"], videos, 6, meta["vocab"])
    n0 = videos[int(queries["video_idx"][0])]["num_segments"]
    queries["times"][0] = [[0, n0 - 1]] * 4                               # the positive fills the whole video
    ds, annotations = _dataset(videos, queries, validate=False)
    samp = vdata.DeviceBatchSampler(120, annotations, ds.num_segments_info, ds, same_length=True, device=DEV)
    with pytest.raises(IndexError):
        next(iter(samp))


def test_moment_pool_prefix_sums_match_numpy():
    """Pooled-moment features (MCN-style, non-reference variant): mean segment feature of all n (n + 1) / 2 moments."""
    rng = np.random.default_rng(2)
    nseg = rng.choice([1, 5, 6, 30, 32], size=40)
    vid_off = np.concatenate([[0], np.cumsum(nseg)])
    seg = rng.random((int(vid_off[-1]), 512), dtype=np.float32)
    out, mom_off = ops.moment_pool(torch.from_numpy(seg).to(DEV), vid_off)
    out = out.cpu().numpy()
    assert mom_off[-1] == sum(n * (n + 1) // 2 for n in nseg) == out.shape[0]
    for v in (0, 7, 23, 39):
        n = int(nseg[v])
        for m, (s, e) in enumerate(orc.generate_moments(n)):
            want = seg[vid_off[v] + s:vid_off[v] + e + 1].astype(np.float64).mean(axis=0)
            # (a difference of fp32 prefix sums: absolute error ~ 2^-24 x the prefix, up to n = 32 summands)
            np.testing.assert_allclose(out[mom_off[v] + m], want, rtol=1e-5, atol=4e-6)


def test_pooled_moment_variant_scores_match_a_torch_restatement(golden):
    """The non-reference pooled-moment variant end to end: pooled features -> K2 -> one-row candidates -> distance ranking,
    against the same computation spelled out with plain torch fp64 / fp32 on the CPU."""
    z, meta = golden("tiny_eval")
    videos, queries, sd = _inputs(meta)
    model = _model(sd, meta["feat_dim"])
    nseg = [v["num_segments"] for v in videos]
    vid_off = np.concatenate([[0], np.cumsum(nseg)])
    seg = np.concatenate([v["segment_features"] for v in videos]).astype(np.float32)
    ctx = np.stack([v["context_features"] for v in videos]).astype(np.float32)
    bank, mom_off = vev.pooled_moment_bank(model, torch.from_numpy(seg).to(DEV), torch.from_numpy(ctx).to(DEV), vid_off)
    assert bank.n_videos == int(mom_off[-1]) == sum(n * (n + 1) // 2 for n in nseg)
    q = torch.from_numpy(z["query_emb"][:8]).to(DEV)
    got = ops.score_full(bank, q).cpu().numpy()                       # [8, M]: one "moment" per single-row video
    rows = []
    for v, n in enumerate(nseg):
        for (s, e) in orc.generate_moments(n):
            f = seg[vid_off[v] + s:vid_off[v] + e + 1].astype(np.float64).mean(0).astype(np.float32)
            rows.append(np.concatenate([f, ctx[v], np.asarray([s / n, (e + 1) / n], dtype=np.float32)]))
    emb = orc.visual_embed(sd, np.stack(rows)).numpy()
    want = np.sqrt((((emb[None] - z["query_emb"][:8, None]) + 1e-6) ** 2).sum(-1))
    assert np.abs(got - want).max() <= 2e-5 * want.max()
