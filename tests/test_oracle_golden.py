"""Pin the CPU oracle (oracle/cal_oracle.py) against outputs of the UNMODIFIED reference
(tests/golden/*.npz, produced by oracle/gen_golden.py).  CPU only."""
import random

import numpy as np
import pytest
import torch

import vfr_b200  # noqa: F401
from vfr_b200 import synth
from oracle import cal_oracle as orc


def _inputs(meta):
    videos = synth.make_videos(meta["seed"], meta["n_videos"], meta["feat_dim"],
                               tuple(meta["seg_choices"]), tuple(meta["seg_probs"]))
    queries = synth.make_queries(meta["seed"], videos, meta["n_queries"], meta["vocab"])
    sd = synth.make_state_dict(meta["seed"], meta["feat_dim"], meta["vocab"], spread=meta["spread"])
    return videos, queries, sd


def _split(z):
    off = z["vid_off"]
    return [z["video_emb"][off[i]:off[i + 1]] for i in range(len(off) - 1)]


@pytest.mark.parametrize("case", ["tiny_eval", "long_eval"])
def test_embeddings_match_reference(golden, case):
    z, meta = golden(case)
    videos, queries, sd = _inputs(meta)
    feats = np.concatenate([synth.clip_features(v) for v in videos])
    vemb = orc.visual_embed(sd, feats).numpy()
    np.testing.assert_allclose(vemb, z["video_emb"], rtol=2e-5, atol=2e-6)
    qemb = orc.text_embed(sd, queries["tokens"]).numpy()
    np.testing.assert_allclose(qemb, z["query_emb"], rtol=2e-5, atol=2e-6)


@pytest.mark.parametrize("nl", [0, 1])
def test_text_embed_matches_reference(golden, nl):
    z, meta = golden(f"text_nl{nl}")
    sd = synth.make_state_dict(meta["seed"], meta["feat_dim"], meta["vocab"], normalize_lang=bool(nl))
    q = synth.make_queries(meta["seed"], synth.make_videos(meta["seed"], 4, meta["feat_dim"]),
                           meta["n_queries"], meta["vocab"])
    emb = orc.text_embed(sd, q["tokens"], normalize_lang=bool(nl)).numpy()
    np.testing.assert_allclose(emb, z["emb_batch1"], rtol=2e-5, atol=2e-6)
    np.testing.assert_allclose(emb, z["emb_batched"], rtol=2e-5, atol=2e-6)


@pytest.mark.parametrize("case", ["tiny_eval", "long_eval", "val_eval"])
def test_scores_match_reference(golden, case):
    z, meta = golden(case)
    vemb = _split(z)
    n_keep = z["scores"].shape[0]
    full = orc.score_matrix(z["video_emb"], z["vid_off"], z["query_emb"][:n_keep]).numpy()
    np.testing.assert_allclose(full, z["scores"], rtol=1e-6, atol=0)
    # loop form == the reference's op sequence: bit-exact
    for q in range(min(n_keep, 3)):
        loop = []
        for v in vemb:
            loop.extend(orc.moment_scores_loop(v, z["query_emb"][q], orc.generate_moments(len(v))))
        assert np.array_equal(np.asarray(loop, dtype=np.float32), z["scores"][q])


@pytest.mark.parametrize("case", ["tiny_eval", "long_eval"])
def test_corpus_metrics_match_reference(golden, case):
    z, meta = golden(case)
    _, queries, _ = _inputs(meta)
    np.random.seed(123)
    metrics, det = orc.evaluate_corpus(_split(z), z["query_emb"], queries["video_idx"], queries["times"],
                                       model_types=("model", "chance"), return_details=True)
    assert {k: {kk: float(vv) for kk, vv in v.items()} for k, v in metrics.items()} == meta["metrics_corpus"]
    assert det["rank"][0.5] == z["rank_05"].tolist()
    assert det["rank"][0.7] == z["rank_07"].tolist()


def test_single_metrics_match_reference(golden):
    z, meta = golden("tiny_eval")
    videos, queries, _ = _inputs(meta)
    prior = synth.make_prior(sorted(set(v["num_segments"] for v in videos)))
    random.seed(123)
    metrics = orc.evaluate_single(_split(z), z["query_emb"], queries["video_idx"], queries["times"], prior,
                                  model_types=("model", "chance", "prior"), py_random=random)
    assert {k: {kk: float(vv) for kk, vv in v.items()} for k, v in metrics.items()} == meta["metrics_single"]


def test_pooling_matches_reference(golden):
    z, meta = golden("pool")
    for i, (seed, nf) in enumerate(zip(meta["frame_seeds"], meta["n_frames"])):
        fr = synth.make_frames(seed, nf, meta["feat_dim"])
        for pooling in ("avg", "max"):
            seg, ctx, n = orc.segment_pool(fr, pooling)
            assert np.array_equal(seg, z[f"{pooling}_v{i}_seg"])
            assert np.array_equal(ctx, z[f"{pooling}_v{i}_ctx"])
    for i, (seed, nf) in enumerate(zip(meta["h5_seeds"], meta["h5_n_frames"])):
        fr = synth.make_frames(seed, nf, meta["feat_dim"])
        seg, ctx, n = orc.segment_pool_h5(fr)
        assert np.array_equal(seg, z[f"h5_h{i}_seg"])
        assert np.array_equal(ctx, z[f"h5_h{i}_ctx"])


@pytest.mark.parametrize("norm", [0, 1])
def test_ranking_loss_matches_reference(golden, norm):
    z, meta = golden("train_step")
    embs = [torch.from_numpy(z[f"emb_{k}"]).requires_grad_(True) for k in ("posit", "intra", "inter", "lang")]
    loss, n = orc.ranking_loss(*embs, z["maskp"], z["maskn"], b=meta["b"], lamb=meta["lamb"],
                               normalize=bool(norm))
    loss.backward()
    assert n == int(z[f"norm{norm}_n"])
    np.testing.assert_allclose(loss.item(), float(z[f"norm{norm}_loss"]), rtol=2e-6)
    for k, t in zip(("posit", "intra", "inter", "lang"), embs):
        np.testing.assert_allclose(t.grad.numpy(), z[f"norm{norm}_grad_{k}"], rtol=1e-4, atol=1e-6)


def test_train_batch_embeddings_match_reference(golden):
    z, meta = golden("train_step")
    sd = synth.make_state_dict(meta["seed"], meta["feat_dim"], meta["vocab"], spread=meta["spread"])
    for k in ("posit", "intra", "inter"):
        np.testing.assert_allclose(orc.visual_embed(sd, z[k]).numpy(), z[f"emb_{k}"], rtol=2e-5, atol=2e-6)
    np.testing.assert_allclose(orc.text_embed(sd, z["lang"]).numpy(), z["emb_lang"], rtol=2e-5, atol=2e-6)


def test_moments_and_iou_tables():
    from vfr_b200 import utils
    for n in (0, 1, 5, 6, 12, 30):
        assert utils.generate_moments(n) == orc.generate_moments(n)
    rng = np.random.default_rng(0)
    for n in (5, 6, 30):
        for _ in range(20):
            times = [sorted(rng.integers(0, n, size=2).tolist()) for _ in range(4)]
            for (s, e) in orc.generate_moments(n):
                assert np.array_equal(utils.get_iou(times, s, e), orc.get_iou(times, s, e))
                inter, union = utils.iou_int(times, s, e)
                for thr, inc in ((0.5, False), (0.7, False), (0.3, True), (1.0, True), (0.0, True)):
                    tab = utils.threshold_table(thr, inc)
                    ref = (orc.get_iou(times, s, e) >= thr) if inc else (orc.get_iou(times, s, e) > thr)
                    assert np.array_equal(tab[inter, union].astype(bool), ref)


# ---- goldens of a TRAINED reference model (informative R@k), validate_epoch and whole training steps ------------
def _trained_inputs(meta):
    videos = synth.make_videos(meta["seed"], meta["n_videos"], meta["feat_dim"], tuple(meta["seg_choices"]),
                               tuple(meta["seg_probs"]))
    queries = synth.make_queries(meta.get("query_seed", meta["seed"]), videos, meta["n_queries"], meta["vocab"])
    return videos, queries


@pytest.mark.parametrize("case", ["val_trained", "mid_trained"])
def test_trained_goldens_are_informative_and_match_the_oracle(golden, case):
    z, meta = golden(case)
    videos, queries = _trained_inputs(meta)
    m = meta["metrics_corpus"]["model, IoU=0.5"]
    assert m["R@1"] > 0 and m["R@10"] > m["R@1"] and m["R@100"] > m["R@10"]          # not the degenerate all-zero case
    # ranks of the first positive from the vectorised restatement == the reference's own (ties aside: none here)
    full = orc.score_matrix(z["video_emb"], z["vid_off"], z["query_emb"]).numpy()
    mom_off = np.concatenate([[0], np.cumsum([n * (n + 1) // 2 for n in np.diff(z["vid_off"])])])
    for thr, key in ((0.5, "rank_05"), (0.7, "rank_07")):
        ranks = []
        for q in range(full.shape[0]):
            vi = int(queries["video_idx"][q])
            gt = orc.gt_bits(queries["times"][q], orc.generate_moments(videos[vi]["num_segments"]), thr)
            own = full[q, mom_off[vi]:mom_off[vi + 1]]
            tau = own[gt == 1].min()
            ranks.append(int((full[q] < tau).sum()))
        ranks = np.asarray(ranks)
        ref = z[key]
        # equal wherever no other score ties with tau at fp32 resolution (the reference's argsort is unstable there)
        assert (ranks == ref).mean() > 0.97 and np.abs(ranks - ref).max() <= 3
        rec = {f"R@{k}": float(np.mean(ref < k) * 100) for k in (1, 10, 100)}
        rec["MR"] = float(np.median(ref))
        assert rec == meta["metrics_corpus"][f"model, IoU={thr}"]


def test_mid_trained_full_protocols_match_reference(golden):
    z, meta = golden("mid_trained")
    videos, queries = _trained_inputs(meta)
    vemb = _split(z)
    np.random.seed(123)
    metrics = orc.evaluate_corpus(vemb, z["query_emb"], queries["video_idx"], queries["times"], model_types=("model", "chance"))
    assert {k: {kk: float(vv) for kk, vv in v.items()} for k, v in metrics.items()} == meta["metrics_corpus"]
    prior = synth.make_prior(sorted(set(v["num_segments"] for v in videos)))
    random.seed(123)
    single = orc.evaluate_single(vemb, z["query_emb"], queries["video_idx"], queries["times"], prior,
                                 model_types=("model", "chance", "prior"), py_random=random)
    assert {k: {kk: float(vv) for kk, vv in v.items()} for k, v in single.items()} == meta["metrics_single"]
    # the embeddings themselves, from the committed trained weights
    sd = {k[2:]: z[k] for k in z.files if k.startswith("w:")}
    feats = np.concatenate([synth.clip_features(v) for v in videos])
    np.testing.assert_allclose(orc.visual_embed(sd, feats).numpy(), z["video_emb"], rtol=2e-5, atol=2e-6)
    np.testing.assert_allclose(orc.text_embed(sd, queries["tokens"]).numpy(), z["query_emb"], rtol=2e-5, atol=2e-6)


@pytest.mark.parametrize("tag,size", [("size250", 250), ("size100", 100), ("all", -1)])
def test_validate_epoch_matches_reference(golden, tag, size):
    z, meta = golden("mid_trained")
    videos, queries = _trained_inputs(meta)
    scalars, pr = orc.validate_epoch(_split(z), z["query_emb"], queries["video_idx"], queries["times"], size=size)
    ref = meta["validate"][tag]
    ref_scalars = {name: vals for name, vals, _ in ref["scalars"]}
    assert scalars == ref_scalars
    assert {k: {str(kk): [float(x) for x in vv] for kk, vv in v.items()} for k, v in pr.items()} == ref["pr_curve"]
    assert all(step == 7 for _, _, step in ref["scalars"])        # written at the trainer's global_step


@pytest.mark.parametrize("norm", [0, 1])
def test_whole_training_steps_match_reference(golden, norm):
    """Four forwards + ranking loss + backward + Adam(lr 5e-4, wd 5e-3), dropout off: loss, mean grad norm and the
    weights after every step against what the reference's Trainer.train_epoch produced."""
    z, meta = golden("train_full_step")
    sd = {k: torch.from_numpy(v) for k, v in
          synth.make_state_dict(meta["seed"], meta["feat_dim"], meta["vocab"], hidden=meta["hidden"], spread=meta["spread"]).items()}
    state = {}
    logged = z[f"norm{norm}_logged"].reshape(-1, 2)
    for step in range(meta["steps"]):
        batch = {k: z[f"b{step}_{k}"] for k in ("posit", "intra", "inter", "lang", "maskp", "maskn")}
        loss, n, gnorm = orc.train_step(sd, batch, state, step + 1, normalize_loss=bool(norm), lr=meta["lr"],
                                        weight_decay=meta["weight_decay"])
        np.testing.assert_allclose(loss / n, logged[step, 0], rtol=2e-5)
        np.testing.assert_allclose(gnorm, logged[step, 1], rtol=2e-4)
        np.testing.assert_allclose(loss / n, float(z[f"norm{norm}_s{step}_epoch_loss"]), rtol=2e-5)
        for k in z.files:
            pre = f"norm{norm}_s{step}_w:"
            if k.startswith(pre):
                np.testing.assert_allclose(sd[k[len(pre):]].numpy(), z[k], rtol=2e-4, atol=2e-6, err_msg=k)
