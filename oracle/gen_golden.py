"""Generate tests/golden/*.npz by RUNNING THE UNMODIFIED REFERENCE (/root/reference/model) on
seeded synthetic inputs.  Run in the build container only (the GPU box has no /root/reference):

    python oracle/gen_golden.py            # all cases
    python oracle/gen_golden.py tiny_eval  # one case

Recipe: put /root/reference/model on sys.path, stub the two missing imports that
never touch the hot path (h5py - except for the fake ``File`` used to pin the .h5 pooling branch -
and matplotlib.pyplot), build ``data.CustomDataset`` through ``__new__`` with synthetic
``video_features`` / ``lang_features`` / ``num_segments_info``, wrap it with the reference's real
samplers and collates, and call the reference's real entry points.  Inputs are pure functions of
the seeds (vfr_b200.synth), so each file stores OUTPUTS plus the seeds/shapes that regenerate the
inputs, and the library versions the numbers were produced with.
"""
import io
import json
import os
import random
import sys
import tempfile
import time
import types
from contextlib import redirect_stdout

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference/model"
sys.path.insert(0, ROOT)

sys.modules["h5py"] = _h5 = types.ModuleType("h5py")   # a stub even if a real h5py exists: case_pool installs a fake File

import vfr_b200  # noqa: E402,F401
from vfr_b200 import synth  # noqa: E402
from oracle import ref_harness  # noqa: E402

_ref = ref_harness.load(REF)   # the reference itself, straight from /root/reference/model
rdata, reval, rsingle, rmodels, rutils = _ref.data, _ref.evaluate, _ref.evaluate_single, _ref.models, _ref.utils
from torch.utils.data import DataLoader  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")
VERSIONS = dict(torch=torch.__version__, numpy=np.__version__)


def _save(name, meta, **arrays):
    meta = dict(meta, versions=VERSIONS, generated_by="oracle/gen_golden.py", case=name)
    path = os.path.join(OUT, name + ".npz")
    np.savez_compressed(path, meta=np.frombuffer(json.dumps(meta).encode(), dtype=np.uint8), **arrays)
    print(f"wrote {path}: {os.path.getsize(path) / 1e6:.2f} MB")


def _ref_model(sd, feat_dim, normalize_lang=False, dropout_rate=0.3):
    return ref_harness.ref_model(_ref, sd, feat_dim, normalize_lang, dropout_rate)


def _ref_dataset(videos, queries, validate=True):
    return ref_harness.ref_dataset(_ref, videos, queries, validate)


def _ref_iters(ds, videos, annotations, max_seg=None):
    return ref_harness.ref_iters(_ref, ds, videos, annotations, max_seg)


def _metrics_json(m):
    return {k: {kk: float(vv) for kk, vv in v.items()} for k, v in m.items()}


def _eval_case(name, seed, n_videos, n_queries, feat_dim, vocab, spread, seg_choices, seg_probs,
               keep_scores, max_seg=None, run_single=True):
    videos = synth.make_videos(seed, n_videos, feat_dim, seg_choices, seg_probs)
    queries = synth.make_queries(seed, videos, n_queries, vocab)
    sd = synth.make_state_dict(seed, feat_dim, vocab, spread=spread)
    model = _ref_model(sd, feat_dim)
    ds, annotations = _ref_dataset(videos, queries)
    vit, lit, moments = _ref_iters(ds, videos, annotations, max_seg)

    # (1) the real corpus-level evaluate(), 'model' + 'chance' rankers, seeds as evaluate.py:95-97
    torch.random.manual_seed(123); random.seed(123); np.random.seed(123)
    t0 = time.time()
    with redirect_stdout(io.StringIO()):
        m_corpus = reval.evaluate(model, vit, lit, annotations, "cpu", preliminary=10 ** 9,
                                  model_types=["model", "chance"])
    t_corpus = time.time() - t0

    # (2) the real single-video evaluate(), all three rankers, seeds as evaluate_single.py:91-93
    m_single = None
    if run_single:
        prior = synth.make_prior(sorted(set(v["num_segments"] for v in videos)))
        torch.random.manual_seed(123); random.seed(123); np.random.seed(123)
        with redirect_stdout(io.StringIO()):
            m_single = rsingle.evaluate(model, vit, lit, annotations, "cpu",
                                        ["model", "chance", "prior"], prior)

    # (3) per-stage intermediates, produced by the same reference ops the two functions run
    #     (evaluate.py:35,44,53-58,71,77) so the kernels can be checked stage by stage
    with torch.no_grad():
        vemb = [model(b["feature"]) for b in vit]
        qemb = torch.cat([model(b["feature"], False, "cpu") for b in lit], dim=0)
    vid_off = np.cumsum([0] + [int(v.shape[0]) for v in vemb]).astype(np.int64)
    n_keep = min(keep_scores, n_queries)
    scores = []
    ranks = {0.5: [], 0.7: []}
    for q in range(n_queries):
        dist_all = []
        gts = {0.5: [], 0.7: []}
        for vi, ve in enumerate(vemb):
            n = ve.size(0)
            dist = torch.nn.functional.pairwise_distance(ve, qemb[q:q + 1].repeat(n, 1))
            for s, e in moments[n]:
                dist_all.append(dist.index_select(0, torch.arange(s, e + 1)).mean().item())
                for thr in gts:
                    if vi == int(queries["video_idx"][q]):
                        gts[thr].append(int((rutils.get_iou(queries["times"][q], s, e) > thr).sum() >= 2))
                    else:
                        gts[thr].append(0)
        order = np.argsort(dist_all)
        for thr in gts:
            ranks[thr].append(int(np.where(np.array(gts[thr])[order] == 1)[0][0]))
        if q < n_keep:
            scores.append(np.asarray(dist_all, dtype=np.float32))
    meta = dict(seed=seed, n_videos=n_videos, n_queries=n_queries, feat_dim=feat_dim, vocab=vocab,
                spread=spread, seg_choices=list(seg_choices), seg_probs=list(seg_probs),
                max_seg=max_seg, metrics_corpus=_metrics_json(m_corpus),
                metrics_single=_metrics_json(m_single) if m_single else None,
                ref_seconds_corpus=t_corpus, ref_threads=torch.get_num_threads())
    _save(name, meta, video_emb=torch.cat(vemb).numpy(), vid_off=vid_off, query_emb=qemb.numpy(),
          scores=np.stack(scores), rank_05=np.asarray(ranks[0.5]), rank_07=np.asarray(ranks[0.7]))


def case_tiny_eval():
    _eval_case("tiny_eval", seed=11, n_videos=14, n_queries=40, feat_dim=64, vocab=300, spread=4.0,
               seg_choices=(6, 5), seg_probs=(0.6, 0.4), keep_scores=40)


def case_val_eval():
    # DiDeMo val shape (1,094 videos, 4096-d), 48-query subsample (0.41 s / query here)
    _eval_case("val_eval", seed=123, n_videos=1094, n_queries=48, feat_dim=4096, vocab=2000,
               spread=4.0, seg_choices=(6, 5), seg_probs=(0.83, 0.17), keep_scores=4)


def case_long_eval():
    # BASELINE config 4: 30 clips -> 465 candidate moments (table injected)
    _eval_case("long_eval", seed=31, n_videos=10, n_queries=12, feat_dim=64, vocab=300, spread=4.0,
               seg_choices=(30, 12), seg_probs=(0.7, 0.3), keep_scores=12, max_seg=30,
               run_single=False)


def case_text():
    seed, vocab = 41, 500
    for nl in (False, True):
        sd = synth.make_state_dict(seed, 16, vocab, normalize_lang=nl)
        model = _ref_model(sd, 16, normalize_lang=nl)
        videos = synth.make_videos(seed, 4, 16)
        q = synth.make_queries(seed, videos, 24, vocab)
        with torch.no_grad():
            one_by_one = torch.cat([model(torch.from_numpy(q["tokens"][i:i + 1]), False, "cpu")
                                    for i in range(24)])
            batched = model(torch.from_numpy(q["tokens"]), False, "cpu")
        _save(f"text_nl{int(nl)}", dict(seed=seed, vocab=vocab, feat_dim=16, n_queries=24,
                                        normalize_lang=nl),
              emb_batch1=one_by_one.numpy(), emb_batched=batched.numpy())


def case_pool():
    out = {}
    meta = dict(frame_seeds=[], n_frames=[], feat_dim=512)
    with tempfile.TemporaryDirectory() as tmp:
        os.makedirs(os.path.join(tmp, "features_vgg19"))
        names = []
        rdata.FEATURE_DIM["vgg19"] = 512      # module constant (data.py:22-24); smaller fixture
        for i, nf in enumerate([150, 151, 130, 125, 26, 25, 7]):
            fr = synth.make_frames(500 + i, nf, 512)
            np.save(os.path.join(tmp, "features_vgg19", f"vgg19_ft_v{i}.npy"), fr)
            names.append(f"v{i}")
            meta["frame_seeds"].append(500 + i); meta["n_frames"].append(nf)
        for pooling in ("avg", "max"):
            ds = rdata.CustomDataset.__new__(rdata.CustomDataset)
            ds.preprocessed = False; ds.ft_directory = tmp; ds.ft_type = "vgg19"; ds.pooling = pooling
            ds.video_features = {}; ds.num_segments_info = {}
            ds.load_video_features(names)
            for n in names:
                out[f"{pooling}_{n}_seg"] = ds.video_features[n]["segment_features"]
                out[f"{pooling}_{n}_ctx"] = ds.video_features[n]["context_features"]
        # the .h5 branch (data.py:144-161) through a fake h5py.File
        store = {}

        class FakeFile:
            def __init__(self, path):
                self.key = os.path.basename(str(path))
            def __getitem__(self, k):
                return store[self.key]
            def close(self):
                pass
        _h5.File = FakeFile
        h5names = []
        for i, nf in enumerate([150, 125, 160, 140]):
            fr = synth.make_frames(600 + i, nf, 512)
            store[f"fc7_subsample5_fps25_h{i}.h5"] = fr
            h5names.append(f"h{i}")
        meta["h5_seeds"] = [600, 601, 602, 603]; meta["h5_n_frames"] = [150, 125, 160, 140]
        ds = rdata.CustomDataset.__new__(rdata.CustomDataset)
        ds.preprocessed = True; ds.ft_directory = tmp; ds.ft_type = "vgg19"; ds.pooling = "avg"
        ds.video_features = {}; ds.num_segments_info = {}
        ds.load_video_features(h5names)
        for n in h5names:
            out[f"h5_{n}_seg"] = ds.video_features[n]["segment_features"]
            out[f"h5_{n}_ctx"] = ds.video_features[n]["context_features"]
    _save("pool", meta, **out)


def case_train_step():
    import main as rmain
    seed, feat_dim, vocab = 51, 64, 300
    videos = synth.make_videos(seed, 60, feat_dim)
    queries = synth.make_queries(seed, videos, 200, vocab)
    sd = synth.make_state_dict(seed, feat_dim, vocab, spread=4.0)
    model = _ref_model(sd, feat_dim)
    ds, annotations = _ref_dataset(videos, queries, validate=False)
    random.seed(123); np.random.seed(123); torch.random.manual_seed(123)
    it = DataLoader(ds, shuffle=False, collate_fn=rdata.custom_collate,
                    batch_sampler=rdata.CustomBatchSampler(120, annotations, ds.num_segments_info,
                                                           same_length=True))
    batch = next(iter(it))
    out = {k: batch[k].numpy() for k in ("posit", "intra", "inter", "lang", "maskp", "maskn")}
    for norm in (False, True):
        tr = rmain.Trainer(device="cpu", normalize_loss=norm)
        embs = [model(batch[k]).detach().requires_grad_(True) for k in ("posit", "intra", "inter")]
        lang = model(batch["lang"], False, "cpu").detach().requires_grad_(True)
        loss, n = tr.ranking_loss(embs[0], embs[1], embs[2], lang, batch["maskp"], batch["maskn"])
        loss.backward()
        tag = f"norm{int(norm)}"
        out[f"{tag}_loss"] = np.asarray(loss.item(), dtype=np.float32)
        out[f"{tag}_n"] = np.asarray(n)
        for nm, t in zip(("posit", "intra", "inter", "lang"), embs + [lang]):
            out[f"{tag}_grad_{nm}"] = t.grad.numpy()
            if not norm:
                out[f"emb_{nm}"] = t.detach().numpy()
    _save("train_step", dict(seed=seed, feat_dim=feat_dim, vocab=vocab, n_videos=60, n_queries=200,
                             spread=4.0, batch_size=120, b=0.1, lamb=0.4), **out)


def _train_with_reference(model, ds, annotations, epochs, batch_size=120):
    """Brief training with the reference's own ``Trainer.train_epoch`` (main.py:43-79), its real negative sampler
    (data.py:249-337) and the optimiser of main.py:358, seeds as main.py:237-239."""
    import torch.optim as optim
    rmain = ref_harness.load(REF, with_main=True).main
    random.seed(123); np.random.seed(123); torch.random.manual_seed(123)
    it = DataLoader(ds, shuffle=False, collate_fn=rdata.custom_collate,
                    batch_sampler=rdata.CustomBatchSampler(batch_size, annotations, ds.num_segments_info, same_length=True))
    tr = rmain.Trainer(train_writer=ref_harness.NullWriter(), device="cpu", compute_grads=False)
    opt = optim.Adam(model.parameters(), lr=5e-4, weight_decay=5e-3)
    losses = [float(tr.train_epoch(model, it, opt)) for _ in range(epochs)]
    model.eval()
    return losses


def _corpus_eval_arrays(model, vit, lit, annotations, queries, moments, thresholds=(0.5, 0.7)):
    """Embeddings of every video / query and, per query, the rank of the first positive in the reference's own
    ordering (the ops of evaluate.py:35,44,53-58,71,77)."""
    with torch.no_grad():
        vemb = [model(b["feature"]) for b in vit]
        qemb = torch.cat([model(b["feature"], False, "cpu") for b in lit], dim=0)
    vid_off = np.cumsum([0] + [int(v.shape[0]) for v in vemb]).astype(np.int64)
    ranks = {t: [] for t in thresholds}
    for q in range(qemb.shape[0]):
        dist_all = []
        gts = {t: [] for t in thresholds}
        for vi, ve in enumerate(vemb):
            n = ve.size(0)
            dist = torch.nn.functional.pairwise_distance(ve, qemb[q:q + 1].repeat(n, 1))
            for s, e in moments[n]:
                dist_all.append(dist.index_select(0, torch.arange(s, e + 1)).mean().item())
                for thr in gts:
                    if vi == int(queries["video_idx"][q]):
                        gts[thr].append(int((rutils.get_iou(queries["times"][q], s, e) > thr).sum() >= 2))
                    else:
                        gts[thr].append(0)
        order = np.argsort(dist_all)
        for thr in gts:
            ranks[thr].append(int(np.where(np.array(gts[thr])[order] == 1)[0][0]))
    return torch.cat(vemb).numpy(), vid_off, qemb.numpy(), {t: np.asarray(r) for t, r in ranks.items()}


def case_val_trained():
    """The val shape (1,094 videos x 4096-d, 21.9 k moments) with a model the reference's own Trainer has trained
    for a few epochs on these queries, so that R@1 / R@10 / R@100 are NOT all zero (the untrained val_eval golden
    only carries information in MR).  The 52 MB of trained weights are not committed: the file holds the
    embeddings the trained reference model produced, the reference's metric dicts and its per-query ranks."""
    seed, n_videos, n_queries, feat_dim, vocab, epochs = 123, 1094, 192, 4096, 2000, 20
    videos = synth.make_videos(seed, n_videos, feat_dim)
    queries = synth.make_queries(seed + 7, videos, n_queries, vocab)
    sd = synth.make_state_dict(seed, feat_dim, vocab, spread=1.0)
    model = _ref_model(sd, feat_dim)
    ds_train, annotations = _ref_dataset(videos, queries, validate=False)
    losses = _train_with_reference(model, ds_train, annotations, epochs)
    ds, annotations = _ref_dataset(videos, queries)
    vit, lit, moments = _ref_iters(ds, videos, annotations)
    torch.random.manual_seed(123); random.seed(123); np.random.seed(123)
    with redirect_stdout(io.StringIO()):
        m_corpus = reval.evaluate(model, vit, lit, annotations, "cpu", preliminary=10 ** 9, model_types=["model", "chance"])
    prior = synth.make_prior(sorted(set(v["num_segments"] for v in videos)))
    torch.random.manual_seed(123); random.seed(123); np.random.seed(123)
    with redirect_stdout(io.StringIO()):
        m_single = rsingle.evaluate(model, vit, lit, annotations, "cpu", ["model", "chance", "prior"], prior)
    vemb, vid_off, qemb, ranks = _corpus_eval_arrays(model, vit, lit, annotations, queries, moments)
    meta = dict(seed=seed, query_seed=seed + 7, n_videos=n_videos, n_queries=n_queries, feat_dim=feat_dim, vocab=vocab,
                spread=1.0, epochs=epochs, train_losses=losses, seg_choices=[6, 5], seg_probs=[0.83, 0.17],
                metrics_corpus=_metrics_json(m_corpus), metrics_single=_metrics_json(m_single))
    _save("val_trained", meta, video_emb=vemb, vid_off=vid_off, query_emb=qemb, rank_05=ranks[0.5], rank_07=ranks[0.7])


MID = dict(seed=71, n_videos=200, n_queries=160, feat_dim=256, vocab=300, hidden=128, epochs=30)


def _mid_setup(train=True):
    c = MID
    videos = synth.make_videos(c["seed"], c["n_videos"], c["feat_dim"])
    queries = synth.make_queries(c["seed"], videos, c["n_queries"], c["vocab"])
    sd = synth.make_state_dict(c["seed"], c["feat_dim"], c["vocab"], hidden=c["hidden"], spread=1.0)
    model = _ref_model(sd, c["feat_dim"])
    losses = []
    if train:
        ds_train, annotations = _ref_dataset(videos, queries, validate=False)
        losses = _train_with_reference(model, ds_train, annotations, c["epochs"])
    return videos, queries, model, losses


def case_mid_trained():
    """A mid-size corpus (200 videos, 4,0xx moments) with a briefly TRAINED reference model whose weights ARE
    committed (hidden_size 128 keeps them at ~2 MB): pins the full drop-in calls - ``evaluate.evaluate``,
    ``evaluate_single.evaluate`` and ``Trainer.validate_epoch`` (main.py:121-212, both the ``size=250`` scalars and
    the ``size=-1`` precision/recall curves) - on a model with informative R@k."""
    rmain = ref_harness.load(REF, with_main=True).main
    videos, queries, model, losses = _mid_setup()
    ds, annotations = _ref_dataset(videos, queries)
    vit, lit, moments = _ref_iters(ds, videos, annotations)
    torch.random.manual_seed(123); random.seed(123); np.random.seed(123)
    with redirect_stdout(io.StringIO()):
        m_corpus = reval.evaluate(model, vit, lit, annotations, "cpu", preliminary=10 ** 9, model_types=["model", "chance"])
    prior = synth.make_prior(sorted(set(v["num_segments"] for v in videos)))
    torch.random.manual_seed(123); random.seed(123); np.random.seed(123)
    with redirect_stdout(io.StringIO()):
        m_single = rsingle.evaluate(model, vit, lit, annotations, "cpu", ["model", "chance", "prior"], prior)
    validate = {}
    for tag, size in (("size250", 250), ("size100", 100), ("all", -1)):
        w = ref_harness.NullWriter()
        tr = rmain.Trainer(val_writer=w, device="cpu")
        tr.global_step = 7
        with redirect_stdout(io.StringIO()):
            pr = tr.validate_epoch(model, vit, lit, annotations, size=size)
        validate[tag] = dict(pr_curve={k: {str(kk): [float(x) for x in vv] for kk, vv in v.items()} for k, v in pr.items()},
                             scalars=[[name, vals, step] for name, vals, step in w.scalar_groups], figures=w.figures)
    vemb, vid_off, qemb, ranks = _corpus_eval_arrays(model, vit, lit, annotations, queries, moments)
    weights = {f"w:{k}": v.detach().numpy() for k, v in model.state_dict().items()}
    meta = dict(MID, spread=1.0, train_losses=losses, metrics_corpus=_metrics_json(m_corpus),
                metrics_single=_metrics_json(m_single), validate=validate, seg_choices=[6, 5], seg_probs=[0.83, 0.17])
    _save("mid_trained", meta, video_emb=vemb, vid_off=vid_off, query_emb=qemb, rank_05=ranks[0.5], rank_07=ranks[0.7],
          **weights)


def case_train_full_step():
    """Whole training steps of the reference (main.py:57-67: four forwards, ranking loss, backward, Adam lr 5e-4 /
    wd 5e-3 of main.py:358) from the untrained mid-size model with dropout disabled (``dropout_rate=0``: the only
    RNG-coupled op of a step).  Stores the batches (one epoch of the 160 queries = 2) the reference's sampler produced, the loss and
    the mean gradient norm (utils.py:85-92) of every step and the weights after the first and after the last step."""
    import torch.optim as optim
    rmain = ref_harness.load(REF, with_main=True).main
    c = MID
    videos = synth.make_videos(c["seed"], c["n_videos"], c["feat_dim"])
    queries = synth.make_queries(c["seed"], videos, c["n_queries"], c["vocab"])
    sd = synth.make_state_dict(c["seed"], c["feat_dim"], c["vocab"], hidden=c["hidden"], spread=1.0)
    out = {}
    for norm in (False, True):
        model = _ref_model(sd, c["feat_dim"], dropout_rate=0.0)
        ds, annotations = _ref_dataset(videos, queries, validate=False)
        random.seed(123); np.random.seed(123); torch.random.manual_seed(123)
        it = DataLoader(ds, shuffle=False, collate_fn=rdata.custom_collate,
                        batch_sampler=rdata.CustomBatchSampler(120, annotations, ds.num_segments_info, same_length=True))
        batches = [b for b in it][:3]
        w = ref_harness.NullWriter()
        tr = rmain.Trainer(train_writer=w, device="cpu", compute_grads=True, normalize_loss=norm)
        opt = optim.Adam(model.parameters(), lr=5e-4, weight_decay=5e-3)
        tag = f"norm{int(norm)}"
        for step, batch in enumerate(batches):
            if not norm:
                for k in ("posit", "intra", "inter", "lang", "maskp", "maskn"):
                    out[f"b{step}_{k}"] = batch[k].numpy()
            epoch_loss = tr.train_epoch(model, [batch], opt)
            out[f"{tag}_s{step}_epoch_loss"] = np.asarray(epoch_loss, dtype=np.float64)
            if step in (0, len(batches) - 1):
                for k, v in model.state_dict().items():
                    # (normalize_loss=True: the two output projections only - the file stays small)
                    if k != "word_embedding.weight" and (not norm or (step > 0 and k.startswith(("lang_fc", "visual_fc.2")))):
                        out[f"{tag}_s{step}_w:{k}"] = v.detach().numpy().copy()
        out[f"{tag}_logged"] = np.asarray([v for _, v, _ in w.scalars], dtype=np.float64)   # loss/n, grad_norm per step
    _save("train_full_step", dict(MID, spread=1.0, lr=5e-4, weight_decay=5e-3, steps=len(batches), batch_size=120), **out)


CASES = dict(val_trained=case_val_trained, mid_trained=case_mid_trained, train_full_step=case_train_full_step,
             tiny_eval=case_tiny_eval, long_eval=case_long_eval, text=case_text, pool=case_pool,
             train_step=case_train_step, val_eval=case_val_eval)

if __name__ == "__main__":
    os.makedirs(OUT, exist_ok=True)
    for c in (sys.argv[1:] or list(CASES)):
        t = time.time()
        CASES[c]()
        print(f"{c}: {time.time() - t:.1f} s")
