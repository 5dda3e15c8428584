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
sys.path.insert(0, REF)

_h5 = types.ModuleType("h5py")
sys.modules["h5py"] = _h5
_mpl, _plt = types.ModuleType("matplotlib"), types.ModuleType("matplotlib.pyplot")
_plt.switch_backend = lambda *a, **k: None
_mpl.pyplot = _plt
sys.modules.setdefault("matplotlib", _mpl)
sys.modules.setdefault("matplotlib.pyplot", _plt)

import vfr_b200  # noqa: E402,F401
from vfr_b200 import synth  # noqa: E402

import data as rdata  # noqa: E402  (reference)
import evaluate as reval  # noqa: E402
import evaluate_single as rsingle  # noqa: E402
import models as rmodels  # noqa: E402
import utils as rutils  # noqa: E402
from torch.utils.data import DataLoader  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")
VERSIONS = dict(torch=torch.__version__, numpy=np.__version__)


def _save(name, meta, **arrays):
    meta = dict(meta, versions=VERSIONS, generated_by="oracle/gen_golden.py", case=name)
    path = os.path.join(OUT, name + ".npz")
    np.savez_compressed(path, meta=np.frombuffer(json.dumps(meta).encode(), dtype=np.uint8), **arrays)
    print(f"wrote {path}: {os.path.getsize(path) / 1e6:.2f} MB")


def _ref_model(sd, feat_dim, normalize_lang=False):
    model = rmodels.CALModel(pretrained_emb=torch.from_numpy(sd["word_embedding.weight"]),
                             visual_input_dim=2 * feat_dim + 2, emb_dim=rdata.EMBEDDING_DIM,
                             normalize_lang=normalize_lang)
    model.load_state_dict({k: torch.from_numpy(v) for k, v in sd.items()})
    return model.eval()


def _ref_dataset(videos, queries, validate=True):
    ds = rdata.CustomDataset.__new__(rdata.CustomDataset)
    ds.validate = validate
    ds.video_features = {v["name"]: dict(segment_features=v["segment_features"],
                                         context_features=v["context_features"],
                                         num_segments=v["num_segments"]) for v in videos}
    ds.num_segments_info = {v["name"]: v["num_segments"] for v in videos}
    ds.lang_features = {a: torch.from_numpy(queries["tokens"][i:i + 1]).long()
                        for i, a in enumerate(queries["annot_id"])}
    annotations = {a: dict(video=videos[int(queries["video_idx"][i])]["name"], description="",
                           times=queries["times"][i]) for i, a in enumerate(queries["annot_id"])}
    return ds, annotations


def _ref_iters(ds, videos, annotations, max_seg=None):
    names = [v["name"] for v in videos]
    vit = DataLoader(ds, shuffle=False, collate_fn=rdata.validate_collate,
                     batch_sampler=rdata.VideoBatchSampler(names, ds.num_segments_info))
    lsamp = rdata.LanguageBatchSampler(annotations, ds.num_segments_info)
    if max_seg is not None:   # the reference caps its table at n<7 (data.py:386): inject
        for n in range(7, max_seg + 1):
            lsamp.moments[n] = rutils.generate_moments(n)
    lit = DataLoader(ds, shuffle=False, collate_fn=rdata.validate_collate, batch_sampler=lsamp)
    return vit, lit, lsamp.moments


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


CASES = dict(tiny_eval=case_tiny_eval, long_eval=case_long_eval, text=case_text, pool=case_pool,
             train_step=case_train_step, val_eval=case_val_eval)

if __name__ == "__main__":
    os.makedirs(OUT, exist_ok=True)
    for c in (sys.argv[1:] or list(CASES)):
        t = time.time()
        CASES[c]()
        print(f"{c}: {time.time() - t:.1f} s")
