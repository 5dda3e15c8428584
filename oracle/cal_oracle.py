"""CPU ORACLE - TEST INFRASTRUCTURE ONLY.  Never imported by the product package.

A CPU restatement (torch CPU fp32 + NumPy, no CUDA) of the moment-scoring hot path of
mariyashcheg/video-fragments-retrieval.  Only ``tests/``, ``__graft_entry__.smoke()`` and
``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs may import this module, and only as
the checker / the reported CPU baseline.

PARITY PIN: the reference ships no tests, fixtures or golden vectors, so the
pin is the reference itself: ``oracle/gen_golden.py`` imports the unmodified reference from
``/root/reference/model`` in the build container, drives its real ``CALModel``, samplers,
collates, ``evaluate.evaluate``, ``evaluate_single.evaluate``, ``Trainer.ranking_loss`` and
``CustomDataset.load_video_features`` on seeded synthetic inputs, and commits the outputs under
``tests/golden/``.  ``tests/test_oracle_golden.py`` checks every function below against those
files (torch 2.11.0 / numpy 2.3.5 recorded inside each file).

The arithmetic of this path lives in un-vendored, unpinned dependencies of the reference
(PyTorch ATen ``linear`` / ``lstm`` / ``pairwise_distance`` / ``mean``; NumPy ``argsort`` /
``mean`` / ``median``); the restatement spells the algorithms out with elementary ops.

Each function cites the reference lines it follows (paths relative to /root/reference).
"""
import itertools
import random as _pyrandom

import numpy as np
import torch

PAIRWISE_EPS = 1e-6   # F.pairwise_distance default eps, added to the DIFFERENCE (evaluate.py:53)
NORM_EPS = 1e-5       # data.py:177,181 ; main.py:220-223 ; models.py:64
FRAMES_PER_SEGMENT = 25  # FRAMES_PER_SEC * SEC_PER_SEGMENT = SELECT_FPS (data.py:19-21)


# --------------------------------------------------------------------------------------------
# a6 / a7 : moment enumeration and temporal IoU
# --------------------------------------------------------------------------------------------
def generate_moments(n):
    """model/utils.py:71-75 - singles, then itertools.combinations(range(n), 2)."""
    return [(j, j) for j in range(n)] + list(itertools.combinations(range(n), 2))


def get_iou(times, s, e):
    """model/utils.py:78-82 - inclusive integer segments, float64 division."""
    t = np.array(times)
    inter = np.maximum(np.minimum(t[:, 1], e) + 1 - np.maximum(t[:, 0], s), 0)
    union = np.maximum(t[:, 1], e) + 1 - np.minimum(t[:, 0], s)
    return inter / union


def gt_bits(times, moments, thr, inclusive=False):
    """model/evaluate.py:59-62 (``>``) / model/main.py:159-161 (``>=``): a moment is positive iff
    at least two annotators have IoU above the threshold."""
    out = np.zeros(len(moments), dtype=np.int64)
    for m, (s, e) in enumerate(moments):
        iou = get_iou(times, s, e)
        hit = (iou >= thr) if inclusive else (iou > thr)
        out[m] = int(hit.sum() >= 2)
    return out


# --------------------------------------------------------------------------------------------
# a1 : frame -> segment pooling
# --------------------------------------------------------------------------------------------
def segment_pool(frames, pooling="avg"):
    """model/data.py:163-181 (.npy branch).  ``frames`` fp32 [F, dim] -> (segment_features
    float64 [n, dim] holding fp32 values, context_features fp32 [dim], n)."""
    pool = {"avg": np.mean, "max": np.max}[pooling]
    F = frames.shape[0]
    n = F // FRAMES_PER_SEGMENT + (0 if F % FRAMES_PER_SEGMENT == 0 else 1)
    seg = np.zeros((n, frames.shape[1]))
    for i in range(n):
        x = pool(frames[i * FRAMES_PER_SEGMENT:(i + 1) * FRAMES_PER_SEGMENT, :], axis=0)
        seg[i, :] = x / (np.linalg.norm(x) + NORM_EPS)
    x = pool(frames, axis=0)
    ctx = x / (np.linalg.norm(x) + NORM_EPS)
    return seg, ctx, n


def segment_pool_h5(frames):
    """model/data.py:144-161 (``prep=True`` branch on MCN's released fc7 features): always six
    25-row windows of the first 150 rows (mean), drop the 6th if all-zero, context = mean of the
    un-normalised segment means, then normalise."""
    seg = np.zeros((6, frames.shape[1]))
    count = 0
    for i in range(0, min(frames.shape[0], 150), FRAMES_PER_SEGMENT):
        seg[count, :] = np.mean(frames[i:i + FRAMES_PER_SEGMENT, :], axis=0)
        count += 1
    if np.sum(seg[5, :]) == 0:
        seg = seg[:5, :]
    x = np.mean(seg, axis=0)
    ctx = x / (np.linalg.norm(x) + NORM_EPS)
    for i in range(seg.shape[0]):
        seg[i, :] = seg[i, :] / (np.linalg.norm(seg[i, :]) + NORM_EPS)
    return seg, ctx, seg.shape[0]


# --------------------------------------------------------------------------------------------
# a2 : [segment | context | tef] assembly
# --------------------------------------------------------------------------------------------
def make_visual_features(seg, ctx, n, s, e):
    """model/data.py:204-213 -> fp32 [e-s+1, 2*dim+2]."""
    seg_t = torch.from_numpy(np.asarray(seg)[s:e + 1]).float()
    ctx_t = torch.from_numpy(np.asarray(ctx).reshape(1, -1)).float()
    idx = torch.arange(s, e + 1).view(-1, 1)
    tef = torch.cat([idx, idx + 1], dim=1).float() / n
    return torch.cat([seg_t, ctx_t.repeat(seg_t.size(0), 1), tef], dim=1)


# --------------------------------------------------------------------------------------------
# a3 / a4 : embeddings
# --------------------------------------------------------------------------------------------
def _t(x):
    return x if isinstance(x, torch.Tensor) else torch.from_numpy(np.asarray(x))


def visual_embed(sd, feats):
    """model/models.py:21-26,56 in eval mode: Linear(2F+2, 500) -> ReLU -> Linear(500, D);
    Dropout is the identity."""
    x = _t(feats).float()
    h = torch.clamp_min(x @ _t(sd["visual_fc.0.weight"]).t() + _t(sd["visual_fc.0.bias"]), 0)
    return h @ _t(sd["visual_fc.2.weight"]).t() + _t(sd["visual_fc.2.bias"])


def text_embed(sd, tokens, normalize_lang=False):
    """model/models.py:61-66: embedding gather (row 0 = pad = zeros) -> optional
    ``x / (|x| + 1e-5) * len[id]`` -> 1-layer BiLSTM run over ALL 20 positions including the
    padding (no packing), h0 = c0 = 0, PyTorch gate order i, f, g, o, two bias vectors -> concat
    [h_fwd(T-1) | h_bwd(0)] -> Linear(2H, D)."""
    ids = _t(tokens).long()
    x = _t(sd["word_embedding.weight"])[ids]                     # [B, L, E]
    if normalize_lang:
        length = _t(sd["learnable_length.weight"])[ids]          # [B, L, 1]
        x = x / (x.norm(dim=-1, keepdim=True) + NORM_EPS) * length
    B, L, _ = x.shape
    finals = []
    for suffix, order in (("", range(L)), ("_reverse", range(L - 1, -1, -1))):
        w_ih, w_hh = _t(sd[f"lstm.weight_ih_l0{suffix}"]), _t(sd[f"lstm.weight_hh_l0{suffix}"])
        bias = _t(sd[f"lstm.bias_ih_l0{suffix}"]) + _t(sd[f"lstm.bias_hh_l0{suffix}"])
        H = w_hh.shape[1]
        h = torch.zeros(B, H)
        c = torch.zeros(B, H)
        for t in order:
            gates = x[:, t] @ w_ih.t() + h @ w_hh.t() + bias
            i, f, g, o = gates.split(H, dim=1)
            c = torch.sigmoid(f) * c + torch.sigmoid(i) * torch.tanh(g)
            h = torch.sigmoid(o) * torch.tanh(c)
        finals.append(h)
    hcat = torch.cat(finals, dim=1)                              # [B, 2H]
    return hcat @ _t(sd["lang_fc.weight"]).t() + _t(sd["lang_fc.bias"])


# --------------------------------------------------------------------------------------------
# a5 : scoring core
# --------------------------------------------------------------------------------------------
def clip_distances(video_emb, q):
    """model/evaluate.py:53: F.pairwise_distance(v, q.repeat(n,1)) = || v - q + 1e-6 ||_2 per
    clip (eps is added to the difference, inside the norm)."""
    diff = _t(video_emb).float() - _t(q).float().reshape(1, -1) + PAIRWISE_EPS
    return torch.sqrt((diff * diff).sum(dim=1))


def moment_scores_loop(video_emb, q, moments):
    """model/evaluate.py:53-58 restated op for op (one ``index_select().mean().item()`` per
    moment).  This is the form timed as the reference arm / CPU baseline."""
    v = _t(video_emb).float()
    n = v.size(0)
    dist = torch.nn.functional.pairwise_distance(v, _t(q).float().reshape(1, -1).repeat(n, 1))
    return [dist.index_select(0, torch.arange(s, e + 1)).mean().item() for (s, e) in moments]


def moment_scores(video_emb, q, moments):
    """Vectorised equivalent of ``moment_scores_loop`` (fp32 sum / count)."""
    d = clip_distances(video_emb, q)
    out = torch.empty(len(moments), dtype=torch.float32)
    for m, (s, e) in enumerate(moments):
        out[m] = d[s:e + 1].sum() / (e - s + 1)
    return out


def score_matrix(bank, vid_off, queries):
    """All scores of ``queries [Q, D]`` against a bank of clip embeddings ``[C, D]`` with CSR video
    offsets ``vid_off [V+1]``, in the reference's global order (video-major, then
    ``generate_moments(n)``) -> fp32 [Q, M_total].  Direct-difference form, no GEMM expansion."""
    bank = _t(bank).float()
    queries = _t(queries).float()
    Q = queries.shape[0]
    outs = []
    # distances [Q, C] in chunks to bound memory
    dist = torch.empty(Q, bank.shape[0])
    step = max(1, (1 << 24) // max(1, bank.shape[1] * Q))
    for c0 in range(0, bank.shape[0], step):
        diff = bank[c0:c0 + step].unsqueeze(0) - queries.unsqueeze(1) + PAIRWISE_EPS
        dist[:, c0:c0 + step] = torch.sqrt((diff * diff).sum(-1))
    vid_off = np.asarray(vid_off)
    for v in range(len(vid_off) - 1):
        a, b = int(vid_off[v]), int(vid_off[v + 1])
        d = dist[:, a:b]
        for (s, e) in generate_moments(b - a):
            acc = d[:, s].clone()
            for k in range(s + 1, e + 1):
                acc = acc + d[:, k]
            outs.append(acc / (e - s + 1))
    return torch.stack(outs, dim=1) if outs else torch.empty(Q, 0)


# --------------------------------------------------------------------------------------------
# a8 / a9 : corpus-level evaluation
# --------------------------------------------------------------------------------------------
def get_metrics(recalls):
    """model/evaluate.py:19-26."""
    out = {}
    for name, value in recalls.items():
        if name == "MR":
            out[name] = np.median(value)
        else:
            out[f"R@{name}"] = np.mean(value) * 100
    return out


def evaluate_corpus(video_embs, query_embs, query_video, query_times, model_types=("model",),
                    iou_thresholds=(0.5, 0.7), np_random=np.random, return_details=False,
                    inclusive=False):
    """model/evaluate.py:28-90 on already-embedded inputs.

    ``video_embs``: list (iteration order of the video iterator) of fp32 [n_v, D];
    ``query_embs`` [Q, D]; ``query_video[q]`` = index of the query's video; ``query_times[q]`` =
    annotator times.  ``np_random`` supplies the 'chance' permutation exactly as
    ``np.random.choice(np.arange(N), size=N, replace=False)`` (evaluate.py:68) does."""
    keys = list(itertools.product(model_types, iou_thresholds))
    recalls = {k: {1: [], 10: [], 100: [], "MR": []} for k in keys}
    moments = {n: generate_moments(n) for n in set(int(v.shape[0]) for v in video_embs)}
    details = dict(scores=[], rank={thr: [] for thr in iou_thresholds})
    for q in range(len(query_embs)):
        distances = []
        gts = {thr: [] for thr in iou_thresholds}
        for vi, v in enumerate(video_embs):
            mom = moments[int(v.shape[0])]
            distances.extend(moment_scores_loop(v, query_embs[q], mom) if not return_details
                             else moment_scores(v, query_embs[q], mom).tolist())
            for thr in iou_thresholds:
                if vi == int(query_video[q]):
                    gts[thr].extend(gt_bits(query_times[q], mom, thr, inclusive).tolist())
                else:
                    gts[thr].extend([0] * len(mom))
        n_all = len(distances)
        ind_rand = np_random.choice(np.arange(n_all), size=n_all, replace=False)
        order = np.argsort(distances)
        for thr in iou_thresholds:
            gt = np.array(gts[thr], dtype=int)
            ranked = {"model": gt[order], "chance": gt[ind_rand]}
            for mt in model_types:
                for k in recalls[(mt, thr)].keys():
                    if k == "MR":
                        recalls[(mt, thr)][k].append(np.where(ranked[mt] == 1)[0][0])
                    else:
                        recalls[(mt, thr)][k].append(int(ranked[mt][:k].sum() > 0))
            details["rank"][thr].append(int(np.where(gt[order] == 1)[0][0]))
        if return_details:
            details["scores"].append(np.asarray(distances, dtype=np.float32))
    metrics = {f"{mt}, IoU={thr}": get_metrics(recalls[(mt, thr)]) for (mt, thr) in keys}
    return (metrics, details) if return_details else metrics


# --------------------------------------------------------------------------------------------
# a10 : single-video protocol
# --------------------------------------------------------------------------------------------
def evaluate_single(video_embs, query_embs, query_video, query_times, prior,
                    model_types=("model",), iou_thresholds=(0.5, 0.7), py_random=_pyrandom,
                    return_details=False):
    """model/evaluate_single.py:28-87.  NOTE the ``[::-1]`` at :54 - the reference ranks the
    FARTHEST moment first; that is what parity means here.  ``prior[n]`` is evaluated
    unconditionally (:56)."""
    rank_metrics = {mt: {1: [], 5: [], 10: [], "mIoU": []} for mt in model_types}
    recall_metrics = {(mt, thr): {1: [], 5: [], 10: []}
                      for mt, thr in itertools.product(model_types, iou_thresholds)}
    details = dict(order=[], scores=[])
    for q in range(len(query_embs)):
        v = video_embs[int(query_video[q])]
        n = int(v.shape[0])
        mom = generate_moments(n)
        scores = moment_scores_loop(v, query_embs[q], mom)
        order = np.argsort(scores)
        predicts = {"model": [mom[i] for i in order][::-1],
                    "chance": py_random.sample(mom, k=len(mom)),
                    "prior": prior[n]}
        details["order"].append(np.asarray(order))
        details["scores"].append(np.asarray(scores, dtype=np.float32))
        times = query_times[q]
        for mt in rank_metrics.keys():
            ranks, ious = [], []
            for t in times:
                ranks.append(predicts[mt].index(tuple(t)) + 1)
                ious.append(get_iou([predicts[mt][0]], t[0], t[1])[0])
            for k in rank_metrics[mt].keys():
                if k == "mIoU":
                    rank_metrics[mt][k].append(np.mean(np.sort(ious)[-3:]))
                else:
                    rank_metrics[mt][k].append(int(np.mean(np.sort(ranks)[:3]) <= k))
        for mt, thr in recall_metrics.keys():
            bits = np.array([(get_iou(times, s, e) > thr).sum() >= 2
                             for s, e in predicts[mt]]).astype(int)
            for k in recall_metrics[(mt, thr)].keys():
                recall_metrics[(mt, thr)][k].append(int(bits[:k].sum() > 0))
    metrics = {}
    for mt, d in rank_metrics.items():
        metrics[mt] = {(k if k == "mIoU" else f"Rank@{k}"): np.mean(v) * 100 for k, v in d.items()}
    for (mt, thr), d in recall_metrics.items():
        metrics[f"{mt}, IoU={thr}"] = {f"Recall@{k}": np.mean(v) * 100 for k, v in d.items()}
    return (metrics, details) if return_details else metrics


# --------------------------------------------------------------------------------------------
# a11 : ranking loss
# --------------------------------------------------------------------------------------------
def ranking_loss(posit, intra, inter, lang, maskp, maskn, b=0.1, lamb=0.4, normalize=False):
    """model/main.py:214-232 (differentiable torch CPU restatement).  Returns (loss, n)."""
    posit, intra, inter, lang = (_t(x) for x in (posit, intra, inter, lang))
    maskp, maskn = _t(maskp).long(), _t(maskn).long()
    n = int(maskp.max().item()) + 1
    if normalize:
        posit, intra, inter, lang = (x / (x.norm(dim=1, keepdim=True) + NORM_EPS)
                                     for x in (posit, intra, inter, lang))

    def cost(rows, i):
        diff = rows - lang[i].reshape(1, -1) + PAIRWISE_EPS
        return torch.sqrt((diff * diff).sum(dim=1)).mean()

    loss = 0
    for i in range(n):
        cp = cost(posit[maskp == i], i)
        cn = cost(intra[maskn == i], i)
        ci = cost(inter[maskp == i], i)
        loss = loss + torch.relu(cp - cn + b) + lamb * torch.relu(cp - ci + b)
    return loss, n


# --------------------------------------------------------------------------------------------
# f1 : Trainer.validate_epoch
# --------------------------------------------------------------------------------------------
def validate_epoch(video_embs, query_embs, query_video, query_times, size=250, iou_thresholds=(0.5, 0.7),
                   atk=(1, 10, 100)):
    """model/main.py:121-212 on already-embedded inputs: the corpus scoring core with the ``>=`` IoU rule (:161),
    thresholds 0.0 ... 1.0 when ``size == -1`` (:138), 1-based rank (:170), stop after ``size`` queries (:187-188).
    Returns ``(scalars, pr_curve)``: ``scalars`` = the three ``add_scalars`` groups written to the val writer
    (CustomRecall / MedianRank / MeanReciprocalRank, :190-199), ``pr_curve`` = the returned dict (:201-212)."""
    from collections import defaultdict
    iou_thresholds = list(iou_thresholds)
    custom = defaultdict(lambda: defaultdict(list))
    recipr_rank, median_rank = defaultdict(list), defaultdict(list)
    true_posit = defaultdict(lambda: defaultdict(lambda: 0))
    total_relevant = defaultdict(lambda: 0)
    thr_range = [i / 10 for i in range(11)] if size == -1 else iou_thresholds
    moments = {n: generate_moments(n) for n in set(int(v.shape[0]) for v in video_embs)}
    li = -1
    for li in range(len(query_embs)):
        distances = []
        predicts = defaultdict(list)
        for vi, v in enumerate(video_embs):
            mom = moments[int(v.shape[0])]
            distances.extend(moment_scores(v, query_embs[li], mom).tolist())
            for thr in thr_range:
                if vi == int(query_video[li]):
                    predicts[thr].extend(gt_bits(query_times[li], mom, thr, inclusive=True).tolist())
                else:
                    predicts[thr].extend([0] * len(mom))
        order = np.argsort(distances)
        for thr in thr_range:
            ranked = np.array(predicts[thr], dtype=int)[order]
            rank = np.where(ranked == 1)[0][0] + 1
            total_relevant[thr] += ranked.sum()
            if thr in iou_thresholds:
                median_rank[thr].append(rank)
                recipr_rank[thr].append(1 / rank)
            for k in atk:
                if size == -1:
                    true_posit[thr][k] += ranked[:k].sum()
                if thr in iou_thresholds:
                    custom[thr][k].append(int(rank <= k))
        if li == size - 1:
            break
    scalars = {"CustomRecall": {}, "MedianRank": {}, "MeanReciprocalRank": {}}
    for thr, values in custom.items():
        scalars["CustomRecall"].update({f"{k}_IoU0{round(thr * 10)}": float(np.mean(v)) for k, v in values.items()})
    scalars["MedianRank"] = {f"IoU0{round(t * 10)}": float(np.median(v)) for t, v in median_rank.items()}
    scalars["MeanReciprocalRank"] = {f"IoU0{round(t * 10)}": float(np.mean(v)) for t, v in recipr_rank.items()}
    pr_curve = defaultdict(lambda: defaultdict(list))
    if size == -1:
        for k in atk:
            for thr in thr_range:
                pr_curve["precision"][k].append(true_posit[thr][k] / (k * (li + 1)))
                pr_curve["recall"][k].append(true_posit[thr][k] / total_relevant[thr])
    return scalars, {key: dict(value) for key, value in dict(pr_curve).items()}


# --------------------------------------------------------------------------------------------
# f2 : one whole training step (four forwards, loss, backward, Adam)
# --------------------------------------------------------------------------------------------
def adam_update(param, grad, exp_avg, exp_avg_sq, step, lr=5e-4, weight_decay=5e-3, betas=(0.9, 0.999), eps=1e-8):
    """torch.optim.Adam (the optimiser of model/main.py:358; torch 2.11 defaults: L2 weight decay folded into the
    gradient, bias-corrected moments, eps added after the sqrt).  In-place on ``param`` / the two moments; ``step`` is the
    1-based step count.  Third-party algorithm restated from its published form (Kingma & Ba 2015, PyTorch docs)."""
    g = grad + weight_decay * param
    exp_avg.mul_(betas[0]).add_(g, alpha=1 - betas[0])
    exp_avg_sq.mul_(betas[1]).addcmul_(g, g, value=1 - betas[1])
    bias1, bias2 = 1 - betas[0] ** step, 1 - betas[1] ** step
    denom = exp_avg_sq.sqrt() / (bias2 ** 0.5) + eps
    param.addcdiv_(exp_avg, denom, value=-lr / bias1)


TRAINABLE = ("visual_fc.0.weight", "visual_fc.0.bias", "visual_fc.2.weight", "visual_fc.2.bias",
             "lstm.weight_ih_l0", "lstm.weight_hh_l0", "lstm.bias_ih_l0", "lstm.bias_hh_l0",
             "lstm.weight_ih_l0_reverse", "lstm.weight_hh_l0_reverse", "lstm.bias_ih_l0_reverse", "lstm.bias_hh_l0_reverse",
             "lang_fc.weight", "lang_fc.bias")


def train_step(sd, batch, state, step, normalize_loss=False, b=0.1, lamb=0.4, lr=5e-4, weight_decay=5e-3):
    """model/main.py:57-67 with dropout disabled: forwards of posit / intra / inter / lang, ``ranking_loss``,
    backward (torch CPU autograd over the restated forwards), Adam.  ``sd``: dict of torch tensors (updated in place),
    ``state``: dict name -> (exp_avg, exp_avg_sq) (created on first use).  Returns (loss, n_samples, mean grad norm) -
    the grad norm is the mean of the per-parameter L2 norms, model/utils.py:85-92."""
    leaves = {k: _t(sd[k]).clone().requires_grad_(True) for k in TRAINABLE}
    full = dict(sd)
    full.update(leaves)
    embs = [visual_embed(full, batch[k]) for k in ("posit", "intra", "inter")]
    lang = text_embed(full, batch["lang"])
    loss, n = ranking_loss(embs[0], embs[1], embs[2], lang, batch["maskp"], batch["maskn"], b=b, lamb=lamb,
                           normalize=normalize_loss)
    grads = torch.autograd.grad(loss, [leaves[k] for k in TRAINABLE])
    gnorm = float(np.mean([g.norm().item() for g in grads]))
    with torch.no_grad():
        for k, g in zip(TRAINABLE, grads):
            if k not in state:
                state[k] = (torch.zeros_like(g), torch.zeros_like(g))
            p = _t(sd[k])
            adam_update(p, g, state[k][0], state[k][1], step, lr=lr, weight_decay=weight_decay)
            sd[k] = p
    return float(loss.item()), n, gnorm
