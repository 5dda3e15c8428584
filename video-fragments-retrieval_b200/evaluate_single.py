"""Single-video (classic DiDeMo) protocol - drop-in for the reference's ``model/evaluate_single.py``.

Same signature and return structure as ``evaluate_single.py:28,75-87``.  The 21 (or 15, or 465)
scores of the query's own video come from ``vfr_score_own``; the ranking, the per-annotator
ranks, the top-1 integer IoU and the first-positive positions come from the K5 kernels; float64
means stay in NumPy.

REFERENCE QUIRK kept by default (``REFERENCE_COMPAT = True``): ``evaluate_single.py:54`` reverses
the ascending-distance order, so the reference's top-1 is the FARTHEST moment.  Set
``REFERENCE_COMPAT = False`` for the ascending (nearest-first) ranking.
"""
import itertools
import random

import numpy as np
import torch

from . import ops
from .evaluate import collect_embeddings
from .utils import generate_moments, moment_index

REFERENCE_COMPAT = True


def _order_tensor(lists, m_stride, device):
    arr = np.full((len(lists), m_stride), -1, dtype=np.int32)
    for q, l in enumerate(lists):
        arr[q, :len(l)] = l
    return torch.from_numpy(arr).to(device)


def evaluate_embedded(bank, q_emb, q_video, times_list, model_types=("model",), prior=(),
                      iou_thresholds=(0.5, 0.7), py_random=random):
    model_types = list(model_types)
    iou_thresholds = list(iou_thresholds)
    dev = bank.device
    Q = len(times_list)
    q_video = np.asarray(q_video)
    nseg = bank.nseg_host[q_video]
    q_video_t = torch.as_tensor(q_video.astype(np.int32), device=dev)
    q_nseg = torch.as_tensor(nseg, device=dev)
    times = ops.pack_times(times_list, dev)
    tables = ops.threshold_tables(iou_thresholds, False, dev)
    own = ops.score_own(bank, q_emb, q_video_t)
    m_stride = own.shape[1]

    # ranked lists per ranker; chance / prior are drawn on the host in the reference's order
    # (random.sample at :55 and prior[num_segments] at :56 run for EVERY query, whatever model_types)
    moments = {int(n): generate_moments(int(n)) for n in np.unique(nseg)}
    chance_lists, prior_lists = [], []
    for q in range(Q):
        n = int(nseg[q])
        sample = py_random.sample(moments[n], k=len(moments[n]))
        chance_lists.append([moment_index(n, s, e) for (s, e) in sample])
        pr = prior[n]
        prior_lists.append([moment_index(n, s, e) for (s, e) in pr])
    orders = {"model": ops.rank_order(own, q_nseg, descending=REFERENCE_COMPAT)}
    if "chance" in model_types:
        orders["chance"] = _order_tensor(chance_lists, m_stride, dev)
    if "prior" in model_types:
        orders["prior"] = _order_tensor(prior_lists, m_stride, dev)

    rank_metrics = {mt: {1: [], 5: [], 10: [], "mIoU": []} for mt in model_types}
    recall_metrics = {(mt, thr): {1: [], 5: [], 10: []} for mt, thr in itertools.product(model_types, iou_thresholds)}
    n_annot = np.array([len(t) for t in times_list])
    for mt in model_types:
        if mt not in orders:
            raise KeyError(mt)
        ranks, t_i, t_u, first = (x.cpu().numpy() for x in ops.single_metrics(orders[mt], q_nseg, times, tables))
        if (ranks < 0).any():
            q, a = np.argwhere(ranks < 0)[0]
            raise ValueError(f"{tuple(times_list[q][a])} is not in list")
        for q in range(Q):
            a = int(n_annot[q])
            r = ranks[q, :a]
            ious = t_i[q, :a] / t_u[q, :a]
            for k in (1, 5, 10):
                rank_metrics[mt][k].append(int(np.mean(np.sort(r)[:3]) <= k))
            rank_metrics[mt]["mIoU"].append(np.mean(np.sort(ious)[-3:]))
        for ti, thr in enumerate(iou_thresholds):
            for k in (1, 5, 10):
                recall_metrics[(mt, thr)][k] = (first[:, ti] < k).astype(int).tolist()

    metrics = {}
    for mt, d in rank_metrics.items():
        metrics[mt] = {(k if k == "mIoU" else f"Rank@{k}"): np.mean(v) * 100 for k, v in d.items()}
    for (mt, thr), d in recall_metrics.items():
        metrics[f"{mt}, IoU={thr}"] = {f"Recall@{k}": np.mean(v) * 100 for k, v in d.items()}
    return metrics


def evaluate(model, video_iterator, lang_iterator, annotations, device, model_types=["model"], prior=[],
             iou_thresholds=[0.5, 0.7]):
    """Drop-in for reference ``model/evaluate_single.py:28-87``."""
    bank, names, q_emb, q_names, q_annots = collect_embeddings(model, video_iterator, lang_iterator, device)
    index = {name: i for i, name in enumerate(names)}
    q_video = [index[n] for n in q_names]
    times_list = [annotations[a]["times"] for a in q_annots]
    print("\nEvaluation:")
    return evaluate_embedded(bank, q_emb, q_video, times_list, model_types, prior, iou_thresholds)
