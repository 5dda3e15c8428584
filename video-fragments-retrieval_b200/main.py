"""``Trainer`` - drop-in for the hot-path members of the reference's ``model/main.py:27-232``.

* ``ranking_loss``   (main.py:214-232)  -> K6 forward/backward kernels (``ops.ranking_loss``);
* ``validate_epoch`` (main.py:121-212)  -> the corpus scoring core of ``evaluate.py`` with the ``>=``
  IoU rule, 1-based ranks, MRR and the precision/recall accumulators;
* ``train_epoch`` / ``test_epoch`` (main.py:43-119) keep the reference's step structure around the
  four ``model(...)`` calls and the loss.

The CLI driver, optimiser/scheduler construction, checkpoint writing and the TensorBoard/matplotlib
plumbing of main.py:235-390 are out of scope; writers passed in by the caller
are used duck-typed (``add_scalar`` / ``add_scalars``).
"""
from collections import defaultdict

import numpy as np
import torch

from . import ops
from .evaluate import collect_embeddings, rank_first_positive


def grad_norm(model):
    """Mean per-parameter gradient L2 norm (reference model/utils.py:85-92): one kernel over all gradient tensors and one
    read-back instead of a ``.norm().item()`` round trip per parameter."""
    grads = [p.grad.detach().float().contiguous() for p in model.parameters() if p.grad is not None]
    norms = ops.grad_norms(grads).cpu().tolist()
    return sum(norms) / len(norms)


class FusedAdam(torch.optim.Optimizer):
    """``torch.optim.Adam(params, lr, weight_decay)`` as the reference configures it (main.py:358) on ONE kernel launch
    per step for all parameter tensors (``vfr_adam_step``): L2 weight decay folded into the gradient, bias-corrected
    moments, eps after the square root.  State layout (``step`` / ``exp_avg`` / ``exp_avg_sq`` per parameter) as torch's."""

    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0):
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay))

    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        for group in self.param_groups:
            ps = [p for p in group["params"] if p.grad is not None]
            if not ps:
                continue
            for p in ps:
                st = self.state[p]
                if not st:
                    st["step"] = 0
                    st["exp_avg"] = torch.zeros_like(p, memory_format=torch.preserve_format)
                    st["exp_avg_sq"] = torch.zeros_like(p, memory_format=torch.preserve_format)
                st["step"] += 1
            by_step = {}
            for p in ps:
                by_step.setdefault(self.state[p]["step"], []).append(p)
            for step, group_ps in by_step.items():
                ops.adam_step([p.data for p in group_ps], [p.grad.contiguous() for p in group_ps],
                              [self.state[p]["exp_avg"] for p in group_ps], [self.state[p]["exp_avg_sq"] for p in group_ps],
                              step, lr=group["lr"], betas=group["betas"], eps=group["eps"], weight_decay=group["weight_decay"])
        return loss


class Trainer:

    def __init__(self, train_writer=None, test_writer=None, val_writer=None, normalize_loss=False,
                 compute_grads=True, device=None, bert=False, b=0.1, lamb=0.4):
        self.train_writer, self.test_writer, self.val_writer = train_writer, test_writer, val_writer
        self.normalize_loss = normalize_loss
        self.compute_grads = compute_grads
        self.device = device
        self.bert = bert
        self.b = b
        self.lamb = lamb
        self.global_step = 0

    # -- K6 ----------------------------------------------------------------------------------
    def ranking_loss(self, posit_emb, intra_emb, inter_emb, lang_emb, maskp, maskn):
        """-> (loss: 0-d tensor, the SUM over samples; n_samples: int) - main.py:214-232."""
        n_samples = maskp.max().item() + 1
        loss = ops.ranking_loss(posit_emb, intra_emb, inter_emb, lang_emb, maskp, maskn, n_samples,
                                normalize=self.normalize_loss, b=self.b, lamb=self.lamb)
        return loss, n_samples

    def _embed_batch(self, model, batch):
        dev = self.device
        posit = model(batch["posit"].to(dev))
        intra = model(batch["intra"].to(dev))
        inter = model(batch["inter"].to(dev))
        lang = model(batch["lang"].to(dev), False, dev, self.bert)
        return posit, intra, inter, lang, batch["maskp"].to(dev), batch["maskn"].to(dev)

    def train_epoch(self, model, iterator, optimizer):
        model = model.to(self.device)
        model.train()
        total, count = 0.0, 0
        for batch in iterator:
            optimizer.zero_grad()
            loss, n = self.ranking_loss(*self._embed_batch(model, batch))
            value = loss.item()
            total += value
            count += n
            loss.backward()
            optimizer.step()
            if self.train_writer is not None:
                self.train_writer.add_scalar("loss", value / n, global_step=self.global_step)
                if self.compute_grads:
                    self.train_writer.add_scalar("grad_norm", grad_norm(model), global_step=self.global_step)
            self.global_step += 1
        return total / count

    def test_epoch(self, model, iterator):
        model = model.to(self.device)
        model.eval()
        total, count = 0.0, 0
        posit = lang = None
        for batch in iterator:
            with torch.no_grad():
                posit, intra, inter, lang, maskp, maskn = self._embed_batch(model, batch)
                loss, n = self.ranking_loss(posit, intra, inter, lang, maskp, maskn)
            total += loss.item()
            count += n
        mean_loss = total / count
        if self.test_writer is not None:
            self.test_writer.add_scalar("loss", mean_loss, global_step=self.global_step)
            self.test_writer.add_scalars("embedding_norm", dict(video=posit.norm(dim=1).mean().item(),
                                                                language=lang.norm(dim=1).mean().item()),
                                         global_step=self.global_step)
        return mean_loss

    # -- validation (corpus protocol inside training) ----------------------------------------------
    def validate_epoch(self, model, video_iterator, lang_iterator, annotations, size=250,
                       iou_thresholds=[0.5, 0.7], atk=[1, 10, 100]):
        """main.py:121-212: ``>=`` IoU rule (:161), thresholds 0.0..1.0 when ``size == -1`` (:138),
        1-based rank (:170), CustomRecall / MedianRank / MRR scalars, precision/recall curves."""
        model.eval()
        bank, names, q_emb, q_names, q_annots = collect_embeddings(model, video_iterator, lang_iterator, self.device,
                                                                      bert=self.bert)
        if size != -1:
            q_emb, q_names, q_annots = q_emb[:size], q_names[:size], q_annots[:size]
        index = {name: i for i, name in enumerate(names)}
        q_video = [index[n] for n in q_names]
        times_list = [annotations[a]["times"] for a in q_annots]
        thr_range = [i / 10 for i in range(11)] if size == -1 else list(iou_thresholds)
        res = rank_first_positive(bank, q_emb, q_video, times_list, thr_range, inclusive=True, want_topk=max(atk))
        if (res["npos"] == 0).any():
            raise IndexError("index 0 is out of bounds for axis 0 with size 0")
        n_q = len(times_list)
        custom = defaultdict(lambda: defaultdict(list))
        recipr_rank, median_rank = defaultdict(list), defaultdict(list)
        true_posit = defaultdict(lambda: defaultdict(lambda: 0))
        total_relevant = defaultdict(lambda: 0)
        for q in range(n_q):
            for ti, thr in enumerate(thr_range):
                rank = int(res["rank"][q, ti]) + 1
                total_relevant[thr] += int(res["npos"][q, ti])
                if thr in iou_thresholds:
                    median_rank[thr].append(rank)
                    recipr_rank[thr].append(1 / rank)
                for k in atk:
                    if size == -1:
                        true_posit[thr][k] += int(res["topk_hits"][q, ti, :k].sum())
                    if thr in iou_thresholds:
                        custom[thr][k].append(int(rank <= k))
        if self.val_writer is not None:
            scalars = {}
            for thr, values in custom.items():
                scalars.update({f"{k}_IoU0{round(thr * 10)}": np.mean(v) for k, v in values.items()})
            self.val_writer.add_scalars("CustomRecall", scalars, global_step=self.global_step)
            self.val_writer.add_scalars("MedianRank", {f"IoU0{round(t * 10)}": np.median(v) for t, v in median_rank.items()},
                                        global_step=self.global_step)
            self.val_writer.add_scalars("MeanReciprocalRank",
                                        {f"IoU0{round(t * 10)}": np.mean(v) for t, v in recipr_rank.items()},
                                        global_step=self.global_step)
        pr_curve = defaultdict(lambda: defaultdict(list))
        if size == -1:
            for k in atk:
                for thr in thr_range:
                    pr_curve["precision"][k].append(true_posit[thr][k] / (k * n_q))
                    pr_curve["recall"][k].append(true_posit[thr][k] / total_relevant[thr])
        self.last_validation = dict(custom={t: dict(v) for t, v in custom.items()}, median_rank=dict(median_rank),
                                    recipr_rank=dict(recipr_rank))
        return {key: dict(value) for key, value in dict(pr_curve).items()}
