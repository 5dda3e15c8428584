"""Seeded DiDeMo-shaped synthetic inputs (NumPy only, no file or network access).

Shapes and statistics follow SURVEY.md section 8(d): videos of 5 or 6 five-second clips (83 % / 17 %),
non-negative fc7-like features, 1-6 queries per video, 4 annotators per query of which at least
two agree (the reference crashes otherwise: ``model/evaluate.py:77``), query lengths with mean
~7.5 tokens zero-padded to 20 (reference ``model/data.py:100-104``).

Used by the tests, ``bench.py``, ``__graft_entry__.smoke`` and ``oracle/gen_golden.py``; everything
is a pure function of the seed through ``np.random.default_rng`` (PCG64, stable across NumPy
versions) so the golden fixtures store outputs only.
"""
import numpy as np

from .utils import generate_moments

MAX_QUERY_LEN = 20
_LEN_VALUES = np.array([1, 2, 3, 4, 5])
_LEN_PROBS = np.array([0.72, 0.22, 0.04, 0.013, 0.007])


def _l2n(x, eps=1e-5):
    return (x / (np.linalg.norm(x, axis=-1, keepdims=True) + eps)).astype(np.float32)


def make_videos(seed, n_videos, feat_dim, seg_choices=(6, 5), seg_probs=(0.83, 0.17)):
    """Per-video pooled features as the reference's ``CustomDataset.video_features`` holds them
    (``model/data.py:182-186``): ``segment_features [n, F]`` (float64 storage of fp32 values),
    ``context_features [F]`` fp32, ``num_segments``."""
    rng = np.random.default_rng(seed)
    n_seg = rng.choice(np.asarray(seg_choices), size=n_videos, p=np.asarray(seg_probs))
    videos = []
    for v in range(n_videos):
        n = int(n_seg[v])
        base = np.maximum(rng.standard_normal((1, feat_dim), dtype=np.float32), 0)
        seg = np.maximum(base + 0.7 * rng.standard_normal((n, feat_dim), dtype=np.float32), 0)
        ctx = seg.mean(axis=0)
        videos.append(dict(name=f"vid{v:07d}", num_segments=n,
                           segment_features=_l2n(seg).astype(np.float64),
                           context_features=_l2n(ctx)))
    return videos


def make_frames(seed, n_frames, feat_dim):
    """One video's raw per-frame features, the ``get_rgb_features.py`` output format: fp32
    ``[F, feat_dim]``, non-negative (post-ReLU fc7)."""
    rng = np.random.default_rng(seed)
    return np.maximum(rng.standard_normal((n_frames, feat_dim), dtype=np.float32) + 0.1, 0)


def make_queries(seed, videos, n_queries, vocab, n_annot=4):
    """Queries: ``tokens int64 [Q, 20]`` (0 = pad), ``video_idx int64 [Q]``, ``times`` list of
    ``n_annot`` inclusive ``[start, end]`` pairs with at least two identical, ``annot_id``."""
    rng = np.random.default_rng(seed + 1)
    n_videos = len(videos)
    vid = np.sort(rng.integers(0, n_videos, size=n_queries))
    # make sure early videos are not starved when Q < V: keep the draw, it is only synthetic
    tokens = np.zeros((n_queries, MAX_QUERY_LEN), dtype=np.int64)
    times = []
    for q in range(n_queries):
        n = videos[int(vid[q])]["num_segments"]
        qlen = int(np.clip(rng.poisson(6.5) + 1, 1, MAX_QUERY_LEN))
        tokens[q, :qlen] = rng.integers(1, vocab, size=qlen)
        length = int(min(rng.choice(_LEN_VALUES, p=_LEN_PROBS), n - 1))
        start = int(rng.integers(0, n - length + 1))
        agreed = [start, start + length - 1]
        ann = [list(agreed), list(agreed)]
        for _ in range(n_annot - 2):
            if rng.random() < 0.5:
                ann.append(list(agreed))
            else:
                l2 = int(min(rng.choice(_LEN_VALUES, p=_LEN_PROBS), n - 1))
                s2 = int(rng.integers(0, n - l2 + 1))
                ann.append([s2, s2 + l2 - 1])
        order = rng.permutation(n_annot)
        times.append([ann[i] for i in order])
    return dict(tokens=tokens, video_idx=vid.astype(np.int64), times=times,
                annot_id=[f"a{q:07d}" for q in range(n_queries)])


def make_state_dict(seed, feat_dim, vocab, emb_dim=100, hidden=1000, normalize_lang=False,
                    spread=1.0):
    """``CALModel`` parameters with the reference's shapes and init ranges
    (``model/models.py:7-10,21-48``): Linear weights U(-0.08, 0.08), biases 0, LSTM
    U(-1/sqrt(H), 1/sqrt(H)), GloVe-like table N(0, 0.4) with a zero pad row.  ``spread`` scales
    the two output projections so that scores are not squeezed into a ~0.03-wide band
    (SURVEY.md H1).  Returned as a dict of fp32 NumPy arrays keyed like the reference's
    ``state_dict``."""
    rng = np.random.default_rng(seed + 2)

    def uni(shape, a):
        return rng.uniform(-a, a, size=shape).astype(np.float32)

    k = 1.0 / np.sqrt(hidden)
    sd = {
        "visual_fc.0.weight": uni((500, 2 * feat_dim + 2), 0.08),
        "visual_fc.0.bias": np.zeros(500, np.float32),
        "visual_fc.2.weight": uni((emb_dim, 500), 0.08) * np.float32(spread),
        "visual_fc.2.bias": np.zeros(emb_dim, np.float32),
        "word_embedding.weight": (0.4 * rng.standard_normal((vocab, emb_dim))).astype(np.float32),
        "lang_fc.weight": uni((emb_dim, 2 * hidden), 0.08) * np.float32(spread),
        "lang_fc.bias": np.zeros(emb_dim, np.float32),
    }
    sd["word_embedding.weight"][0] = 0
    for suffix in ("", "_reverse"):
        sd[f"lstm.weight_ih_l0{suffix}"] = uni((4 * hidden, emb_dim), k)
        sd[f"lstm.weight_hh_l0{suffix}"] = uni((4 * hidden, hidden), k)
        sd[f"lstm.bias_ih_l0{suffix}"] = uni((4 * hidden,), k)
        sd[f"lstm.bias_hh_l0{suffix}"] = uni((4 * hidden,), k)
    if normalize_lang:
        ll = (1.0 + 0.1 * rng.standard_normal((vocab, 1))).astype(np.float32)
        ll[0] = 0
        sd["learnable_length.weight"] = ll
    return sd


def clip_features(video):
    """The reference's eval-time visual input for one video, ``[n, 2F+2]`` fp32 =
    ``[segment | context | (i/n, (i+1)/n)]`` (``model/data.py:204-213`` with start=0, end=n-1)."""
    n = video["num_segments"]
    seg = video["segment_features"].astype(np.float32)
    ctx = np.repeat(video["context_features"].reshape(1, -1).astype(np.float32), n, axis=0)
    i = np.arange(n, dtype=np.float32).reshape(-1, 1)
    tef = np.concatenate([i, i + 1], axis=1) / np.float32(n)
    return np.concatenate([seg, ctx, tef.astype(np.float32)], axis=1)


def make_bank(seed, n_videos, n_seg, dim, dtype=np.float32, scale=0.25):
    """Corpus-scale shortcut: clip embeddings generated directly,
    ``[n_videos * n_seg, dim]`` - a shared per-video component plus per-clip variation."""
    rng = np.random.default_rng(seed + 3)
    base = rng.standard_normal((n_videos, 1, dim), dtype=np.float32)
    clip = rng.standard_normal((n_videos, n_seg, dim), dtype=np.float32)
    return ((base + 0.6 * clip) * np.float32(scale)).reshape(-1, dim).astype(dtype)


def make_query_embeddings(seed, n_queries, dim, scale=0.25):
    rng = np.random.default_rng(seed + 4)
    return (rng.standard_normal((n_queries, dim), dtype=np.float32) * np.float32(scale))


def make_prior(seg_counts=(5, 6)):
    """A deterministic stand-in for the reference's train-set moment-frequency prior
    (``model/evaluate_single.py:129-146``): shorter moments first."""
    return {n: sorted(generate_moments(n), key=lambda m: (m[1] - m[0], m[0])) for n in seg_counts}
