"""Feature assembly - drop-in for the hot-path members of the reference's ``model/data.py``.

* ``CustomDataset.load_video_features`` (data.py:142-188)  -> K1 ``vfr_segment_pool`` (one batched
  device pass over all videos instead of a per-video NumPy loop);
* ``CustomDataset.make_visual_features`` (data.py:204-213), ``__getitem__`` (data.py:215-246);
* ``custom_collate`` (data.py:340-356, defines the ``maskp``/``maskn`` layout K6 consumes),
  ``validate_collate`` (data.py:412-418);
* ``VideoBatchSampler`` / ``LanguageBatchSampler`` (data.py:359-409) - one video / one query per
  batch, ``.moments`` table read by ``evaluate``.

* ``DeviceBatchSampler`` - the training-batch stream of ``CustomBatchSampler`` + ``custom_collate`` (data.py:249-356) built
  ON THE DEVICE: negative sampling for a whole epoch in one kernel, rows gathered from the pooled features in HBM.

Host-side string work is out of scope: the GloVe text-file parser
``WordIndexer`` (data.py:33-118) is meant to be reused from the reference unchanged (as is its host sampler
``CustomBatchSampler`` when the reference's exact Mersenne-Twister draw sequence matters); ``CustomDataset`` here
accepts any object with the reference's ``items2tensor(list_of_token_lists, max_len)`` method.
"""
import re
from pathlib import Path

import numpy as np
import torch
from torch.utils.data.dataset import Dataset
from torch.utils.data.sampler import BatchSampler

from . import ops
from .utils import generate_moments

SELECT_FPS = 25
FRAMES_PER_SEC = 5
SEC_PER_SEGMENT = 5
FEATURE_DIM = dict(vgg19=4096, resnet152=2048)
EMBEDDING_DIM = 100
_WINDOW = FRAMES_PER_SEC * SEC_PER_SEGMENT

_TOKEN_RE = re.compile(r"('\w )|([\w\d]+)")


def tokenize(description):
    """Query tokeniser of data.py:190-193 (lower-case, alphanumeric runs)."""
    query = description.rstrip("\n ").lower()
    return [m[1] for m in _TOKEN_RE.findall(query) if m[1] != ""]


def pool_videos(frame_arrays, pooling="avg", preprocessed=False, device="cuda", max_frames_per_call=1 << 16):
    """K1 over a list of per-video frame arrays ``[F_v, dim]`` -> list of dicts in the reference's
    ``video_features`` format: ``segment_features`` float64 [n, dim] (fp32 values),
    ``context_features`` fp32 [dim], ``num_segments``."""
    mode = "h5" if preprocessed else pooling
    if mode not in ops.POOL_MODES:
        raise KeyError(pooling)
    out = []
    i = 0
    while i < len(frame_arrays):
        j, tot = i, 0
        while j < len(frame_arrays) and (j == i or tot + len(frame_arrays[j]) <= max_frames_per_call):
            tot += len(frame_arrays[j])
            j += 1
        chunk = [np.ascontiguousarray(a, dtype=np.float32).reshape(len(a), -1) for a in frame_arrays[i:j]]
        off = np.concatenate([[0], np.cumsum([len(a) for a in chunk])])
        frames = torch.from_numpy(np.concatenate(chunk)).to(device, non_blocking=True)
        seg, ctx, n_seg = ops.segment_pool(frames, off, mode, _WINDOW)
        seg, ctx, n_seg = seg.cpu().numpy(), ctx.cpu().numpy(), n_seg.cpu().numpy()
        for k in range(j - i):
            n = int(n_seg[k])
            out.append(dict(segment_features=seg[k, :n].astype(np.float64), context_features=ctx[k].copy(),
                            num_segments=n))
        i = j
    return out


def pool_video_files(readers, pooling="avg", preprocessed=False, device="cuda", chunk_frames=1 << 15):
    """Feature ingestion: per-video frame files -> K1, streamed through PINNED host staging.

    ``readers``: one callable per video returning its frames ``[F_v, dim]`` (``np.load(..., mmap_mode="r")`` for the
    ``get_rgb_features.py`` ``.npy`` files of data.py:164-166, an h5py dataset for MCN's ``.h5`` files of :145-148 - the
    array is copied ONCE, from the file mapping straight into the pinned buffer).  Two pinned buffers of
    ``chunk_frames`` frames alternate: while the host fills one, the other one's host->device copy (its own stream) and
    K1 run; the pooled outputs stay on the device until the end, so nothing in the loop waits for the GPU except the
    reuse of a staging buffer.  Returns the same list of dicts as ``pool_videos``."""
    mode = "h5" if preprocessed else pooling
    if mode not in ops.POOL_MODES:
        raise KeyError(pooling)
    dev = torch.device(device)
    copy_stream = torch.cuda.Stream(device=dev)
    main = torch.cuda.current_stream(dev)
    pinned, staged, copied, freed = [None, None], [None, None], [torch.cuda.Event(), torch.cuda.Event()], [None, None]
    results = []
    pending = None                      # frames of the video that did not fit the previous chunk
    it = iter(readers)
    b, dim, done = 0, None, False
    while not done:
        counts, n = [], 0
        if freed[b] is not None:
            freed[b].synchronize()      # K1 of the chunk that used this staging buffer two rounds ago has finished
        while True:
            if pending is None:
                try:
                    arr = next(it)()
                except StopIteration:
                    done = True
                    break
                arr = arr.reshape(arr.shape[0], -1)
            else:
                arr, pending = pending, None
            if dim is None:
                dim = int(arr.shape[1])
            if pinned[b] is None or (n == 0 and arr.shape[0] > pinned[b].shape[0]):
                pinned[b] = torch.empty((max(chunk_frames, arr.shape[0]), dim), dtype=torch.float32).pin_memory()
                staged[b] = torch.empty(pinned[b].shape, dtype=torch.float32, device=dev)
            if n + arr.shape[0] > pinned[b].shape[0]:
                pending = arr
                break
            np.copyto(pinned[b][n:n + arr.shape[0]].numpy(), arr, casting="same_kind")   # file mapping -> pinned, one copy
            counts.append(int(arr.shape[0]))
            n += int(arr.shape[0])
        if n == 0:
            break
        with torch.cuda.stream(copy_stream):
            staged[b][:n].copy_(pinned[b][:n], non_blocking=True)
            copied[b].record(copy_stream)
        main.wait_event(copied[b])
        off = np.concatenate([[0], np.cumsum(counts)])
        results.append(ops.segment_pool(staged[b][:n], off, mode, _WINDOW))
        freed[b] = torch.cuda.Event()
        freed[b].record(main)
        b ^= 1
    out = []
    for seg, ctx, n_seg in results:
        seg, ctx, n_seg = seg.cpu().numpy(), ctx.cpu().numpy(), n_seg.cpu().numpy()
        for k in range(len(n_seg)):
            m = int(n_seg[k])
            out.append(dict(segment_features=seg[k, :m].astype(np.float64), context_features=ctx[k].copy(), num_segments=m))
    return out


class CustomDataset(Dataset):
    """Same constructor and item format as the reference's ``CustomDataset`` (data.py:121-246)."""

    def __init__(self, videos, annotations, ft_directory, ft_type, word_indexer=None, bert_tokenizer=None,
                 bert_model=None, validate=False, max_query_len=20, pooling="avg", prep=False, device="cuda"):
        self.word_indexer = word_indexer
        self.bert_tokenizer = bert_tokenizer
        self.bert_model = bert_model
        self.max_query_len = max_query_len
        self.ft_directory = ft_directory
        self.ft_type = ft_type
        self.validate = validate
        self.pooling = pooling
        self.preprocessed = prep
        self.device = device
        self.num_segments_info = {}
        self.video_features = {}
        self.load_video_features(videos)
        self.lang_features = {}
        self.load_lang_features(annotations)

    def _read_frames(self, video):
        if self.preprocessed:
            import h5py  # only needed for MCN's released .h5 features (data.py:144-148)
            with h5py.File(Path(self.ft_directory).joinpath(f"fc7_subsample5_fps25_{video}.h5")) as f:
                return np.array(f["features"])
        # memory-mapped: the frames are copied once, from the page cache into the pinned staging buffer
        arr = np.load(Path(self.ft_directory).joinpath(f"features_{self.ft_type}/{self.ft_type}_ft_{video}.npy"), mmap_mode="r")
        return arr.reshape((arr.shape[0], FEATURE_DIM[self.ft_type]))

    def load_video_features(self, videos):
        """data.py:142-188 for all videos: the frame files stream through pinned staging buffers into K1."""
        videos = list(videos)
        pooled = pool_video_files([(lambda v=v: self._read_frames(v)) for v in videos], self.pooling, self.preprocessed,
                                  self.device)
        for video, feats in zip(videos, pooled):
            self.video_features[video] = feats
            self.num_segments_info[video] = feats["num_segments"]

    def load_lang_features(self, annotations):
        for annot_id, info in annotations.items():
            words = tokenize(info["description"])
            if self.word_indexer is not None:
                self.lang_features[annot_id] = self.word_indexer.items2tensor([words], self.max_query_len)
            elif self.bert_tokenizer is not None:
                with torch.no_grad():
                    tokens = self.bert_tokenizer.encode_plus(" ".join(words), return_tensors="pt")
                    self.lang_features[annot_id] = self.bert_model(**tokens)[1]

    def make_visual_features(self, video, start_t, end_t):
        """``[segment | context | (i/n, (i+1)/n)]`` rows for clips start_t..end_t (data.py:204-213)."""
        info = self.video_features[video]
        n = info["num_segments"]
        seg = torch.from_numpy(info["segment_features"][start_t:end_t + 1]).float()
        ctx = torch.from_numpy(np.asarray(info["context_features"]).reshape(1, -1)).float().expand(seg.size(0), -1)
        left = torch.arange(start_t, end_t + 1).view(-1, 1)
        tef = torch.cat([left, left + 1], dim=1).float() / n
        return torch.cat([seg, ctx, tef], dim=1)

    def __getitem__(self, sample):
        if self.validate:
            if "annotation_id" in sample:
                return dict(features=self.lang_features[sample["annotation_id"]], video=sample["video_pos"],
                            annot_id=sample["annotation_id"])
            return dict(features=self.make_visual_features(sample["video_pos"], sample["start_t"], sample["end_t"]),
                        video=sample["video_pos"])
        return dict(
            posit=self.make_visual_features(sample["video_pos"], sample["start_t"], sample["end_t"]),
            intra=self.make_visual_features(sample["video_pos"], sample["start_tn"], sample["end_tn"]),
            inter=self.make_visual_features(sample["video_neg"], sample["start_t"], sample["end_t"]),
            lang=self.lang_features[sample["annotation_id"]])


def custom_collate(batch):
    """Training collate (data.py:340-356): rows of ``posit``/``inter`` belong to sample
    ``maskp[r]``, rows of ``intra`` to ``maskn[r]``; runs are contiguous and ascending."""
    keys = ("posit", "intra", "inter", "lang")
    cat = {k: torch.cat([s[k] for s in batch], dim=0) for k in keys}
    cat["maskp"] = torch.LongTensor([i for i, s in enumerate(batch) for _ in range(s["posit"].size(0))])
    cat["maskn"] = torch.LongTensor([i for i, s in enumerate(batch) for _ in range(s["intra"].size(0))])
    return cat


def validate_collate(batch):
    """Evaluation collate (data.py:412-418): batches hold exactly one video or one query."""
    item = batch[0]
    return dict(feature=item["features"], video=item["video"], annot_id=item.get("annot_id", []))


class VideoBatchSampler(BatchSampler):
    """One whole video per batch (data.py:359-379)."""

    def __init__(self, videos, num_segments_info):
        self.videos = videos
        self.num_segments_info = num_segments_info

    def __iter__(self):
        for video in self.videos:
            yield [dict(video_pos=video, start_t=0, end_t=self.num_segments_info[video] - 1)]

    def __len__(self):
        return len(self.videos)


class LanguageBatchSampler(BatchSampler):
    """One query per batch; ``moments[n]`` is the candidate table (data.py:382-409)."""

    def __init__(self, annotations, num_segments_info, max_segments=6):
        self.annotations = annotations
        self.num_segments_info = num_segments_info
        self.moments = {n: generate_moments(n) for n in range(max_segments + 1)}

    def __iter__(self):
        for annot_id, info in self.annotations.items():
            yield [dict(annotation_id=annot_id, video_pos=info["video"])]

    def __len__(self):
        return len(self.annotations)


class DeviceBatchSampler:
    """Training batches of the reference's ``DataLoader(dataset, batch_sampler=CustomBatchSampler(...),
    collate_fn=custom_collate)`` (data.py:249-356), built on the device.

    Per epoch: one host permutation of the queries (``train=True``), ONE kernel draws every query's positive annotation,
    intra-video negative and inter-video negative (``vfr_sample_negatives``: the reference's distributions from a
    counter-based RNG - not its ``random`` stream), the batches are cut by the reference's rule (a batch closes when
    ``max(posit rows, intra rows) >= batch_size``, :329-332) and every batch's ``posit`` / ``intra`` / ``inter`` rows are
    gathered from the pooled features resident in HBM (``vfr_gather_clip_rows``).  Yields the collate's dict
    ``{posit, intra, inter, lang, maskp, maskn}`` as DEVICE tensors: ``Trainer.train_epoch`` consumes it unchanged.
    A query whose positive fills its whole video under ``same_length=True`` has no intra-video candidate; the reference
    raises ``IndexError`` there (``random.choice([])``, :306) and so does this iterator."""

    def __init__(self, batch_size, annotations, num_segments_info, dataset, train=True, drop_last=False, same_length=True,
                 seed=123, device="cuda"):
        self.batch_size, self.train, self.drop_last, self.same_length = int(batch_size), train, drop_last, same_length
        self.seed, self.epoch, self.device = int(seed), 0, torch.device(device)
        self.annot_ids = list(annotations.keys())
        self.videos = list(num_segments_info.keys())
        index = {v: i for i, v in enumerate(self.videos)}
        feats = [dataset.video_features[v] for v in self.videos]
        nseg = np.asarray([int(num_segments_info[v]) for v in self.videos], dtype=np.int64)
        dev = self.device
        self.seg = torch.from_numpy(np.concatenate([np.asarray(f["segment_features"])[:n] for f, n in zip(feats, nseg)]).astype(np.float32)).to(dev)
        self.ctx = torch.from_numpy(np.stack([np.asarray(f["context_features"], dtype=np.float32).reshape(-1) for f in feats])).to(dev)
        self.vid_off = torch.from_numpy(np.concatenate([[0], np.cumsum(nseg)]).astype(np.int32)).to(dev)
        self.nseg = torch.from_numpy(nseg.astype(np.int32)).to(dev)
        self.times = ops.pack_times([annotations[a]["times"] for a in self.annot_ids], dev)
        self.q_video = torch.from_numpy(np.asarray([index[annotations[a]["video"]] for a in self.annot_ids], dtype=np.int32)).to(dev)
        self.lang = torch.cat([dataset.lang_features[a] for a in self.annot_ids], dim=0).to(dev)
        self._rng = np.random.default_rng(self.seed)

    def __len__(self):
        return len(self.annot_ids)

    def __iter__(self):
        Q = len(self.annot_ids)
        order = self._rng.permutation(Q) if self.train else np.arange(Q)
        samples = ops.sample_negatives(self.times, self.q_video, self.nseg, self.same_length, self.seed, self.epoch).cpu().numpy()
        self.epoch += 1
        if (samples[:, 6] == 1).any():
            raise IndexError("Cannot choose from an empty sequence")          # random.choice([]) in the reference
        if (samples[:, 6] != 0).any():
            raise IndexError("no valid annotation / negative video for some query")
        batch, posit_rows, intra_rows = [], 0, 0
        for q in order:
            batch.append(int(q))
            posit_rows += int(samples[q, 3] - samples[q, 2] + 1)
            intra_rows += int(samples[q, 5] - samples[q, 4] + 1)
            if max(posit_rows, intra_rows) >= self.batch_size:
                yield self._build(batch, samples)
                batch, posit_rows, intra_rows = [], 0, 0
        if batch and not self.drop_last:
            yield self._build(batch, samples)

    def _build(self, batch, samples):
        dev = self.device
        rows = {"posit": ([], []), "intra": ([], []), "inter": ([], [])}
        maskp, maskn = [], []
        for i, q in enumerate(batch):
            vp, vn, st, en, sn, enn = (int(x) for x in samples[q, :6])
            for name, v, a, b in (("posit", vp, st, en), ("intra", vp, sn, enn), ("inter", vn, st, en)):
                rows[name][0].extend([v] * (b - a + 1))
                rows[name][1].extend(range(a, b + 1))
            maskp.extend([i] * (en - st + 1))
            maskn.extend([i] * (enn - sn + 1))
        out = {}
        for name, (rv, rc) in rows.items():
            out[name] = ops.gather_clip_rows(self.seg, self.ctx, self.vid_off, torch.tensor(rv, dtype=torch.int32, device=dev),
                                             torch.tensor(rc, dtype=torch.int32, device=dev))
        out["lang"] = self.lang[torch.tensor(batch, dtype=torch.int64, device=dev)]
        out["maskp"] = torch.tensor(maskp, dtype=torch.int64, device=dev)
        out["maskn"] = torch.tensor(maskn, dtype=torch.int64, device=dev)
        out["samples"] = samples[batch]
        return out
