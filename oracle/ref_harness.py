"""Harness that imports and drives the UNMODIFIED reference - TEST INFRASTRUCTURE ONLY.

Used by ``oracle/gen_golden.py`` (from ``/root/reference/model``, in the build container), by ``bench.py``'s
``--impl reference`` / ``cpu_baseline`` / ``library_baseline.reference_cuda`` legs and by the CPU tests (from
``oracle/_ref``, the verbatim copy ``oracle/build_ref.py`` makes; it travels to the GPU box).  Never imported
by the product package.

What is stubbed, and why it does not touch the timed path: ``h5py`` (imported at
``model/data.py:8``, used only by the ``prep=True`` branch ``:146``) and ``matplotlib.pyplot``
(``model/main.py:9,24``, figures of ``validate_epoch``).  Everything else - ``CALModel``, the samplers and
collates, ``evaluate.evaluate``, ``evaluate_single.evaluate``, ``Trainer`` - is the reference's own code.
"""
import hashlib
import json
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REF_COPY = os.path.join(HERE, "_ref")
REF_SOURCE = "/root/reference/model"
_MODULES = ("utils", "data", "models", "evaluate", "evaluate_single")
_loaded = {}


def _install_stubs():
    if "h5py" not in sys.modules:
        try:
            import h5py  # noqa: F401
        except Exception:
            sys.modules["h5py"] = types.ModuleType("h5py")
    if "matplotlib.pyplot" not in sys.modules:
        try:
            import matplotlib.pyplot  # noqa: F401
        except Exception:
            mpl, plt = types.ModuleType("matplotlib"), types.ModuleType("matplotlib.pyplot")
            plt.switch_backend = lambda *a, **k: None
            plt.figure = lambda *a, **k: None
            plt.plot = lambda *a, **k: None
            plt.grid = lambda *a, **k: None
            mpl.pyplot = plt
            sys.modules["matplotlib"] = mpl
            sys.modules["matplotlib.pyplot"] = plt


def available(root=None):
    root = root or (REF_COPY if os.path.isdir(REF_COPY) else REF_SOURCE)
    return all(os.path.exists(os.path.join(root, m + ".py")) for m in _MODULES)


def verify_copy(root=REF_COPY):
    """sha256 of every module of ``oracle/_ref`` against the manifest written by build_ref.py."""
    with open(os.path.join(root, "MANIFEST.json")) as f:
        manifest = json.load(f)
    for name, digest in manifest["files"].items():
        with open(os.path.join(root, name), "rb") as f:
            if hashlib.sha256(f.read()).hexdigest() != digest:
                raise RuntimeError(f"oracle/_ref/{name} differs from the manifest: not the unmodified reference")
    return manifest


def load(root=None, with_main=False):
    """Import the reference modules (bare names ``data``, ``models``, ... as they import each other) from
    ``root`` (default: ``oracle/_ref`` if built, else ``/root/reference/model``).  Returns a namespace."""
    root = root or (REF_COPY if os.path.isdir(REF_COPY) else REF_SOURCE)
    key = (root, with_main)
    if key in _loaded:
        return _loaded[key]
    if not available(root):
        raise RuntimeError(f"the reference is not available at {root}: run `python oracle/build_ref.py` in the build "
                           f"container (needs /root/reference)")
    if os.path.abspath(root) == os.path.abspath(REF_COPY):
        verify_copy(root)
    _install_stubs()
    if root not in sys.path:
        sys.path.insert(0, root)
    import importlib
    ns = types.SimpleNamespace(root=root)
    for name in _MODULES + (("main",) if with_main else ()):
        mod = importlib.import_module(name)
        if os.path.dirname(os.path.abspath(mod.__file__)) != os.path.abspath(root):
            raise RuntimeError(f"module {name!r} resolved to {mod.__file__}, not to the reference under {root}")
        setattr(ns, name, mod)
    _loaded[key] = ns
    return ns


# ---------------------------------------------------------------------------------------------------
# building the reference's own objects around synthetic inputs
# ---------------------------------------------------------------------------------------------------
def ref_model(ref, sd, feat_dim, normalize_lang=False, dropout_rate=0.3):
    """The reference's ``CALModel`` (models.py:12-68) holding the fp32 NumPy / torch state dict ``sd``."""
    sd = {k: (torch.from_numpy(v) if isinstance(v, np.ndarray) else v.detach().cpu()) for k, v in sd.items()}
    model = ref.models.CALModel(pretrained_emb=sd["word_embedding.weight"].clone(), visual_input_dim=2 * feat_dim + 2,
                                emb_dim=sd["lang_fc.weight"].shape[0], hidden_size=sd["lstm.weight_hh_l0"].shape[1],
                                dropout_rate=dropout_rate, normalize_lang=normalize_lang)
    model.load_state_dict(sd)
    return model.eval()


def ref_dataset(ref, videos, queries, validate=True):
    """``data.CustomDataset`` built through ``__new__`` (its ``__init__`` reads feature files) and filled with the
    synthetic ``video_features`` / ``lang_features`` / ``num_segments_info``."""
    ds = ref.data.CustomDataset.__new__(ref.data.CustomDataset)
    ds.validate = validate
    ds.video_features = {v["name"]: dict(segment_features=v["segment_features"], context_features=v["context_features"],
                                         num_segments=v["num_segments"]) for v in videos}
    ds.num_segments_info = {v["name"]: v["num_segments"] for v in videos}
    ds.lang_features = {a: torch.from_numpy(queries["tokens"][i:i + 1]).long() for i, a in enumerate(queries["annot_id"])}
    annotations = {a: dict(video=videos[int(queries["video_idx"][i])]["name"], description="", times=queries["times"][i])
                   for i, a in enumerate(queries["annot_id"])}
    return ds, annotations


def ref_iters(ref, ds, videos, annotations, max_seg=None):
    """The reference's real ``VideoBatchSampler`` / ``LanguageBatchSampler`` + ``validate_collate`` loaders."""
    from torch.utils.data import DataLoader
    names = [v["name"] for v in videos]
    vit = DataLoader(ds, shuffle=False, collate_fn=ref.data.validate_collate,
                     batch_sampler=ref.data.VideoBatchSampler(names, ds.num_segments_info))
    lsamp = ref.data.LanguageBatchSampler(annotations, ds.num_segments_info)
    if max_seg is not None:   # the reference caps its table at n<7 (data.py:386): inject
        for n in range(7, max_seg + 1):
            lsamp.moments[n] = ref.utils.generate_moments(n)
    lit = DataLoader(ds, shuffle=False, collate_fn=ref.data.validate_collate, batch_sampler=lsamp)
    return vit, lit, lsamp.moments


class NullWriter:
    """Stand-in for the TensorBoard ``SummaryWriter`` the reference's ``Trainer`` writes to (main.py:69-77,
    107-117, 190-210): records every call so tests can compare the logged scalars."""

    def __init__(self):
        self.scalars, self.scalar_groups, self.figures = [], [], []

    def add_scalar(self, name, value, global_step=None):
        self.scalars.append((name, float(value), global_step))

    def add_scalars(self, name, values, global_step=None):
        self.scalar_groups.append((name, {k: float(v) for k, v in values.items()}, global_step))

    def add_figure(self, name, fig, global_step=None):
        self.figures.append(name)


# ---------------------------------------------------------------------------------------------------
# the corpus workload of bench.py through the reference's own evaluate() functions
# ---------------------------------------------------------------------------------------------------
def bank_dataset(ref, clip_emb, n_seg, tokens, q_video, times):
    """A ``data.CustomDataset`` whose videos are rows of a bank of CLIP EMBEDDINGS (the corpus workload of
    BASELINE configs[4] is generated at embedding level: 1 M videos of 8194-d features would be 197 GB).
    ``make_visual_features`` - feature assembly, not part of the timed scoring loop - hands the video's embedding rows
    through, and the model the caller passes to ``evaluate`` has ``visual_fc = nn.Identity()``; the text branch,
    ``evaluate.evaluate`` and ``evaluate_single.evaluate`` run unmodified."""
    n_videos = clip_emb.shape[0] // n_seg
    names = [f"v{i:07d}" for i in range(n_videos)]

    class BankDataset(ref.data.CustomDataset):
        def make_visual_features(self, video, start_t, end_t):
            return self.video_features[video]["embedding"][start_t:end_t + 1]

    ds = BankDataset.__new__(BankDataset)
    ds.validate = True
    emb = torch.as_tensor(clip_emb, dtype=torch.float32).reshape(n_videos, n_seg, -1)
    ds.video_features = {n: dict(embedding=emb[i], num_segments=n_seg) for i, n in enumerate(names)}
    ds.num_segments_info = {n: n_seg for n in names}
    annot = [f"a{i:07d}" for i in range(len(tokens))]
    ds.lang_features = {a: torch.as_tensor(tokens[i:i + 1]).long() for i, a in enumerate(annot)}
    annotations = {a: dict(video=names[int(q_video[i])], description="", times=times[i]) for i, a in enumerate(annot)}
    return ds, annotations, names
