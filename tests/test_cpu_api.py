"""CPU-only checks of the host-side mirror of the reference interface (no GPU, no compute calls)."""
import inspect

import numpy as np
import pytest
import torch

import vfr_b200  # noqa: F401
from vfr_b200 import data as vdata
from vfr_b200 import evaluate as vev
from vfr_b200 import evaluate_single as vsingle
from vfr_b200 import main as vmain
from vfr_b200 import models, synth


def test_state_dict_keys_and_shapes_match_reference():
    # the reference's state_dict layout, 13,163,700 trainable parameters
    emb = torch.zeros(50, 100)
    m = models.CALModel(visual_input_dim=8194, pretrained_emb=emb, normalize_lang=True)
    sd = m.state_dict()
    want = {"visual_fc.0.weight": (500, 8194), "visual_fc.0.bias": (500,), "visual_fc.2.weight": (100, 500),
            "visual_fc.2.bias": (100,), "word_embedding.weight": (50, 100), "learnable_length.weight": (50, 1),
            "lstm.weight_ih_l0": (4000, 100), "lstm.weight_hh_l0": (4000, 1000), "lstm.bias_ih_l0": (4000,),
            "lstm.bias_hh_l0": (4000,), "lstm.weight_ih_l0_reverse": (4000, 100),
            "lstm.weight_hh_l0_reverse": (4000, 1000), "lstm.bias_ih_l0_reverse": (4000,),
            "lstm.bias_hh_l0_reverse": (4000,), "lang_fc.weight": (100, 2000), "lang_fc.bias": (100,)}
    assert {k: tuple(v.shape) for k, v in sd.items()} == want
    m2 = models.CALModel(visual_input_dim=8194, pretrained_emb=emb)
    assert sum(p.numel() for p in m2.parameters() if p.requires_grad) == 13163700
    assert not m2.word_embedding.weight.requires_grad
    bert = models.CALModel(visual_input_dim=8194)
    assert tuple(bert.lang_fc.weight.shape) == (100, 768) and not hasattr(bert, "lstm")


def test_signatures_mirror_reference():
    assert list(inspect.signature(models.CALModel.__init__).parameters) == [
        "self", "visual_input_dim", "pretrained_emb", "emb_dim", "hidden_size", "bert_emb", "dropout_rate", "normalize_lang"]
    assert list(inspect.signature(models.CALModel.forward).parameters) == ["self", "batch", "visual", "device", "bert"]
    assert list(inspect.signature(vev.evaluate).parameters) == [
        "model", "video_iterator", "lang_iterator", "annotations", "device", "preliminary", "model_types", "iou_thresholds"]
    assert list(inspect.signature(vsingle.evaluate).parameters) == [
        "model", "video_iterator", "lang_iterator", "annotations", "device", "model_types", "prior", "iou_thresholds"]
    assert list(inspect.signature(vmain.Trainer.ranking_loss).parameters) == [
        "self", "posit_emb", "intra_emb", "inter_emb", "lang_emb", "maskp", "maskn"]
    tr = vmain.Trainer()
    assert (tr.b, tr.lamb, tr.normalize_loss) == (0.1, 0.4, False)


def test_no_cpu_fallback():
    m = models.CALModel(visual_input_dim=10, pretrained_emb=torch.zeros(5, 100))
    with pytest.raises(RuntimeError, match="CUDA"):
        m(torch.zeros(3, 10))


def test_feature_assembly_and_collates_match_oracle():
    from oracle import cal_oracle as orc
    videos = synth.make_videos(3, 5, 16)
    ds = vdata.CustomDataset.__new__(vdata.CustomDataset)
    ds.validate = False
    ds.video_features = {v["name"]: v for v in videos}
    ds.lang_features = {"a": torch.arange(20).view(1, 20), "b": torch.arange(20).view(1, 20) + 1}
    v0, v1 = videos[0], videos[1]
    f = ds.make_visual_features(v0["name"], 1, 3)
    want = orc.make_visual_features(v0["segment_features"], v0["context_features"], v0["num_segments"], 1, 3)
    assert f.dtype == torch.float32 and torch.equal(f, want)
    np.testing.assert_array_equal(synth.clip_features(v0), ds.make_visual_features(v0["name"], 0, v0["num_segments"] - 1).numpy())
    s1 = ds[dict(annotation_id="a", video_pos=v0["name"], video_neg=v1["name"], start_t=0, end_t=1, start_tn=2, end_tn=2)]
    s2 = ds[dict(annotation_id="b", video_pos=v1["name"], video_neg=v0["name"], start_t=2, end_t=2, start_tn=0, end_tn=2)]
    batch = vdata.custom_collate([s1, s2])
    assert batch["maskp"].tolist() == [0, 0, 1] and batch["maskn"].tolist() == [0, 1, 1, 1]
    assert batch["posit"].shape == (3, 34) and batch["intra"].shape == (4, 34) and batch["lang"].shape == (2, 20)
    ds.validate = True
    item = vdata.validate_collate([ds[dict(video_pos=v0["name"], start_t=0, end_t=v0["num_segments"] - 1)]])
    assert item["annot_id"] == [] and item["feature"].shape[0] == v0["num_segments"]
    samp = vdata.LanguageBatchSampler({"a": dict(video=v0["name"])}, {})
    assert sorted(samp.moments) == list(range(7)) and len(samp.moments[6]) == 21
    assert vdata.tokenize("The Dog's ball, 2nd time!\n") == ["the", "dog", "ball", "2nd", "time"]


def test_get_metrics_matches_reference_reducer():
    from oracle import cal_oracle as orc
    rec = {1: [0, 1, 1], 10: [1, 1, 1], 100: [1, 1, 1], "MR": [5, 0, 2]}
    assert vev.get_metrics(rec) == orc.get_metrics(rec)
