"""vfr_b200 - B200-native moment-scoring hot path of mariyashcheg/video-fragments-retrieval.

Module names mirror the reference's ``model/`` directory so that a user of the reference finds
the same entry points: ``models.CALModel``, ``evaluate.evaluate``, ``evaluate_single.evaluate``,
``main.Trainer.ranking_loss``, ``utils.generate_moments`` / ``utils.get_iou``,
``data.CustomDataset`` feature assembly.  All device work goes through the C-ABI library
``csrc/libvfr.so`` (hand-written sm_100a CUDA) loaded by ``_lib``; there is no CPU fallback.
"""
from . import utils  # noqa: F401  (pure python, importable without the CUDA library)

__all__ = ["utils", "synth", "_lib", "ops", "models", "evaluate", "evaluate_single", "main", "data",
           "retrieval"]
