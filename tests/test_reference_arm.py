"""The reference arm of bench.py: oracle/_ref is a verbatim, sha256-checked copy of the reference's modules
(oracle/build_ref.py), and one tiny step of the unmodified evaluate.evaluate + evaluate_single.evaluate runs through
the harness.  CPU only; skipped when neither oracle/_ref nor /root/reference exists."""
import os

import pytest

from oracle import ref_harness

pytestmark = pytest.mark.skipif(not ref_harness.available(), reason="the reference is not available here")


def test_ref_copy_is_verbatim():
    if not os.path.isdir(ref_harness.REF_COPY):
        pytest.skip("oracle/_ref not built")
    manifest = ref_harness.verify_copy()
    assert sorted(manifest["files"]) == ["data.py", "evaluate.py", "evaluate_single.py", "main.py", "models.py", "utils.py"]
    if os.path.isdir(ref_harness.REF_SOURCE):
        import hashlib
        for name, digest in manifest["files"].items():
            with open(os.path.join(ref_harness.REF_SOURCE, name), "rb") as f:
                assert hashlib.sha256(f.read()).hexdigest() == digest


def test_reference_arm_runs_the_unmodified_evaluate():
    import bench
    arm = bench.ReferenceArm(64 * 8, n_q=3, n_v=8)
    assert arm.kind == "reference"
    corpus, single = arm.step()
    assert set(corpus) == {"model, IoU=0.5", "model, IoU=0.7"} and set(corpus["model, IoU=0.5"]) == {"R@1", "R@10", "R@100", "MR"}
    assert set(single["model"]) == {"Rank@1", "Rank@5", "Rank@10", "mIoU"}
    assert arm.pairs_per_step == 3 * 8 * 21
    # the functions that ran are the files of oracle/_ref (or of /root/reference itself)
    assert os.path.dirname(os.path.abspath(arm.ref.evaluate.__file__)) == os.path.abspath(arm.ref.root)
