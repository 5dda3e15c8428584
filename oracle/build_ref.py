"""Recipe that makes the UNMODIFIED reference importable on the GPU box - TEST INFRASTRUCTURE ONLY.

    python oracle/build_ref.py          # in the build container, where /root/reference exists

The reference (mariyashcheg/video-fragments-retrieval) is 1.3 k lines of plain Python under
``/root/reference/model``: there is nothing to compile and nothing to ``pip install`` (no setup.py /
pyproject).  "Building" it therefore means placing byte-identical copies of its six modules where the GPU box
can import them: ``oracle/_ref/`` - git-ignored (reference sources never enter this repository's history),
NOT gpurun-ignored (so the directory travels with the snapshot like our own built ``.so``).  A manifest with
the sha256 of every file is written next to them and ``oracle/ref_harness.load()`` re-checks it at import
time, so what ``bench.py --impl reference`` times is provably the reference's own code.

``__graft_entry__.build()`` calls ``build()`` below whenever ``/root/reference`` is present.
"""
import hashlib
import json
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = "/root/reference/model"
DST = os.path.join(HERE, "_ref")
FILES = ("data.py", "evaluate.py", "evaluate_single.py", "main.py", "models.py", "utils.py")


def _sha(path):
    with open(path, "rb") as f:
        return hashlib.sha256(f.read()).hexdigest()


def build(src=SRC, dst=DST):
    """Copy the reference modules verbatim; returns the manifest dict (or None when the source is absent and
    a previously built copy is kept)."""
    if not os.path.isdir(src):
        return None
    os.makedirs(dst, exist_ok=True)
    manifest = {"source": src, "files": {}}
    for name in FILES:
        shutil.copyfile(os.path.join(src, name), os.path.join(dst, name))
        os.chmod(os.path.join(dst, name), 0o644)
        manifest["files"][name] = _sha(os.path.join(dst, name))
        assert manifest["files"][name] == _sha(os.path.join(src, name))
    with open(os.path.join(dst, "MANIFEST.json"), "w") as f:
        json.dump(manifest, f, indent=1)
    return manifest


if __name__ == "__main__":
    m = build()
    if m is None:
        print(f"{SRC} not present: nothing built", file=sys.stderr)
        sys.exit(1)
    for k, v in m["files"].items():
        print(f"{v[:16]}  oracle/_ref/{k}")
