"""CPU-side checks of the drop-in boundary: the C-ABI library loads and exports every symbol that
include/vfr.h declares, and the ctypes prototypes cover exactly that set.  No compute calls."""
import ctypes
import os
import re

import pytest

import vfr_b200  # noqa: F401
from vfr_b200 import _lib


def _declared():
    src = open(_lib.HEADER_PATH).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(vfr_[a-z0-9_]+)\s*\(", src)))


def test_header_symbols_exported():
    if not os.path.exists(_lib.LIB_PATH):
        import __graft_entry__
        __graft_entry__.build()
    lib = ctypes.CDLL(_lib.LIB_PATH)
    names = _declared()
    assert len(names) >= 10
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/vfr.h but not exported by libvfr.so"


def test_prototypes_cover_header():
    assert sorted(_lib.PROTOTYPES) == _declared()


def test_argument_errors_are_reported_not_thrown():
    lib = _lib.load()
    assert lib.vfr_version() >= 100
    assert lib.vfr_bank_pack_bytes(10, 64, 100) == 0          # n_max > 32: unsupported
    assert lib.vfr_bank_pack_bytes(16, 6, 100) == (5 * 20 * 96 + 96) * 4
    rc = lib.vfr_bank_pack(None, None, 10, 6, 100, None, None)
    assert rc == -1 and b"null" in lib.vfr_last_error()
    with pytest.raises(_lib.VfrError):
        _lib.call("vfr_topk_merge", None, None, 1, 1, 1, None, None, None)


def test_ops_refuse_cpu_tensors():
    import torch
    from vfr_b200 import ops
    with pytest.raises(_lib.VfrError):
        ops.pack_queries(torch.zeros(4, 100))
