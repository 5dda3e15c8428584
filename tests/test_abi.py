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


def test_sel_host_planning_functions_without_a_gpu():
    """The host-only planning entry points of the filter + refine engine (no kernel is launched): candidate lists per
    query, sample sizes, workspace sizes."""
    lib = _lib.load()
    clips_1m = 6_000_000
    # one list per query when a CTA serves two query tiles and the bank is not split; two (tile parities) otherwise
    assert lib.vfr_sel_sample_lists(37888, clips_1m, 1, 100) == 1
    assert lib.vfr_sel_sample_lists(100, clips_1m, 1, 100) == 2
    assert lib.vfr_sel_sample_lists(300, clips_1m, 3, 100) == 3
    # the sample: whole tiles of 256 clips per list, more for larger k, at most 1/32 of a list's share of the bank,
    # none for banks below 128 tiles per list
    for k, want_tiles in ((1, 64), (10, 128), (100, 512)):
        n = lib.vfr_sel_sample_clips(37888, clips_1m, k, 1, 100)
        assert n == want_tiles * 256, (k, n)
    assert lib.vfr_sel_sample_clips(37888, 750_000, 100, 1, 100) == 64 * 256          # 2930 tiles: 1/32 rule
    assert lib.vfr_sel_sample_clips(37888, 127 * 256, 100, 1, 100) == 0
    assert lib.vfr_sel_sample_clips(100, clips_1m, 100, 1, 100) == 2 * 256 * 256     # two lists, 1/32 of 11 719 tiles each
    # workspace: lists dominate (1024 keys of 8 bytes per query and list), grows with the batch
    small, large = lib.vfr_sel_topk_bytes(1000, clips_1m, 1), lib.vfr_sel_topk_bytes(37888, clips_1m, 1)
    assert 1000 * 1024 * 8 <= small < large and large >= 37888 * 1024 * 8
    assert lib.vfr_sel_tiles(clips_1m) == 23438
