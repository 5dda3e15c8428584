"""ctypes binding of the C-ABI library ``csrc/libvfr.so`` (declared in ``include/vfr.h``).

There is no fallback of any kind: if the shared library is missing or a call returns a non-zero
status this module raises.  Build the library with ``python -c "import __graft_entry__ as g;
g.build()"`` or ``make -C video-fragments-retrieval_b200/csrc``.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "csrc", "libvfr.so")
HEADER_PATH = os.path.join(os.path.dirname(_HERE), "include", "vfr.h")

VFR_TILE_Q = 128
VFR_TILE_C = 96
VFR_TOPK_MAX = 128
VFR_MAX_TAU = 16


class VfrError(RuntimeError):
    pass


_p = C.c_void_p
_i = C.c_int
_l = C.c_int64
_z = C.c_size_t
_f = C.c_float

class SearchPlan(C.Structure):
    """Mirror of ``vfr_search_plan`` (include/vfr.h)."""
    _fields_ = [("table", _p), ("vocab", _l), ("length_table", _p), ("emb", _i),
                ("lstm_fwd", _p), ("lstm_bwd", _p), ("hidden", _i),
                ("fc_w", _p), ("fc_b", _p), ("dim", _i), ("seq_len", _i),
                ("bank_packed", _p), ("vid_off", _p), ("mom_off", _p),
                ("n_videos", _l), ("n_max", _i), ("id_base", _l),
                ("tokens_dev", _p), ("q_emb", _p), ("q_packed", _p), ("text_ws", _p), ("topk_ws", _p),
                ("out_scores_dev", _p), ("out_ids_dev", _p), ("n_split", _i), ("max_queries", _l),
                ("engine", _i), ("bank_tc", _p), ("bank_clips", _p), ("uniform6", _i), ("q_tc", _p),
                ("text_engine", _i), ("text_tc", _p), ("n_clips", _l)]


# name -> (restype, argtypes); kept in the order of include/vfr.h
PROTOTYPES = {
    "vfr_last_error": (C.c_char_p, []),
    "vfr_version": (_i, []),
    "vfr_device_sms": (_i, []),
    "vfr_launch_count": (_l, []),
    "vfr_bank_pack_bytes": (_z, [_l, _i, _i]),
    "vfr_bank_pack": (_i, [_p, _p, _l, _i, _i, _p, _p]),
    "vfr_query_pack_bytes": (_z, [_l, _i]),
    "vfr_query_pack": (_i, [_p, _l, _i, _p, _p]),
    "vfr_score_full": (_i, [_p, _p, _p, _l, _i, _i, _p, _l, _p, _l, _p]),
    "vfr_score_own": (_i, [_p, _p, _i, _p, _l, _p, _p, _i, _p]),
    "vfr_score_count": (_i, [_p, _p, _l, _i, _i, _p, _l, _p, _i, _p, _p, _p, _i, _p]),
    "vfr_score_topk_bytes": (_z, [_l, _i]),
    "vfr_score_topk": (_i, [_p, _p, _p, _l, _i, _i, _p, _l, _i, _l, _p, _p, _p, _i, _p]),
    "vfr_topk_merge": (_i, [_p, _p, _i, _l, _i, _p, _p, _p]),
    "vfr_tc_bank_bytes": (_z, [_l]),
    "vfr_tc_bank_pack": (_i, [_p, _p, _l, _i, _i, _i, _p, _p]),
    "vfr_tc_query_bytes": (_z, [_l]),
    "vfr_tc_query_pack": (_i, [_p, _l, _i, _i, _p, _p]),
    "vfr_score_topk_tc_bytes": (_z, [_l, _l, _i]),
    "vfr_score_topk_tc": (_i, [_p, _p, _p, _p, _l, _i, _i, _i, _p, _p, _l, _i, _l, _p, _p, _p, _i, _p]),
    "vfr_score_full_tc": (_i, [_p, _p, _p, _p, _l, _i, _i, _i, _p, _p, _l, _p, _l, _p]),
    "vfr_sel_bank_bytes": (_z, [_l, _i]),
    "vfr_sel_bank_pack": (_i, [_p, _l, _i, _p, _p]),
    "vfr_sel_query_bytes": (_z, [_l, _i]),
    "vfr_sel_query_pack": (_i, [_p, _l, _i, _p, _l, _p, _p]),
    "vfr_sel_topk_bytes": (_z, [_l, _l, _i]),
    "vfr_sel_topk": (_i, [_p, _p, _p, _p, _l, _l, _i, _i, _p, _p, _l, _i, _l, _p, _p, _p, _i, _p]),
    "vfr_sel_flags": (_p, [_p, _l, _i]),
    "vfr_sel_topk_b16": (_i, [_p, _p, _p, _p, _l, _l, _i, _i, _p, _p, _l, _i, _l, _p, _p, _p, _i, _p]),
    "vfr_sel_tiles": (_l, [_l]),
    "vfr_sel_pool_levels": (_i, [_p, _i, _l, _i, _p, _i, _i, _p, _p]),
    "vfr_sel_count_levels": (_i, [_p, _l, _l, _i, _i, _p, _i, _p, _i, _p, _p]),
    "vfr_sel_pick_put": (_i, [_p, _l, _l, _i, _i, _p, _i, _p, _p, _i, _p]),
    "vfr_topk_block_bytes": (_z, [_l, _i]),
    "vfr_sel_refine_blocks": (_i, [_p, _p, _p, _l, _l, _i, _i, _p, _p, _l, _i, _l, _l, _p, _p, _i, _p]),
    "vfr_topk_merge_blocks": (_i, [_p, _i, _l, _l, _i, _p, _p, _p, _p, _p]),
    "vfr_sel_stats": (_i, [_p, _l, _l, _i, _i, _p, _i, _p, _p]),
    "vfr_sel_sample_rank": (_i, [_i, _l, _l]),
    "vfr_sel_sample_lists": (_i, [_l, _l, _i, _i]),
    "vfr_sel_sample_clips": (_l, [_l, _l, _i, _i, _i]),
    "vfr_sel_sample": (_i, [_p, _l, _i, _p, _l, _i, _p, _i, _p, _p, _p]),
    "vfr_sel_count_under": (_i, [_p, _l, _l, _i, _i, _p, _i, _p, _p, _p]),
    "vfr_sel_filter": (_i, [_p, _l, _i, _p, _l, _i, _p, _i, _l, _l, _i, _p]),
    "vfr_sel_bound_get": (_i, [_p, _l, _l, _i, _i, _p, _i, _p, _p]),
    "vfr_sel_bound_put": (_i, [_p, _l, _l, _i, _i, _p, _i, _p, _p]),
    "vfr_sel_refine": (_i, [_p, _p, _p, _l, _l, _i, _i, _p, _p, _l, _i, _l, _p, _p, _p, _i, _p]),
    "vfr_sel_refine_range": (_i, [_p, _p, _p, _l, _l, _i, _i, _p, _p, _l, _i, _l, _p, _p, _p, _i, _l, _l, _p]),
    "vfr_gt_select": (_i, [_p, _i, _p, _p, _i, _p, _i, _l, _p, _p, _p, _p, _p, _p]),
    "vfr_rank_order": (_i, [_p, _i, _p, _l, _i, _p, _p]),
    "vfr_single_metrics": (_i, [_p, _i, _p, _p, _i, _p, _i, _l, _p, _p, _p, _p, _p]),
    "vfr_linear": (_i, [_p, _l, _i, _i, _p, _p, _i, _i, _p, _i, _p]),
    "vfr_visual_embed": (_i, [_p, _l, _i, _p, _p, _i, _p, _p, _i, _p, _p, _p]),
    "vfr_visual_pack_bytes": (_z, [_i, _i, _i]),
    "vfr_visual_pack": (_i, [_p, _p, _p, _p, _i, _i, _i, _p, _p]),
    "vfr_visual_embed_tc_bytes": (_z, [_l, _l, _i, _i, _i, _i]),
    "vfr_visual_embed_tc": (_i, [_p, _l, _i, _p, _i, _i, _p, _p, _p]),
    "vfr_visual_embed_split": (_i, [_p, _p, _p, _l, _l, _i, _p, _i, _i, _p, _p, _p]),
    "vfr_lstm_pack_bytes": (_z, [_i, _i]),
    "vfr_lstm_pack": (_i, [_p, _p, _p, _p, _i, _i, _p, _p]),
    "vfr_text_embed_bytes": (_z, [_l, _i, _i, _i]),
    "vfr_text_embed": (_i, [_p, _l, _i, _p, _l, _p, _i, _p, _p, _i, _p, _p, _i, _p, _p, _p]),
    "vfr_tc_weight_bytes": (_z, [_i, _i]),
    "vfr_tc_weight_pack": (_i, [_p, _i, _i, _p, _p]),
    "vfr_linear_tc_bytes": (_z, [_l, _i]),
    "vfr_linear_tc": (_i, [_p, _l, _i, _l, _p, _p, _i, _i, _p, _l, _p, _p]),
    "vfr_text_pack_tc_bytes": (_z, [_i, _i, _i]),
    "vfr_text_pack_tc": (_i, [_p] * 10 + [_i, _i, _i, _p, _p]),
    "vfr_text_embed_tc_bytes": (_z, [_l, _i, _i, _i]),
    "vfr_text_embed_tc": (_i, [_p, _l, _i, _p, _l, _p, _i, _p, _i, _i, _p, _p, _p]),
    "vfr_segment_pool_bytes": (_z, [_l, _i, _i]),
    "vfr_segment_pool": (_i, [_p, _p, _l, _i, _i, _i, _p, _i, _p, _p, _p, _p]),
    "vfr_ranking_loss_bytes": (_z, [_i, _i, _i, _i]),
    "vfr_ranking_loss_fwd": (_i, [_p, _p, _p, _p, _p, _p, _i, _i, _i, _i, _i, _i, _f, _f, _p, _p, _p]),
    "vfr_ranking_loss_bwd": (_i, [_p, _p, _p, _p, _p, _p, _i, _i, _i, _i, _i, _i, _f, _f, _p, _p, _p, _p, _p, _p, _p]),
    "vfr_text_train_bytes": (_z, [_i, _i, _i, _i, _i]),
    "vfr_text_train_fwd": (_i, [_p, _i, _i, _p, _l, _p, _i, _p, _p, _p, _p, _i, _p, _p, _i, _p, _p, _p]),
    "vfr_text_train_bwd": (_i, [_p, _i, _i, _l, _i, _i, _p, _p, _i, _p, _i, _p, _p, _p, _p, _p, _p, _p, _p, _p, _p]),
    "vfr_text_train_flag": (_p, [_p]),
    "vfr_visual_train_fwd_bytes": (_z, [_l, _i, _i]),
    "vfr_visual_train_fwd": (_i, [_p, _l, _i, _p, _p, _i, _p, _p, _i, _p, _p, _p, _p]),
    "vfr_visual_train_bwd": (_i, [_p, _l, _i, _p, _i, _p, _p, _i, _p, _p, _p, _p, _p, _p, _p, _p]),
    "vfr_adam_step": (_i, [_p, _p, _p, _p, _p, _i, _l, _f, _f, _f, _f, _f, _p]),
    "vfr_grad_norms": (_i, [_p, _p, _i, _p, _p]),
    "vfr_sample_negatives": (_i, [_p, _i, _p, _p, _l, _l, _i, C.c_uint64, C.c_uint64, _p, _p]),
    "vfr_gather_clip_rows": (_i, [_p, _p, _p, _p, _p, _l, _i, _p, _p]),
    "vfr_moment_pool": (_i, [_p, _p, _p, _l, _i, _i, _p, _p]),
    "vfr_search_embed_device": (_i, [C.POINTER(SearchPlan), _p, _l, _p, _p]),
    "vfr_search_score_device": (_i, [C.POINTER(SearchPlan), _l, _i, _p, _p, _p]),
    "vfr_search_device": (_i, [C.POINTER(SearchPlan), _p, _l, _i, _p, _p, _p]),
    "vfr_search_host": (_i, [C.POINTER(SearchPlan), _p, _l, _i, _p, _p, _p]),
}

_lib = None


def load():
    """Load libvfr.so once; raises ``VfrError`` if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise VfrError(f"{LIB_PATH} not found: the CUDA library has not been built "
                       f"(run __graft_entry__.build()); there is no CPU fallback")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in PROTOTYPES.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(status, what):
    if status != 0:
        msg = load().vfr_last_error().decode(errors="replace")
        raise VfrError(f"{what} failed with status {status}: {msg}")


def call(name, *args):
    """Call an int-returning entry point and raise on a non-zero status."""
    check(getattr(load(), name)(*args), name)
