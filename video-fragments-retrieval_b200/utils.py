"""Pure helpers shared by every stage of the moment-scoring path.

Drop-in for the two hot-path helpers of the reference's ``model/utils.py``:

* ``generate_moments``  <- reference ``model/utils.py:71-75``
* ``get_iou``           <- reference ``model/utils.py:78-82``

plus the integer forms the CUDA kernels use (moment index <-> (start, end), threshold tables).
The moment enumeration order defined here is THE index map shared by the kernels
(csrc/vfr_common.cuh: ``moment_se``), the oracle and the metric reducers.
"""
import numpy as np

MAX_SEGMENTS = 32  # device-side cap (kernels keep one video's clip distances in registers/smem)


def generate_moments(num_segments):
    """All contiguous candidate moments of an ``num_segments``-clip video as inclusive
    ``(start, end)`` tuples: the ``n`` single clips first, then every ``start < end`` pair in
    lexicographic order -> ``n(n+1)/2`` entries (21 for 6 clips, 15 for 5, 465 for 30).
    Same order as reference ``model/utils.py:71-75``."""
    n = int(num_segments)
    out = [(c, c) for c in range(n)]
    for start in range(n):
        for end in range(start + 1, n):
            out.append((start, end))
    return out


def num_moments(num_segments):
    n = int(num_segments)
    return n * (n + 1) // 2


def moment_index(num_segments, start, end):
    """Inverse of ``generate_moments(n)[m] == (start, end)`` in O(1)."""
    n = int(num_segments)
    if start == end:
        return start
    # pairs (s, e>s) in lexicographic order: s' < s contribute (n-1-s') each
    return n + start * (n - 1) - start * (start - 1) // 2 + (end - start - 1)


def moment_table(num_segments):
    """``int32 [M, 2]`` array form of ``generate_moments``."""
    return np.asarray(generate_moments(num_segments), dtype=np.int32).reshape(-1, 2)


def get_iou(times, start_t, end_t):
    """Temporal IoU between the inclusive clip range ``[start_t, end_t]`` and each annotated
    ``[s, e]`` in ``times``; float64 array of length ``len(times)``
    (reference ``model/utils.py:78-82``)."""
    t = np.asarray(times)
    lo, hi = t[:, 0], t[:, 1]
    inter = np.clip(np.minimum(hi, end_t) - np.maximum(lo, start_t) + 1, 0, None)
    union = np.maximum(hi, end_t) - np.minimum(lo, start_t) + 1
    return inter / union


def iou_int(times, start_t, end_t):
    """Integer ``(intersection, union)`` of the same quantity (what the K5 kernel computes)."""
    t = np.asarray(times, dtype=np.int64)
    lo, hi = t[:, 0], t[:, 1]
    inter = np.clip(np.minimum(hi, end_t) - np.maximum(lo, start_t) + 1, 0, None)
    union = np.maximum(hi, end_t) - np.minimum(lo, start_t) + 1
    return inter, union


def threshold_table(thr, inclusive=False, max_len=2 * MAX_SEGMENTS):
    """``uint8 [max_len+1, max_len+1]`` table ``T[inter, union] = (inter/union > thr)`` evaluated
    in float64 exactly as NumPy evaluates ``get_iou(...) > thr`` (``>=`` when ``inclusive``, the
    ``Trainer.validate_epoch`` variant, reference ``model/main.py:161``).  The device kernel
    indexes this table with integer intersection/union, so the test is bit-identical to the
    reference for ANY python-float threshold (0.7 is not exactly representable)."""
    inter = np.arange(max_len + 1, dtype=np.int64).reshape(-1, 1)
    union = np.arange(max_len + 1, dtype=np.int64).reshape(1, -1)
    with np.errstate(divide="ignore", invalid="ignore"):
        ratio = inter / np.maximum(union, 1)
    tab = (ratio >= thr) if inclusive else (ratio > thr)
    tab[:, 0] = False
    return np.ascontiguousarray(tab.astype(np.uint8))
