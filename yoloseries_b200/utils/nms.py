"""Drop-in mirrors of utils/nms.py (reference lines cited per function); compute runs in libysb_postproc.so."""
import numpy as np
import torch

from .. import _lib
from ._common import nms_indices, to_cuda_f32

__all__ = ["numba_nms", "gpu_nms", "gpu_linear_soft_nms", "gpu_exponential_soft_nms"]


def numba_nms(boxes, scores, iou_threshold, max_keep=0):
    """utils/nms.py:10-27 -- (M,4) f32 ndarray, (M,) f32 ndarray, float -> list[int].

    Same order (descending score, ties by lower index), ``>=`` test on the float64-mixed IoU of numba_iou, zero scores
    never kept.  ``max_keep`` (extension, default 0 = run to exhaustion like the reference) stops after that many keeps.
    Inputs are never modified.  Scores must be >= 0 (the reference's ``while sum > 0`` loop is only meaningful then).
    """
    assert boxes.shape[0] == scores.shape[0]
    if boxes.shape[0] == 0:
        return []
    b = to_cuda_f32(np.asarray(boxes).reshape(-1, 4) if isinstance(boxes, np.ndarray) else boxes.reshape(-1, 4))
    s = to_cuda_f32(np.asarray(scores).reshape(-1) if isinstance(scores, np.ndarray) else scores.reshape(-1), b.device)
    return nms_indices(b, s, iou_threshold, _lib.IOU_NUMBA_F64MIX, _lib.CMP_GE, max_keep)


def gpu_nms(boxes, scores, iou_type, iou_threshold, max_keep=0):
    """utils/nms.py:30-65 -- Tensor(M,4), Tensor(M) or (M,1), str, float -> list[int].

    Greedy loop with ``iou.gt(threshold)`` on the float32 IoU flavour named by ``iou_type``.  ``'iou'`` raises
    IndexError in the reference as shipped; here it runs with the intended gpu_iou arithmetic.
    """
    assert isinstance(boxes, torch.Tensor) and isinstance(scores, torch.Tensor)
    assert boxes.shape[0] == scores.shape[0]
    kind = iou_type.lower()
    if kind not in ("iou", "giou", "diou", "ciou"):
        raise ValueError(f"Uknown paramemter: <{iou_type}>")
    if boxes.shape[0] == 0:
        return []
    b = to_cuda_f32(boxes.reshape(-1, 4))
    s = to_cuda_f32(scores.reshape(-1), b.device)
    return nms_indices(b, s, iou_threshold, _lib.IOU_KIND_IDS[kind], _lib.CMP_GT, max_keep)


def gpu_linear_soft_nms(boxes, scores, iou_type, iou_threshold=0.3, thresh=0.001):
    """utils/nms.py:68-103 -- no caller anywhere in the reference; listed as a 'next' row."""
    raise NotImplementedError("soft-NMS is outside the round-1 hot-path scope (SURVEY.md section 8f, rank 4)")


def gpu_exponential_soft_nms(boxes, scores, iou_type, iou_threshold, sigmma=0.5, thresh=0.001):
    """utils/nms.py:106-140 -- see gpu_linear_soft_nms."""
    raise NotImplementedError("soft-NMS is outside the round-1 hot-path scope (SURVEY.md section 8f, rank 4)")
