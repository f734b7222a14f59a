"""Drop-in mirrors of utils/nms.py (reference lines cited per function); compute runs in libysb_postproc.so."""
import numpy as np
import torch

from .. import _lib
from ._common import cuda_device, nms_indices, stream_ptr, to_cuda_f32

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


def _soft_nms(boxes, scores, iou_type, iou_threshold, mode, sigma, thresh):
    assert isinstance(boxes, torch.Tensor) and isinstance(scores, torch.Tensor)
    assert boxes.shape[0] == scores.shape[0]
    kind = iou_type.lower() if isinstance(iou_type, str) else iou_type
    if kind not in ("iou", "giou", "diou", "ciou"):
        raise ValueError(f"Uknown paramemter: <{iou_type}>")
    m = boxes.shape[0]
    b = to_cuda_f32(boxes.reshape(-1, 4)) if m else torch.empty((0, 4), dtype=torch.float32, device=cuda_device())
    if m == 0:
        return torch.zeros(0, dtype=torch.bool, device=scores.device)
    s = to_cuda_f32(scores.reshape(-1), b.device)
    lib = _lib.load()
    ws = torch.empty(m, dtype=torch.float32, device=b.device)
    processed = torch.empty(m, dtype=torch.float32, device=b.device)
    with torch.cuda.device(b.device):
        _lib.check(lib.ysb_soft_nms(b.data_ptr(), s.data_ptr(), m, float(iou_threshold), _lib.IOU_KIND_IDS[kind], mode,
                                    float(sigma), ws.data_ptr(), ws.numel() * 4, processed.data_ptr(), stream_ptr()),
                   "ysb_soft_nms")
    return (processed > thresh).to(scores.device)


def gpu_linear_soft_nms(boxes, scores, iou_type, iou_threshold=0.3, thresh=0.001):
    """utils/nms.py:68-103 -- Tensor(M,4), Tensor(M,1), str -> bool Tensor(M): ``processed > thresh``.

    Each pick records its current score and multiplies every score whose float32 IoU flavour with the pick exceeds
    ``iou_threshold`` by ``1 - iou`` (the pick itself has IoU 1 and drops to 0).  ``'iou'`` raises IndexError in the
    reference as shipped (same indexing slip as gpu_nms); here it runs with the gpu_iou arithmetic.
    """
    return _soft_nms(boxes, scores, iou_type, iou_threshold, 0, 0.0, thresh)


def gpu_exponential_soft_nms(boxes, scores, iou_type, iou_threshold, sigmma=0.5, thresh=0.001):
    """utils/nms.py:106-140 -- as gpu_linear_soft_nms with the decay ``exp(-iou**2 / sigmma)``.

    A pick only decays itself by exp(-1/sigmma), so the reference keeps re-picking every box until its score underflows
    float32 and ``processed`` ends up holding the last (denormal) values: the returned mask is all False for any
    ``thresh`` above ~1e-38.  Reproduced as is (the loop runs on the device, ~104*sigmma picks per box).
    """
    return _soft_nms(boxes, scores, iou_type, iou_threshold, 1, sigmma, thresh)
