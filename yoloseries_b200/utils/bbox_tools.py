"""Drop-in mirrors of the IoU routines of utils/bbox_tools.py; compute runs in libysb_postproc.so (forward only)."""
import numpy as np
import torch

from .. import _lib
from ._common import stream_ptr, to_cuda_f32

__all__ = ["numba_iou", "gpu_iou", "gpu_Giou", "gpu_DIoU", "gpu_CIoU"]


def _no_grad_only(*tensors):
    if any(isinstance(t, torch.Tensor) and t.requires_grad for t in tensors):
        raise NotImplementedError(
            "the CUDA IoU kernels are forward-only; loss functions that differentiate through gpu_iou/gpu_CIoU "
            "(loss/yolov5_loss.py:110, loss/yolox_loss.py:133) are outside the post-processing path (SURVEY.md 8f rank 4)")


def numba_iou(bbox1, bbox2):
    """utils/bbox_tools.py:12-35 -- (M,4) f32, (N,4) f32 ndarrays -> (M,N) float64 ndarray (no clamp, NaN for 0/0)."""
    b1 = to_cuda_f32(np.asarray(bbox1).reshape(-1, 4))
    b2 = to_cuda_f32(np.asarray(bbox2).reshape(-1, 4), b1.device)
    out = torch.empty((b1.shape[0], b2.shape[0]), dtype=torch.float64, device=b1.device)
    with torch.cuda.device(b1.device):
        _lib.check(_lib.load().ysb_pairwise_iou(b1.data_ptr(), b1.shape[0], b2.data_ptr(), b2.shape[0],
                                                _lib.IOU_NUMBA_F64MIX, out.data_ptr(), stream_ptr()), "ysb_pairwise_iou")
    return out.cpu().numpy()


def gpu_iou(bbox1, bbox2):
    """utils/bbox_tools.py:164-190 -- (N,4), (M,4) tensors -> (N,M) float32 tensor on the inputs' device."""
    _no_grad_only(bbox1, bbox2)
    b1, b2 = to_cuda_f32(bbox1.reshape(-1, 4)), to_cuda_f32(bbox2.reshape(-1, 4))
    out = torch.empty((b1.shape[0], b2.shape[0]), dtype=torch.float32, device=b1.device)
    with torch.cuda.device(b1.device):
        _lib.check(_lib.load().ysb_pairwise_iou(b1.data_ptr(), b1.shape[0], b2.data_ptr(), b2.shape[0], _lib.IOU_F32,
                                                out.data_ptr(), stream_ptr()), "ysb_pairwise_iou")
    return out


def _rowwise(kind, bbox1, bbox2):
    assert isinstance(bbox1, torch.Tensor)
    assert isinstance(bbox2, torch.Tensor)
    assert bbox1.shape[-1] == bbox2.shape[-1] == 4
    assert bbox1.device == bbox2.device
    _no_grad_only(bbox1, bbox2)
    b1, b2 = to_cuda_f32(bbox1.reshape(-1, 4)), to_cuda_f32(bbox2.reshape(-1, 4))
    if b1.shape[0] not in (1, b2.shape[0]):
        raise RuntimeError(f"The size of tensor a ({b1.shape[0]}) must match the size of tensor b ({b2.shape[0]})")
    out = torch.empty((b2.shape[0],), dtype=torch.float32, device=b2.device)
    with torch.cuda.device(b2.device):
        _lib.check(_lib.load().ysb_elementwise_iou(b1.data_ptr(), b1.shape[0], b2.data_ptr(), b2.shape[0], kind,
                                                   out.data_ptr(), stream_ptr()), "ysb_elementwise_iou")
    return out


def gpu_Giou(bbox1, bbox2):
    """utils/bbox_tools.py:193-229 -- (N|1,4), (N,4) -> (N,).  The reference asserts x2>=x1, y2>=y1 (:201-202)."""
    assert (bbox1[:, [2, 3]] >= bbox1[:, [0, 1]]).bool().all()
    assert (bbox2[:, [2, 3]] >= bbox2[:, [0, 1]]).bool().all()
    return _rowwise(_lib.GIOU, bbox1, bbox2)


def gpu_DIoU(bbox1, bbox2):
    """utils/bbox_tools.py:232-283 -- (N|1,4), (N,4) -> (N,), clamped to [-1, 1]."""
    assert (bbox1[:, [2, 3]] >= bbox1[:, [0, 1]]).bool().all()
    assert (bbox2[:, [2, 3]] >= bbox2[:, [0, 1]]).bool().all()
    return _rowwise(_lib.DIOU, bbox1, bbox2)


def gpu_CIoU(bbox1, bbox2):
    """utils/bbox_tools.py:286-339 -- (N,4), (N,4) -> (N,) (a 0-d tensor when N == 1, like ``.squeeze()``)."""
    out = _rowwise(_lib.CIOU, bbox1, bbox2)
    return out.squeeze()
