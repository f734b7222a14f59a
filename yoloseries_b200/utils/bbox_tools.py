"""Drop-in mirrors of the IoU routines of utils/bbox_tools.py: torch.ops.ysb.* (the thin torch extension,
csrc/torch_adapter.cpp) over the C ABI of libysb_postproc.so.

The row-wise GIoU / DIoU / CIoU are differentiable (torch.autograd.Function over ysb_elementwise_iou /
ysb_elementwise_iou_backward) because the reference's losses differentiate through them (loss/yolov5_loss.py:110,
loss/yolov7_loss.py:130, loss/yolov8_loss.py:306, loss/loss.py:109).  The pairwise gpu_iou is differentiable too
(ysb_pairwise_iou_backward): the label assignment of loss/yolox_loss.py:133 and loss/yolov7_loss.py:312 calls it on
predictions that require grad with grad mode ON (`torch.no_grad()` at yolox_loss.py:92 is a bare statement, not a
decorator), so the result must carry a grad_fn exactly like the reference's torch expression does."""
import numpy as np
import torch

from .. import _lib, _ops
from ._common import stream_ptr, to_cuda_f32

__all__ = ["numba_iou", "gpu_iou", "gpu_Giou", "gpu_DIoU", "gpu_CIoU"]


def numba_iou(bbox1, bbox2):
    """utils/bbox_tools.py:12-35 -- (M,4) f32, (N,4) f32 ndarrays -> (M,N) float64 ndarray (no clamp, NaN for 0/0)."""
    b1 = to_cuda_f32(np.asarray(bbox1).reshape(-1, 4))
    b2 = to_cuda_f32(np.asarray(bbox2).reshape(-1, 4), b1.device)
    return _ops.load().pairwise_iou(b1, b2, _lib.IOU_NUMBA_F64MIX).cpu().numpy()


def _pairwise_forward(b1, b2):
    return _ops.load().pairwise_iou(b1, b2, _lib.IOU_F32)


class _PairwiseIoU(torch.autograd.Function):
    """forward: ysb_pairwise_iou (float32 flavour); backward: ysb_pairwise_iou_backward (first order only)."""

    @staticmethod
    def forward(ctx, bbox1, bbox2):
        b1 = to_cuda_f32(bbox1.reshape(-1, 4))
        b2 = to_cuda_f32(bbox2.reshape(-1, 4), b1.device)
        ctx.save_for_backward(b1, b2)
        ctx.meta = (bbox1.shape, bbox1.dtype, bbox1.device, bbox2.shape, bbox2.dtype, bbox2.device)
        return _pairwise_forward(b1, b2)

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, grad_out):
        b1, b2 = ctx.saved_tensors
        shape1, dtype1, dev1, shape2, dtype2, dev2 = ctx.meta
        need1, need2 = ctx.needs_input_grad[0], ctx.needs_input_grad[1]
        g = to_cuda_f32(grad_out.reshape(b1.shape[0], b2.shape[0]), b1.device)
        g1, g2 = _ops.load().pairwise_iou_backward(b1, b2, g, need1, need2)
        if need1:
            g1 = g1.reshape(shape1).to(device=dev1, dtype=dtype1)
        if need2:
            g2 = g2.reshape(shape2).to(device=dev2, dtype=dtype2)
        return g1, g2


def gpu_iou(bbox1, bbox2):
    """utils/bbox_tools.py:164-190 -- (N,4), (M,4) tensors -> (N,M) float32 tensor on the inputs' device; differentiable
    like the reference's torch expression (callers: loss/yolox_loss.py:133, loss/yolov7_loss.py:312)."""
    if torch.is_grad_enabled() and (bbox1.requires_grad or bbox2.requires_grad):
        return _PairwiseIoU.apply(bbox1, bbox2)
    b1 = to_cuda_f32(bbox1.reshape(-1, 4))
    return _pairwise_forward(b1, to_cuda_f32(bbox2.reshape(-1, 4), b1.device))


def _rowwise_forward(kind, b1, b2):
    return _ops.load().elementwise_iou(b1, b2, kind)


class _RowwiseIoU(torch.autograd.Function):
    """forward: ysb_elementwise_iou; backward: ysb_elementwise_iou_backward (first order only)."""

    @staticmethod
    def forward(ctx, bbox1, bbox2, kind):
        b1 = to_cuda_f32(bbox1.reshape(-1, 4))
        b2 = to_cuda_f32(bbox2.reshape(-1, 4), b1.device)
        ctx.save_for_backward(b1, b2)
        ctx.kind = kind
        ctx.meta = (bbox1.shape, bbox1.dtype, bbox1.device, bbox2.shape, bbox2.dtype, bbox2.device)
        return _rowwise_forward(kind, b1, b2)

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, grad_out):
        b1, b2 = ctx.saved_tensors
        shape1, dtype1, dev1, shape2, dtype2, dev2 = ctx.meta
        need1, need2 = ctx.needs_input_grad[0], ctx.needs_input_grad[1]
        g = to_cuda_f32(grad_out.reshape(-1), b2.device)
        g1, g2 = _ops.load().elementwise_iou_backward(b1, b2, ctx.kind, g, need1, need2)
        if need1:
            g1 = g1.reshape(shape1).to(device=dev1, dtype=dtype1)
        if need2:
            g2 = g2.reshape(shape2).to(device=dev2, dtype=dtype2)
        return g1, g2, None


def _rowwise(kind, bbox1, bbox2):
    assert isinstance(bbox1, torch.Tensor)
    assert isinstance(bbox2, torch.Tensor)
    assert bbox1.shape[-1] == bbox2.shape[-1] == 4
    assert bbox1.device == bbox2.device
    n1, n2 = bbox1.reshape(-1, 4).shape[0], bbox2.reshape(-1, 4).shape[0]
    if n1 not in (1, n2):
        raise RuntimeError(f"The size of tensor a ({n1}) must match the size of tensor b ({n2})")
    if torch.is_grad_enabled() and (bbox1.requires_grad or bbox2.requires_grad):
        return _RowwiseIoU.apply(bbox1, bbox2, kind)
    b1, b2 = to_cuda_f32(bbox1.reshape(-1, 4)), to_cuda_f32(bbox2.reshape(-1, 4))
    return _rowwise_forward(kind, b1, b2)


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
