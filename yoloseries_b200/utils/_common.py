"""Shared plumbing for the utils mirrors: device staging and the ctypes calls."""
import ctypes

import numpy as np
import torch

from .. import _lib, _ops


def cuda_device():
    if not torch.cuda.is_available():
        raise RuntimeError("yoloseries_b200 needs a CUDA device: the engine has no CPU fallback")
    return torch.device("cuda", torch.cuda.current_device())


def stream_ptr():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def to_cuda_f32(x, device=None):
    """numpy array / tensor -> contiguous float32 CUDA tensor (a copy only when needed)."""
    if isinstance(x, np.ndarray):
        x = torch.from_numpy(np.ascontiguousarray(x, dtype=np.float32))
    t = x.detach()
    if t.device.type != "cuda":
        t = t.to(device or cuda_device())
    t = t.to(torch.float32).contiguous()
    if t.data_ptr() % 16:   # an offset view of a larger buffer: the kernels read boxes as float4
        t = t.clone()
    return t


def nms_indices(boxes, scores, iou_threshold, iou_kind, cmp, max_keep=0):
    """torch.ops.ysb.nms (ysb_nms behind the torch extension) on CUDA tensors -> python list of kept indices (visiting
    order).  Workspace and outputs come from torch's caching allocator inside the operator."""
    keep, cnt = _ops.load().nms(boxes, scores, float(iou_threshold), int(cmp), int(iou_kind), int(max_keep))
    return keep[: int(cnt.item())].cpu().tolist()
