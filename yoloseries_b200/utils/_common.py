"""Shared plumbing for the utils mirrors: device staging and the ctypes calls."""
import ctypes

import numpy as np
import torch

from .. import _lib


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
    """ysb_nms on CUDA tensors -> python list of kept indices (visiting order)."""
    lib = _lib.load()
    m = boxes.shape[0]
    ws_bytes = ctypes.c_size_t()
    _lib.check(lib.ysb_nms_workspace_bytes(m, ctypes.byref(ws_bytes)), "ysb_nms_workspace_bytes")
    dev = boxes.device
    ws = torch.empty(max(ws_bytes.value, 1), dtype=torch.uint8, device=dev)
    cap = m if max_keep <= 0 else min(m, max_keep)
    keep = torch.empty(max(cap, 1), dtype=torch.int32, device=dev)
    cnt = torch.zeros(1, dtype=torch.int32, device=dev)
    with torch.cuda.device(dev):
        _lib.check(lib.ysb_nms(boxes.data_ptr(), scores.data_ptr(), m, float(iou_threshold), cmp, iou_kind,
                               int(max_keep), ws.data_ptr(), ws.numel(), keep.data_ptr(), cnt.data_ptr(),
                               stream_ptr()), "ysb_nms")
    n = int(cnt.item())
    return keep[:n].cpu().tolist()
