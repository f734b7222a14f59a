"""ctypes front-end for oracle/oracle_nms.c (TEST INFRASTRUCTURE ONLY)."""
import ctypes
import os

import numpy as np

from . import build_oracle

_LIB = None


def load_library():
    global _LIB
    if _LIB is None:
        path = build_oracle.LIB
        if not os.path.exists(path) or os.path.getmtime(path) < os.path.getmtime(build_oracle.SRC):
            path = build_oracle.build()
        lib = ctypes.CDLL(path)
        f32p = ctypes.POINTER(ctypes.c_float)
        f64p = ctypes.POINTER(ctypes.c_double)
        i32p = ctypes.POINTER(ctypes.c_int32)
        u8p = ctypes.POINTER(ctypes.c_uint8)
        lib.orc_numba_iou.argtypes = [f32p, ctypes.c_int, f32p, ctypes.c_int, f64p]
        lib.orc_numba_iou.restype = None
        lib.orc_numba_nms.argtypes = [f32p, f32p, ctypes.c_int, ctypes.c_double, ctypes.c_int, i32p]
        lib.orc_numba_nms.restype = ctypes.c_int
        lib.orc_postprocess_count.argtypes = [f32p, f32p, ctypes.c_int, i32p, ctypes.c_int, ctypes.c_double,
                                              f32p, ctypes.c_int, u8p, f32p]
        lib.orc_postprocess_count.restype = None
        lib.orc_gpu_iou.argtypes = [f32p, ctypes.c_int, f32p, ctypes.c_int, f32p]
        lib.orc_gpu_iou.restype = None
        lib.orc_gpu_nms_iou.argtypes = [f32p, f32p, ctypes.c_int, ctypes.c_float, ctypes.c_int, i32p]
        lib.orc_gpu_nms_iou.restype = ctypes.c_int
        _LIB = lib
    return _LIB


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def _ptr(a, ct):
    return a.ctypes.data_as(ctypes.POINTER(ct))


def numba_iou(b1, b2):
    """utils/bbox_tools.py:12-35 -> (M, N) float64."""
    b1, b2 = _f32(b1).reshape(-1, 4), _f32(b2).reshape(-1, 4)
    out = np.empty((b1.shape[0], b2.shape[0]), dtype=np.float64)
    load_library().orc_numba_iou(_ptr(b1, ctypes.c_float), b1.shape[0], _ptr(b2, ctypes.c_float), b2.shape[0],
                                 _ptr(out, ctypes.c_double))
    return out


def numba_nms(boxes, scores, iou_threshold, max_keep=0):
    """utils/nms.py:10-27 -> list[int] (descending score, ties by lower index)."""
    boxes, scores = _f32(boxes).reshape(-1, 4), _f32(scores).reshape(-1)
    assert boxes.shape[0] == scores.shape[0]
    m = boxes.shape[0]
    keep = np.empty(max(m, 1), dtype=np.int32)
    n = load_library().orc_numba_nms(_ptr(boxes, ctypes.c_float), _ptr(scores, ctypes.c_float), m,
                                     float(iou_threshold), int(max_keep), _ptr(keep, ctypes.c_int32))
    return keep[:n].tolist()


def gpu_iou_f32(b1, b2):
    """utils/bbox_tools.py:164-190 -> (N, M) float32."""
    b1, b2 = _f32(b1).reshape(-1, 4), _f32(b2).reshape(-1, 4)
    out = np.empty((b1.shape[0], b2.shape[0]), dtype=np.float32)
    load_library().orc_gpu_iou(_ptr(b1, ctypes.c_float), b1.shape[0], _ptr(b2, ctypes.c_float), b2.shape[0],
                               _ptr(out, ctypes.c_float))
    return out


def gpu_nms_iou(boxes, scores, iou_threshold, max_keep=0):
    """Intended behaviour of utils/nms.py:30-65 with iou_type='iou' (float32 IoU, strict '>')."""
    boxes, scores = _f32(boxes).reshape(-1, 4), _f32(scores).reshape(-1)
    m = boxes.shape[0]
    keep = np.empty(max(m, 1), dtype=np.int32)
    n = load_library().orc_gpu_nms_iou(_ptr(boxes, ctypes.c_float), _ptr(scores, ctypes.c_float), m,
                                       float(iou_threshold), int(max_keep), _ptr(keep, ctypes.c_int32))
    return keep[:n].tolist()


def postprocess_count(boxes_off, scores, keep, iou_threshold, raw_boxes=None, merge=False):
    """Count filter (and RetinaNet merge) of trainer/eval_yolov5.py:306-315 / eval_retinanet.py:342-352.

    Returns (pass_flags bool (K,), merged float32 (K,4) or None).
    """
    boxes_off, scores = _f32(boxes_off).reshape(-1, 4), _f32(scores).reshape(-1)
    keep = np.ascontiguousarray(keep, dtype=np.int32)
    k = keep.shape[0]
    flags = np.zeros(max(k, 1), dtype=np.uint8)
    merged = np.zeros((max(k, 1), 4), dtype=np.float32)
    raw = _f32(raw_boxes).reshape(-1, 4) if raw_boxes is not None else boxes_off
    load_library().orc_postprocess_count(_ptr(boxes_off, ctypes.c_float), _ptr(scores, ctypes.c_float),
                                         boxes_off.shape[0], _ptr(keep, ctypes.c_int32), k, float(iou_threshold),
                                         _ptr(raw, ctypes.c_float), int(bool(merge)), _ptr(flags, ctypes.c_uint8),
                                         _ptr(merged, ctypes.c_float))
    return flags[:k].astype(bool), (merged[:k] if merge else None)
