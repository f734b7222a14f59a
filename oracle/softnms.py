"""TEST INFRASTRUCTURE ONLY -- CPU restatement (numpy float32) of the reference's torch IoU flavours, soft-NMS and the
letterbox undo.  Never imported by the product path (yoloseries_b200/); see oracle/__init__.py.

Pinned against tests/golden/utils_nms_iou.npz (outputs of the reference itself, oracle/gen_golden.py::utils_case):
  giou / diou / ciou   utils/bbox_tools.py:193-339   row-wise, box1 broadcast when it has one row
  linear_soft_nms      utils/nms.py:68-103
  exponential_soft_nms utils/nms.py:106-140
  undo_letterbox       val_yolov5.py:166-172
"""
import numpy as np

F = np.float32


def _parts(b1, b2):
    b1 = np.asarray(b1, dtype=F).reshape(-1, 4)
    b2 = np.asarray(b2, dtype=F).reshape(-1, 4)
    a1 = (b1[:, 2] - b1[:, 0]) * (b1[:, 3] - b1[:, 1])
    a2 = (b2[:, 2] - b2[:, 0]) * (b2[:, 3] - b2[:, 1])
    iw = np.maximum(np.minimum(b1[:, 2], b2[:, 2]) - np.maximum(b1[:, 0], b2[:, 0]), F(0))
    ih = np.maximum(np.minimum(b1[:, 3], b2[:, 3]) - np.maximum(b1[:, 1], b2[:, 1]), F(0))
    inter = iw * ih
    union = a1 + a2 - inter
    cw = np.maximum(b1[:, 2], b2[:, 2]) - np.minimum(b1[:, 0], b2[:, 0])
    ch = np.maximum(b1[:, 3], b2[:, 3]) - np.minimum(b1[:, 1], b2[:, 1])
    return b1, b2, inter, union, cw, ch


def giou(b1, b2):
    """utils/bbox_tools.py:193-229."""
    _, _, inter, union, cw, ch = _parts(b1, b2)
    iou = inter / np.maximum(union, F(1e-6))
    c_area = cw * ch
    return (iou - np.abs(c_area - union) / np.abs(np.maximum(c_area, F(1e-6)))).astype(F)


def _center_term(b1, b2, cw, ch, eps):
    diag = cw * cw + ch * ch
    dx = (b1[:, 2] + b1[:, 0]) / F(2) - (b2[:, 2] + b2[:, 0]) / F(2)
    dy = (b1[:, 3] + b1[:, 1]) / F(2) - (b2[:, 3] + b2[:, 1]) / F(2)
    return (dx * dx + dy * dy) / np.maximum(diag, F(eps))


def diou(b1, b2):
    """utils/bbox_tools.py:232-283 (clamped to [-1, 1])."""
    b1, b2, inter, union, cw, ch = _parts(b1, b2)
    iou = inter / np.maximum(union, F(1e-6))
    return np.clip(iou - _center_term(b1, b2, cw, ch, 1e-6), F(-1), F(1)).astype(F)


def ciou(b1, b2):
    """utils/bbox_tools.py:286-339."""
    eps = F(1e-9)
    b1, b2, inter, union, cw, ch = _parts(b1, b2)
    iou = inter / np.maximum(union, eps)
    w1, h1 = b1[:, 2] - b1[:, 0], b1[:, 3] - b1[:, 1]
    w2, h2 = b2[:, 2] - b2[:, 0], b2[:, 3] - b2[:, 1]
    da = np.arctan(w1 / np.maximum(h1, eps)) - np.arctan(w2 / np.maximum(h2, eps))
    v = F(4 / (np.pi ** 2)) * (da * da)
    alpha = v / np.maximum(F(1) - iou + v, eps)
    return (iou - (_center_term(b1, b2, cw, ch, 1e-9) + v * alpha)).astype(F)


IOU_FLAVOURS = {"giou": giou, "diou": diou, "ciou": ciou}


def _soft_nms(boxes, scores, kind, iou_threshold, decay, max_picks):
    boxes = np.asarray(boxes, dtype=F).reshape(-1, 4)
    score = np.asarray(scores, dtype=F).reshape(-1).copy()
    processed = np.zeros_like(score)
    fn = IOU_FLAVOURS[kind]
    thr = F(iou_threshold)
    picks = 0
    while score.size and score.max() > 0 and picks < max_picks:  # "while score.sum() > 0" for non-negative scores
        i = int(np.argmax(score))                                # first maximum, as torch.argmax on CPU
        processed[i] = score[i]
        v = fn(boxes[i:i + 1], boxes)
        sel = v > thr
        score[sel] = score[sel] * decay(v[sel])
        picks += 1
    return processed


def linear_soft_nms(boxes, scores, kind, iou_threshold=0.3, thresh=0.001):
    """utils/nms.py:68-103 -> bool (M,)."""
    m = np.asarray(scores).size
    return _soft_nms(boxes, scores, kind, iou_threshold, lambda v: F(1) - v, 64 * m + 1024) > F(thresh)


def exponential_soft_nms(boxes, scores, kind, iou_threshold, sigma=0.5, thresh=0.001):
    """utils/nms.py:106-140 -> bool (M,).  A pick decays itself by exp(-1/sigma) only: the loop re-picks every box
    until its score underflows float32 (~104*sigma picks per box)."""
    m = np.asarray(scores).size
    cap = int(max(64, 104 * sigma + 2)) * m + 1024
    return _soft_nms(boxes, scores, kind, iou_threshold, lambda v: np.exp(-(v * v) / F(sigma)).astype(F), cap) > F(thresh)


def undo_letterbox(rows, scale, pad_top, pad_left, org_h, org_w):
    """val_yolov5.py:166-172 on (K, >=4) float32 rows; returns a copy."""
    out = np.array(rows, dtype=F, copy=True)
    out[:, [0, 2]] = np.clip((out[:, [0, 2]] - F(pad_left)) / F(scale), F(1), F(org_w - 1))
    out[:, [1, 3]] = np.clip((out[:, [1, 3]] - F(pad_top)) / F(scale), F(1), F(org_h - 1))
    return out
