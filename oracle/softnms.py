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


# ---- backward of the row-wise flavours (torch autograd of utils/bbox_tools.py:193-339, restated analytically) -----------
def _split_max(p, q, g):
    """d max(p, q): torch.maximum sends the whole gradient to the larger input and half to each at a tie."""
    wp = np.where(p > q, 1.0, np.where(p == q, 0.5, 0.0))
    return g * wp, g * (1.0 - wp)


def _split_min(p, q, g):
    wp = np.where(p < q, 1.0, np.where(p == q, 0.5, 0.0))
    return g * wp, g * (1.0 - wp)


def iou_backward(kind, b1, b2, grad_out):
    """d sum(grad_out * flavour(b1, b2)) / d b1, d b2 in float64 on float32 inputs; b1 (1,4) is broadcast and its
    gradient summed.  clamp passes the gradient on its bounds, abs'(0) = 0, CIoU's alpha is a constant (:334-335)."""
    a = np.asarray(b1, dtype=np.float64).reshape(-1, 4)
    b = np.asarray(b2, dtype=np.float64).reshape(-1, 4)
    G = np.asarray(grad_out, dtype=np.float64).reshape(-1)
    bc = a.shape[0] == 1 and b.shape[0] != 1
    if bc:
        a = np.repeat(a, b.shape[0], axis=0)
    eps_u = 1e-9 if kind == "ciou" else 1e-6
    w1, h1, w2, h2 = a[:, 2] - a[:, 0], a[:, 3] - a[:, 1], b[:, 2] - b[:, 0], b[:, 3] - b[:, 1]
    tw = np.minimum(a[:, 2], b[:, 2]) - np.maximum(a[:, 0], b[:, 0])
    th = np.minimum(a[:, 3], b[:, 3]) - np.maximum(a[:, 1], b[:, 1])
    iw, ih = np.maximum(tw, 0.0), np.maximum(th, 0.0)
    inter = iw * ih
    u_raw = w1 * h1 + w2 * h2 - inter
    uc = np.maximum(u_raw, eps_u)
    iou = inter / uc
    cw = np.maximum(a[:, 2], b[:, 2]) - np.minimum(a[:, 0], b[:, 0])
    ch = np.maximum(a[:, 3], b[:, 3]) - np.minimum(a[:, 1], b[:, 1])
    z = np.zeros_like(G)
    d_uraw, d_cw, d_ch, d_w1, d_h1, d_w2, d_h2, d_dx, d_dy = (z.copy() for _ in range(9))
    if kind == "giou":
        C = cw * ch
        cc = np.maximum(C, 1e-6)
        diff = C - u_raw
        d_iou = G
        d_diff = (-G / np.abs(cc)) * np.sign(diff)
        d_C = d_diff + np.where(C >= 1e-6, (G * np.abs(diff) / cc ** 2) * np.sign(cc), 0.0)
        d_uraw = d_uraw - d_diff
        d_cw, d_ch = d_C * ch, d_C * cw
    else:
        eps_c = 1e-9 if kind == "ciou" else 1e-6
        c2 = cw * cw + ch * ch
        c2c = np.maximum(c2, eps_c)
        dx = (a[:, 2] + a[:, 0]) / 2 - (b[:, 2] + b[:, 0]) / 2
        dy = (a[:, 3] + a[:, 1]) / 2 - (b[:, 3] + b[:, 1]) / 2
        rho2 = dx * dx + dy * dy
        g_pre = G.copy()
        if kind == "diou":
            pre = iou - rho2 / c2c
            g_pre = np.where((pre >= -1) & (pre <= 1), G, 0.0)
        d_iou = g_pre
        d_c2 = np.where(c2 >= eps_c, g_pre * rho2 / c2c ** 2, 0.0)
        d_cw, d_ch = 2 * cw * d_c2, 2 * ch * d_c2
        d_dx, d_dy = 2 * dx * (-g_pre / c2c), 2 * dy * (-g_pre / c2c)
        if kind == "ciou":
            h1c, h2c = np.maximum(h1, 1e-9), np.maximum(h2, 1e-9)
            r1, r2 = w1 / h1c, w2 / h2c
            delta = np.arctan(r1) - np.arctan(r2)
            k = 4 / np.pi ** 2
            v = k * delta ** 2
            alpha = v / np.maximum(1 - iou + v, 1e-9)
            d_delta = (-g_pre * alpha) * k * 2 * delta
            d_r1, d_r2 = d_delta / (1 + r1 * r1), -d_delta / (1 + r2 * r2)
            d_w1 = d_w1 + d_r1 / h1c
            d_h1 = d_h1 + np.where(h1 >= 1e-9, -d_r1 * w1 / h1c ** 2, 0.0)
            d_w2 = d_w2 + d_r2 / h2c
            d_h2 = d_h2 + np.where(h2 >= 1e-9, -d_r2 * w2 / h2c ** 2, 0.0)
    d_inter = d_iou / uc
    d_uraw = d_uraw + np.where(u_raw >= eps_u, -d_iou * inter / uc ** 2, 0.0)
    d_w1, d_h1 = d_w1 + d_uraw * h1, d_h1 + d_uraw * w1
    d_w2, d_h2 = d_w2 + d_uraw * h2, d_h2 + d_uraw * w2
    d_inter = d_inter - d_uraw
    d_tw = np.where(tw >= 0, d_inter * ih, 0.0)
    d_th = np.where(th >= 0, d_inter * iw, 0.0)
    ga = np.stack([-d_w1 + d_dx / 2, -d_h1 + d_dy / 2, d_w1 + d_dx / 2, d_h1 + d_dy / 2], axis=1)
    gb = np.stack([-d_w2 - d_dx / 2, -d_h2 - d_dy / 2, d_w2 - d_dx / 2, d_h2 - d_dy / 2], axis=1)
    for col_lo, col_hi, d_t, d_c in ((0, 2, d_tw, d_cw), (1, 3, d_th, d_ch)):
        p, q = _split_max(a[:, col_lo], b[:, col_lo], -d_t)     # intersection low edge = max
        ga[:, col_lo] += p; gb[:, col_lo] += q
        p, q = _split_min(a[:, col_hi], b[:, col_hi], d_t)      # intersection high edge = min
        ga[:, col_hi] += p; gb[:, col_hi] += q
        p, q = _split_min(a[:, col_lo], b[:, col_lo], -d_c)     # enclosing low edge = min
        ga[:, col_lo] += p; gb[:, col_lo] += q
        p, q = _split_max(a[:, col_hi], b[:, col_hi], d_c)      # enclosing high edge = max
        ga[:, col_hi] += p; gb[:, col_hi] += q
    if bc:
        ga = ga.sum(axis=0, keepdims=True)
    return ga, gb


def pairwise_iou_backward(b1, b2, grad_out):
    """Backward of utils.gpu_iou (utils/bbox_tools.py:164-190): d sum(grad_out * iou(b1, b2)) / d b1 (N,4), d b2 (M,4),
    float64 on float32 inputs.  prod -> w*h; torch.min/max split ties evenly; clamp(min=0.0) and clamp(1e-9) pass the
    gradient on their bound.  Pinned by tests/golden/utils_extra.npz (torch autograd through the reference's function)."""
    a = np.asarray(b1, dtype=np.float64).reshape(-1, 1, 4)
    b = np.asarray(b2, dtype=np.float64).reshape(1, -1, 4)
    G = np.asarray(grad_out, dtype=np.float64).reshape(a.shape[0], b.shape[1])
    w1, h1 = a[..., 2] - a[..., 0], a[..., 3] - a[..., 1]
    w2, h2 = b[..., 2] - b[..., 0], b[..., 3] - b[..., 1]
    tw = np.minimum(a[..., 2], b[..., 2]) - np.maximum(a[..., 0], b[..., 0])
    th = np.minimum(a[..., 3], b[..., 3]) - np.maximum(a[..., 1], b[..., 1])
    iw, ih = np.maximum(tw, 0.0), np.maximum(th, 0.0)
    inter = iw * ih
    u_raw = w1 * h1 + w2 * h2 - inter
    uc = np.maximum(u_raw, 1e-9)
    d_uraw = np.where(u_raw >= 1e-9, -G * inter / uc ** 2, 0.0)
    d_inter = G / uc - d_uraw
    d_tw = np.where(tw >= 0, d_inter * ih, 0.0)
    d_th = np.where(th >= 0, d_inter * iw, 0.0)
    n, m = G.shape
    ga = np.zeros((n, m, 4))
    gb = np.zeros((n, m, 4))
    ga[..., 0], ga[..., 1], ga[..., 2], ga[..., 3] = -d_uraw * h1, -d_uraw * w1, d_uraw * h1, d_uraw * w1
    gb[..., 0], gb[..., 1], gb[..., 2], gb[..., 3] = -d_uraw * h2, -d_uraw * w2, d_uraw * h2, d_uraw * w2
    for lo, hi, d_t in ((0, 2, d_tw), (1, 3, d_th)):
        p, q = _split_max(np.broadcast_to(a[..., lo], (n, m)), np.broadcast_to(b[..., lo], (n, m)), -d_t)
        ga[..., lo] += p; gb[..., lo] += q
        p, q = _split_min(np.broadcast_to(a[..., hi], (n, m)), np.broadcast_to(b[..., hi], (n, m)), d_t)
        ga[..., hi] += p; gb[..., hi] += q
    return ga.sum(axis=1), gb.sum(axis=0)
