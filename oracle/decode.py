"""numpy restatement of the ``do_inference`` decode step of every evaluator family.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  All arithmetic is float32 with one
rounding per operation, in the reference's operation order.  ``np.exp`` on float32 may
differ from torch's CPU/CUDA ``exp`` by an ulp, therefore decoded tensors are compared
with the north-star tolerance ``|a-b| <= 1e-5 * max(|b|, 1)`` and never bit-for-bit.

Each function takes the raw head tensors as numpy arrays in the layout the reference
model emits and returns the ``(b, N, C')`` array ``do_inference`` returns.
"""
import numpy as np

F32 = np.float32

# val_yolov5.py:405 -- pixel anchors per stage (w, h)
V5_ANCHORS = np.array(
    [[[10, 13], [16, 30], [33, 23]], [[30, 61], [62, 45], [59, 119]], [[116, 90], [156, 198], [373, 326]]],
    dtype=np.int64,
)


def sigmoid_f32(x):
    """1 / (1 + exp(-x)) evaluated in float32 (torch.sigmoid semantics)."""
    x = np.asarray(x, dtype=F32)
    with np.errstate(over="ignore"):
        return (F32(1.0) / (F32(1.0) + np.exp(-x))).astype(F32)


def _xy_grid(h, w):
    """(h, w, 2) float32 grid of [x, y] cell indices (eval_yolov5.py:229-234, eval_yolox.py:170-174)."""
    ys, xs = np.meshgrid(np.arange(h), np.arange(w), indexing="ij")
    return np.stack((xs, ys), axis=2).astype(F32)


def decode_yolov5(heads, anchors=V5_ANCHORS, strides=(8, 16, 32), num_class=80):
    """trainer/eval_yolov5.py:181-209.  heads[i]: (b, A*(5+C), H, W) -> (b, N, 5+C).

    p = sigmoid(raw) on all channels (:195); xy = (p*2 - 0.5 + grid) * s (:203);
    wh = (p*2)**2 * (anchor/s) * s (:205); candidate order = stage, anchor, y, x (:207-209).
    """
    out = []
    for i, raw in enumerate(heads):
        raw = np.asarray(raw, dtype=F32)
        b, _, h, w = raw.shape
        a_num = anchors.shape[1]
        s = F32(strides[i])
        p = sigmoid_f32(raw.reshape(b, a_num, 5 + num_class, h, w).transpose(0, 1, 3, 4, 2))
        stage_anchor = (anchors[i].astype(np.float64) / float(strides[i])).astype(F32)[None, :, None, None, :]
        grid = _xy_grid(h, w)[None, None]
        xy = ((p[..., 0:2] * F32(2) - F32(0.5)) + grid) * s
        t = p[..., 2:4] * F32(2)
        wh = ((t * t) * stage_anchor) * s
        p = p.copy()
        p[..., 0:2] = xy
        p[..., 2:4] = wh
        out.append(p.reshape(b, -1, 5 + num_class))
    return np.concatenate(out, axis=1)


def decode_yolov7(heads, anchors=V5_ANCHORS, strides=(8, 16, 32), num_class=80):
    """trainer/eval_yolov7.py:123-151.  heads[i]: (b, A, H, W, 5+C) channels-last; same formula as v5."""
    out = []
    for i, raw in enumerate(heads):
        raw = np.asarray(raw, dtype=F32)
        b, a_num, h, w, _ = raw.shape
        s = F32(strides[i])
        p = sigmoid_f32(raw)
        stage_anchor = (anchors[i].astype(np.float64) / float(strides[i])).astype(F32)[None, :, None, None, :]
        grid = _xy_grid(h, w)[None, None]
        xy = ((p[..., 0:2] * F32(2) - F32(0.5)) + grid) * s
        t = p[..., 2:4] * F32(2)
        wh = ((t * t) * stage_anchor) * s
        p[..., 0:2] = xy
        p[..., 2:4] = wh
        out.append(p.reshape(b, -1, 5 + num_class))
    return np.concatenate(out, axis=1)


def decode_yolox(heads, input_h, num_class=80):
    """trainer/eval_yolox.py:123-150.  heads[i]: (b, na, 5+C, H, W).

    sigmoid on channels 4: only (:140); xy = (raw + grid) * (in_h/H) (:144); wh = exp(raw) * (in_h/H) (:146).
    """
    out = []
    for raw in heads:
        raw = np.asarray(raw, dtype=F32)
        b, na, _, h, w = raw.shape
        s = F32(input_h / h)
        cur = raw.transpose(0, 1, 3, 4, 2).copy()
        cur[..., 4:] = sigmoid_f32(cur[..., 4:])
        grid = _xy_grid(h, w)[None, None]
        xy = (cur[..., 0:2] + grid) * s
        with np.errstate(over="ignore"):
            wh = np.exp(cur[..., 2:4]) * s
        cur[..., 0:2] = xy
        cur[..., 2:4] = wh
        out.append(cur.reshape(b, -1, 5 + num_class))
    return np.concatenate(out, axis=1)


def decode_yolov8(heads, input_h, reg=16, num_class=80):
    """trainer/eval_yolov8.py:75-102 with utils/bbox_tools.py:392-407.

    heads[i]: (b, 4*reg + C, H, W), regression channels first.  Per side a softmax over
    ``reg`` bins dotted with [1..reg] (:80,84 -- bins start at 1, a reference quirk) gives
    [t, b, l, r]; xyxy = (gx - l, gy - t, gx + r, gy + b) * stride; grid from make_grid
    (:122-141), which for flat index n of an (h, w) level yields (n % h + .5, n // h + .5).
    Output (b, N, 4 + C) = [x1, y1, x2, y2, sigmoid(cls)...].
    """
    boxes, clss = [], []
    bins = np.arange(1, reg + 1, dtype=F32)
    for raw in heads:
        raw = np.asarray(raw, dtype=F32)
        b, sf, h, w = raw.shape
        s = F32(input_h / h)
        flat = raw.reshape(b, sf, h * w).transpose(0, 2, 1)  # (b, hw, sf)
        logits = flat[..., : 4 * reg].reshape(b, h * w, 4, reg)
        mx = logits.max(axis=-1, keepdims=True)
        e = np.exp(logits - mx)
        prob = e / e.sum(axis=-1, keepdims=True, dtype=F32)
        tblr = (prob * bins).sum(axis=-1, dtype=F32)  # (b, hw, 4) [t, b, l, r]
        n = np.arange(h * w)
        gx = ((n % h).astype(F32) + F32(0.5))[None, :]
        gy = ((n // h).astype(F32) + F32(0.5))[None, :]
        x1 = (gx - tblr[..., 2]) * s
        y1 = (gy - tblr[..., 0]) * s
        x2 = (gx + tblr[..., 3]) * s
        y2 = (gy + tblr[..., 1]) * s
        boxes.append(np.stack((x1, y1, x2, y2), axis=-1))
        clss.append(sigmoid_f32(flat[..., 4 * reg:]))
    return np.concatenate((np.concatenate(boxes, axis=1), np.concatenate(clss, axis=1)), axis=-1).astype(F32)


def retinanet_base_anchors(size):
    """utils/anchor.py:176-191 -- the 9 base anchors [x1,y1,x2,y2] of one level, float32, ratio-major."""
    scales = np.array([1, 2 ** (1 / 3), 2 ** (2 / 3)], dtype=F32)
    ratios = np.array([0.5, 1, 2], dtype=F32)
    base = np.zeros((9, 4), dtype=F32)
    side = F32(size) * np.tile(scales, 3)
    base[:, 2] = side
    base[:, 3] = side
    areas = base[:, 2] * base[:, 3]
    rr = np.repeat(ratios, 3)
    base[:, 2] = np.sqrt(areas / rr)
    base[:, 3] = base[:, 2] * rr
    half_w = base[:, 2] / F32(2)
    base[:, 0] -= half_w
    base[:, 2] -= half_w
    half_h = base[:, 3] / F32(2)
    base[:, 1] -= half_h
    base[:, 3] -= half_h
    return base


def retinanet_levels(img_h, img_w):
    """Feature-map shapes/strides/sizes of pyramid levels 3..7 (utils/anchor.py:139-150,213-222)."""
    out = []
    for lvl in (3, 4, 5, 6, 7):
        fh = (img_h - 1) // (2 ** lvl) + 1
        fw = (img_w - 1) // (2 ** lvl) + 1
        out.append((fh, fw, 2 ** lvl, 2 ** (lvl + 2)))
    return out


def retinanet_anchors(img_h, img_w):
    """utils/anchor.py:159-211 -- (N, 4) float32 anchors, order = level, y, x, anchor."""
    chunks = []
    for fh, fw, stride, size in retinanet_levels(img_h, img_w):
        base = retinanet_base_anchors(size)
        sx = (np.arange(fw).astype(F32) + F32(0.5)) * F32(stride)
        sy = (np.arange(fh).astype(F32) + F32(0.5)) * F32(stride)
        gx, gy = np.meshgrid(sx, sy, indexing="xy")  # (fh, fw): gx[y, x] = sx[x]
        shifts = np.stack((gx.ravel(), gy.ravel(), gx.ravel(), gy.ravel()), axis=1)  # (K, 4)
        chunks.append((shifts[:, None, :] + base[None, :, :]).reshape(-1, 4))
    return np.concatenate(chunks, axis=0).astype(F32)


def decode_retinanet(reg, cls, img_h, img_w, scale_factor=(0.1, 0.1, 0.2, 0.2)):
    """trainer/eval_retinanet.py:22-75,185-200 (and eval_retinanet_experiment.py:75 for a 5th conf column).

    reg: (b, N, 4|5), cls: (b, N, C) -> (b, N, C+4|5) = [sigmoid(cls), x1, y1, x2, y2(, sigmoid(conf))].
    deltas *= scale (:39); centre = a_ctr + d * a_wh (:46-47); wh = exp(d) * a_wh (:48-49);
    corners = centre -/+ wh*0.5 (:51-54); round half-to-even then clamp to the image (:195-199).
    """
    reg = np.asarray(reg, dtype=F32)
    cls = np.asarray(cls, dtype=F32)
    anc = retinanet_anchors(img_h, img_w)
    aw = anc[:, 2] - anc[:, 0]
    ah = anc[:, 3] - anc[:, 1]
    acx = anc[:, 0] + aw * F32(0.5)
    acy = anc[:, 1] + ah * F32(0.5)
    d = reg[..., :4] * np.asarray(scale_factor, dtype=F32)
    pcx = acx + d[..., 0] * aw
    pcy = acy + d[..., 1] * ah
    with np.errstate(over="ignore"):
        pw = np.exp(d[..., 2]) * aw
        ph = np.exp(d[..., 3]) * ah
    x1 = pcx - pw * F32(0.5)
    y1 = pcy - ph * F32(0.5)
    x2 = pcx + pw * F32(0.5)
    y2 = pcy + ph * F32(0.5)
    box = np.rint(np.stack((x1, y1, x2, y2), axis=-1)).astype(F32)
    box[..., 0] = np.clip(box[..., 0], 0, img_w)
    box[..., 1] = np.clip(box[..., 1], 0, img_h)
    box[..., 2] = np.clip(box[..., 2], 0, img_w)
    box[..., 3] = np.clip(box[..., 3], 0, img_h)
    parts = [sigmoid_f32(cls), box]
    if reg.shape[-1] == 5:
        parts.append(sigmoid_f32(reg[..., 4:5]))
    return np.concatenate(parts, axis=-1).astype(F32)


def decode_fcos(cls_fms, reg_fms, ctr_fms, input_h):
    """trainer/eval_fcos.py:125-161,181-191.

    cls_fms[i]: (b, C, H, W); reg_fms[i]: (b, 4, H, W) as [l, t, r, b]; ctr_fms[i]: (b, 1, H, W).
    stride = in_h / H; box = [gx - l*s, gy - t*s, gx + r*s, gy + b*s] with
    gx = x*s + s//2, gy = y*s + s//2 (square maps only -- the reference grid does not
    broadcast otherwise).  Output (b, N, 5+C) = [x1, y1, x2, y2, sigmoid(ctr), sigmoid(cls)...].
    """
    out = []
    for c, r, t in zip(cls_fms, reg_fms, ctr_fms):
        c, r, t = (np.asarray(v, dtype=F32) for v in (c, r, t))
        b, _, h, w = c.shape
        s = F32(input_h / h)
        half = F32((input_h / h) // 2)
        regs = r.transpose(0, 2, 3, 1) * s  # (b, h, w, 4)
        gx = (np.arange(w).astype(F32) * s + half)[None, None, :]
        gy = (np.arange(h).astype(F32) * s + half)[None, :, None]
        x1 = gx - regs[..., 0]
        y1 = gy - regs[..., 1]
        x2 = gx + regs[..., 2]
        y2 = gy + regs[..., 3]
        box = np.stack((x1, y1, x2, y2), axis=-1).reshape(b, h * w, 4)
        ctr = sigmoid_f32(t.transpose(0, 2, 3, 1)).reshape(b, h * w, 1)
        cl = sigmoid_f32(c.transpose(0, 2, 3, 1)).reshape(b, h * w, -1)
        out.append(np.concatenate((box, ctr, cl), axis=-1))
    return np.concatenate(out, axis=1).astype(F32)
