"""Seeded synthetic head tensors in the exact layouts the reference models emit.

Used by bench.py, the tests and the golden-fixture generator (there is no network for
datasets or checkpoints, so every measurement runs on synthetic heads of the named
shapes -- SURVEY.md section 8d).  Three value distributions:

  dense   every logit ~ N(0, 1): nearly all candidates pass conf=0.001 (M ~ N), random
          boxes, almost no suppression -- stresses decode + compaction + selection.
  sparse  objectness ~ N(-9, 1.5), classes ~ N(-2, 1): M ~ 0.5 % of N, the realistic
          survivor rate; the postprocess_bbox window (1 < M < 3000) is active.
  crowd   (anchor-grid families) G objects, each written into up to 36 candidates through
          the inverse decode with small jitter; background objectness -12.  Real
          suppression depth for the NMS stage.
  deepcrowd  the same with 320 objects of 20-34 px (they fit nearly every anchor, ~30 near-identical
          duplicates each, jitter 0.4 px, one confidence level per object): M ~ 7 000 and the 300th
          keep sits thousands of sorted ranks deep -- the NMS walk crosses several selection tranches.

Layouts (C = num_class):
  yolov5  list of (b, 3*(5+C), H, W)            strides 8,16,32   utils/layer_tools.py:454-470
  yolov7  list of (b, 3, H, W, 5+C)             strides 8,16,32   models/normal/yolov7.py:379-405
  yolox   list of (b, 1, 5+C, H, W)             strides 8,16,32   models/normal/yolox_s.py:129-162
  yolov8  list of (b, 64+C, H, W)               strides 4,8,16,32 models/normal/yolov8.py:70-81,124
  retinanet (reg (b,N,4|5), cls (b,N,C))        levels 3..7, 9 anchors/cell
  fcos    (cls list (b,C,H,W), reg list (b,4,H,W), ctr list (b,1,H,W)) strides 8..128
"""
import math

import torch

V5_ANCHORS_PX = [[[10, 13], [16, 30], [33, 23]], [[30, 61], [62, 45], [59, 119]], [[116, 90], [156, 198], [373, 326]]]

FAMILY_STRIDES = {
    "yolov5": (8, 16, 32),
    "yolov7": (8, 16, 32),
    "yolox": (8, 16, 32),
    "yolov8": (4, 8, 16, 32),
    "fcos": (8, 16, 32, 64, 128),
}


def level_shapes(family, img_h, img_w):
    if family in ("retinanet", "retinanet_exp"):
        return [((img_h - 1) // 2 ** l + 1, (img_w - 1) // 2 ** l + 1) for l in (3, 4, 5, 6, 7)]
    return [(img_h // s, img_w // s) for s in FAMILY_STRIDES[family]]


def num_candidates(family, img_h, img_w):
    per_cell = {"yolov5": 3, "yolov7": 3, "yolox": 1, "yolov8": 1, "fcos": 1, "retinanet": 9, "retinanet_exp": 9}[family]
    return per_cell * sum(h * w for h, w in level_shapes(family, img_h, img_w))


def _gen(seed, device):
    g = torch.Generator(device=device)
    g.manual_seed(int(seed))
    return g


def _randn(shape, g, device, mean=0.0, std=1.0):
    return torch.randn(shape, generator=g, device=device, dtype=torch.float32) * std + mean


def make_heads(family, batch, img_h=640, img_w=640, num_class=80, dist="dense", seed=1234, device="cpu"):
    """Returns the head tensors of ``family`` (see module docstring) for ``batch`` images."""
    g = _gen(seed, device)
    C = num_class
    shapes = level_shapes(family, img_h, img_w)
    if dist in ("crowd", "deepcrowd"):
        if family not in ("yolov5", "yolov7"):
            raise ValueError("crowd distributions are implemented for the anchor-grid families")
        if dist == "crowd":
            heads = _crowd_v5(batch, shapes, C, g, device)
        else:
            heads = _crowd_v5(batch, shapes, C, g, device, objects_per_image=330, jitter_px=0.4, jitter_log=0.01,
                              size_lo=20.0, size_span=1.7, per_object_score=True)
        if family == "yolov7":
            heads = [h.view(batch, 3, 5 + C, h.shape[2], h.shape[3]).permute(0, 1, 3, 4, 2).contiguous() for h in heads]
        return heads
    if dist not in ("dense", "sparse"):
        raise ValueError(f"unknown distribution {dist!r}")
    sparse = dist == "sparse"
    if family in ("yolov5", "yolov7", "yolox"):
        na = 1 if family == "yolox" else 3
        heads = []
        for h, w in shapes:
            t = _randn((batch, na, 5 + C, h, w), g, device)
            if family == "yolox":
                t[:, :, 2:4] *= 0.5  # keep exp() boxes well below the 4096 class offset
            if sparse:
                t[:, :, 4] = t[:, :, 4] * 1.5 - 9.0
                t[:, :, 5:] = t[:, :, 5:] - 2.0
            if family == "yolov5":
                t = t.view(batch, na * (5 + C), h, w)
            elif family == "yolov7":
                t = t.permute(0, 1, 3, 4, 2).contiguous()
            heads.append(t.contiguous())
        return heads
    if family == "yolov8":
        heads = []
        for h, w in shapes:
            t = _randn((batch, 64 + C, h, w), g, device)
            if sparse:
                t[:, 64:] = t[:, 64:] * 1.5 - 8.0
            heads.append(t.contiguous())
        return heads
    if family in ("retinanet", "retinanet_exp"):
        n = 9 * sum(h * w for h, w in shapes)
        reg = _randn((batch, n, 5 if family == "retinanet_exp" else 4), g, device)
        cls = _randn((batch, n, C), g, device)
        if sparse:
            cls = cls * 1.5 - 8.0
        return reg.contiguous(), cls.contiguous()
    if family == "fcos":
        cls_l, reg_l, ctr_l = [], [], []
        for h, w in shapes:
            c = _randn((batch, C, h, w), g, device)
            if sparse:
                c = c * 1.5 - 5.0
            cls_l.append(c.contiguous())
            reg_l.append(_randn((batch, 4, h, w), g, device).abs().contiguous() * 2.0)
            ctr_l.append(_randn((batch, 1, h, w), g, device).contiguous())
        return cls_l, reg_l, ctr_l
    raise ValueError(f"unknown family {family!r}")


def _logit(p):
    p = p.clamp(1e-4, 1 - 1e-4)
    return torch.log(p) - torch.log1p(-p)


def _logit64(p):
    """logit evaluated in float64 and rounded once to float32: vectorised float32 log/exp differ by an ulp between CPU
    generations (AVX2 vs AVX-512 code paths), float64 results rounded to float32 practically never do -- seed-only
    fixtures (tests/golden/c1_*) need heads that are bit-identical wherever they are regenerated."""
    p = p.double().clamp(1e-4, 1 - 1e-4)
    return (torch.log(p) - torch.log1p(-p)).float()


def _crowd_v5(batch, shapes, C, g, device, objects_per_image=700, jitter_px=3.0, jitter_log=0.08, size_lo=16.0,
              size_span=12.0, per_object_score=False):
    """Objects written through the inverse of the v5 decode (trainer/eval_yolov5.py:203-205)."""
    strides = FAMILY_STRIDES["yolov5"]
    img_w = shapes[0][1] * strides[0]
    img_h = shapes[0][0] * strides[0]
    heads = []
    for h, w in shapes:
        t = _randn((batch, 3, 5 + C, h, w), g, device)
        t[:, :, 4] = -12.0
        heads.append(t)
    anchors = torch.tensor(V5_ANCHORS_PX, dtype=torch.float32, device=device)  # (3, 3, 2)
    for b in range(batch):
        G = objects_per_image
        cx = torch.rand(G, generator=g, device=device) * (img_w - 64) + 32
        cy = torch.rand(G, generator=g, device=device) * (img_h - 64) + 32
        f64 = per_object_score   # deepcrowd: transcendental steps in float64 (see _logit64); 'crowd' keeps its float32 bits
        if f64:
            bw = (torch.exp(torch.rand(G, generator=g, device=device).double() * math.log(size_span)) * size_lo).float()
            bh = (bw.double() * torch.exp(_randn((G,), g, device, std=0.35).double())).float()
        else:
            bw = torch.exp(torch.rand(G, generator=g, device=device) * math.log(size_span)) * size_lo
            bh = bw * torch.exp(_randn((G,), g, device, std=0.35))
        cls_id = torch.randint(0, C, (G,), generator=g, device=device)
        # per_object_score: every object has its own confidence level shared by all its duplicates, so the duplicates of a
        # confident object precede the first candidate of a less confident one in the NMS visiting order
        base = (torch.rand(G, generator=g, device=device) * 9.0 - 4.0) if per_object_score else None
        for lvl, ((h, w), s) in enumerate(zip(shapes, strides)):
            for a in range(3):
                aw, ah = anchors[lvl, a, 0], anchors[lvl, a, 1]
                fit = (bw < 3.6 * aw) & (bh < 3.6 * ah) & (bw > 0.05 * aw) & (bh > 0.05 * ah)
                for dx in (0, 1):
                    for dy in (0, 1):
                        gx = torch.floor(cx / s - 0.5).long() + dx
                        gy = torch.floor(cy / s - 0.5).long() + dy
                        ok = fit & (gx >= 0) & (gx < w) & (gy >= 0) & (gy < h)
                        idx = ok.nonzero().flatten()
                        if idx.numel() == 0:
                            continue
                        # two objects can land in the same cell: an indexed write with duplicate indices is resolved in an
                        # unspecified order (it depends on the thread count), so keep only the LAST object per cell --
                        # what a sequential write does -- and the heads are identical wherever they are generated
                        lin = gy[idx] * w + gx[idx]
                        order = torch.sort(lin, stable=True).indices
                        lin_s = lin[order]
                        last = torch.ones_like(lin_s, dtype=torch.bool)
                        last[:-1] = lin_s[1:] != lin_s[:-1]
                        idx = idx[torch.sort(order[last]).values]
                        n = idx.numel()
                        jx = cx[idx] + _randn((n,), g, device, std=jitter_px)
                        jy = cy[idx] + _randn((n,), g, device, std=jitter_px)
                        if f64:
                            jw = (bw[idx].double() * torch.exp(_randn((n,), g, device, std=jitter_log).double())).float()
                            jh = (bh[idx].double() * torch.exp(_randn((n,), g, device, std=jitter_log).double())).float()
                        else:
                            jw = bw[idx] * torch.exp(_randn((n,), g, device, std=jitter_log))
                            jh = bh[idx] * torch.exp(_randn((n,), g, device, std=jitter_log))
                        px = ((jx / s - gx[idx].float()) + 0.5) / 2.0
                        py = ((jy / s - gy[idx].float()) + 0.5) / 2.0
                        pw = torch.sqrt(jw / aw) / 2.0
                        ph = torch.sqrt(jh / ah) / 2.0
                        tgt = heads[lvl][b, a]
                        yy, xx = gy[idx], gx[idx]
                        lg = _logit64 if f64 else _logit
                        tgt[0, yy, xx] = lg(px)
                        tgt[1, yy, xx] = lg(py)
                        tgt[2, yy, xx] = lg(pw)
                        tgt[3, yy, xx] = lg(ph)
                        if per_object_score:
                            tgt[4, yy, xx] = base[idx] + _randn((n,), g, device, std=0.1)
                        else:
                            tgt[4, yy, xx] = _randn((n,), g, device, mean=2.0, std=1.5)
                        tgt[5:, yy, xx] = _randn((C, n), g, device, mean=-4.0, std=1.0)
                        tgt[5 + cls_id[idx], yy, xx] = _randn((n,), g, device, mean=2.5, std=0.05 if per_object_score else 1.0)
    return [t.view(batch, 3 * (5 + C), t.shape[3], t.shape[4]).contiguous() for t in heads]


def map_profile_hyp(**over):
    """The reference's mAP-profile post-processing settings (config/train_yolov5.yaml:85-111 'compute_metric_*',
    config/validation.yaml:3-20) as the flat ``hyp`` dict engine.make_params reads; ``over`` replaces entries."""
    hyp = dict(
        num_class=80, conf_threshold=0.001, cls_threshold=0.001, iou_threshold=0.65, max_predictions_per_img=300,
        min_prediction_box_wh=2, mutil_label=False, agnostic=True, postprocess_bbox=True, pre_nms_topk=1000,
        pre_nms_thresh=0.05, thresh_with_ctr=True,
    )
    hyp.update(over)
    return hyp
