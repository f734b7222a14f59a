"""Generate tests/golden/*.npz by running the UNMODIFIED reference (build container only).

TEST INFRASTRUCTURE ONLY.  Usage:  python -m oracle.gen_golden
The reference has no tests or golden vectors of its own, so these
fixtures -- outputs of the reference's own code on seeded synthetic inputs -- are what pins
the oracle (tests/test_oracle.py) and, through it, the CUDA path.

Generated with torch 2.11.0 (CPU), numba 0.65.0, numpy 2.3.5.
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import refharness  # noqa: E402
from yoloseries_b200 import synth  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")


def _pack_outputs(outs):
    """list[ndarray(K,6) | None] -> (rows (b, Kmax, 6), counts (b,), -1 == None)."""
    kmax = max([o.shape[0] for o in outs if o is not None] + [1])
    rows = np.zeros((len(outs), kmax, 6), dtype=np.float32)
    cnt = np.zeros(len(outs), dtype=np.int32)
    for i, o in enumerate(outs):
        if o is None:
            cnt[i] = -1
        else:
            cnt[i] = o.shape[0]
            rows[i, : o.shape[0]] = o
    return rows, cnt


def _clone(x):
    if isinstance(x, torch.Tensor):
        return x.clone()
    if isinstance(x, (list, tuple)):
        return type(x)(_clone(v) for v in x)
    if isinstance(x, dict):
        return type(x)((k, _clone(v)) for k, v in x.items())
    return x


def _flat_np(prefix, heads, store):
    if isinstance(heads, torch.Tensor):
        store[prefix] = heads.numpy()
    else:
        for i, h in enumerate(heads):
            _flat_np(f"{prefix}_{i}", h, store)


def evaluator_case(trainer, family, dist, img, batch, seed, C=80, **hyp_over):
    from collections import OrderedDict

    hyp = refharness.reference_hyp((img, img), num_class=C, **hyp_over)
    heads = synth.make_heads(family, batch, img, img, C, dist, seed, "cpu")
    dummy = torch.zeros(batch, 3, img, img)
    anchors = torch.tensor(synth.V5_ANCHORS_PX)
    # the reference mutates model outputs in place for v7 / retinanet: hand it clones
    if family == "yolov5":
        ev = trainer.YOLOV5Evaluator(lambda x: _clone(heads), anchors, hyp, compute_metric=True)
    elif family == "yolov7":
        ev = trainer.YOLOV7Evaluator(lambda x: OrderedDict((f"p{i}", h.clone()) for i, h in enumerate(heads)),
                                     anchors, hyp, compute_metric=True)
    elif family == "yolox":
        ev = trainer.YOLOXEvaluator(lambda x: OrderedDict((f"p{i}", h.clone()) for i, h in enumerate(heads)),
                                    hyp, compute_metric=True)
    elif family == "yolov8":
        ev = trainer.YOLOV8Evaluator(lambda x: OrderedDict((f"p{i}", h.clone()) for i, h in enumerate(heads)),
                                     hyp, compute_metric=True)
    elif family == "retinanet":
        ev = trainer.RetinaNetEvaluator(lambda x: _clone(heads), hyp, compute_metric=True)
    elif family == "retinanet_exp":
        ev = trainer.RetinaNetEvaluatorExperiment(lambda x: _clone(heads), hyp, compute_metric=True)
    elif family == "fcos":
        ev = trainer.FCOSEvaluator(lambda x: _clone(heads), hyp, compute_metric=True)
    else:
        raise ValueError(family)
    decoded = ev.do_inference(dummy)
    outs = ev.numba_nms(decoded.clone())
    rows, cnt = _pack_outputs(outs)
    store = {"decoded": decoded.numpy().astype(np.float32), "rows": rows, "counts": cnt}
    _flat_np("head", heads, store)
    meta = dict(family=family, dist=dist, img=img, batch=batch, seed=seed, num_class=C)
    meta.update({k: v for k, v in hyp.items() if isinstance(v, (int, float, bool, str))})
    store["meta"] = np.array(repr(meta))
    return store


def tta_case(trainer, family, img_h, img_w, batch, seed, C=4, **hyp_over):
    """use_tta=True: the reference's own test_time_augmentation + numba_nms over three passes (scale 1 / 0.83 / 0.67,
    flip none / h / w).  The stand-in model hands out a different seeded head set on every forward."""
    from collections import OrderedDict

    hyp = refharness.reference_hyp((img_h, img_w), num_class=C, use_tta=True, **hyp_over)
    sets = [synth.make_heads(family, batch, img_h, img_w, C, "dense", seed + k, "cpu") for k in range(3)]
    calls = {"n": 0}

    def model(x):
        assert tuple(x.shape[2:]) == (img_h, img_w), x.shape  # multiples of 32: every pass is padded back to the input size
        heads = _clone(sets[calls["n"] % 3])
        calls["n"] += 1
        if family in ("yolov7", "yolox", "yolov8"):
            return OrderedDict((f"p{i}", h) for i, h in enumerate(heads))
        return heads

    anchors = torch.tensor(synth.V5_ANCHORS_PX)
    cls = {"yolov5": "YOLOV5Evaluator", "yolov7": "YOLOV7Evaluator", "yolox": "YOLOXEvaluator", "yolov8": "YOLOV8Evaluator",
           "retinanet": "RetinaNetEvaluator", "retinanet_exp": "RetinaNetEvaluatorExperiment", "fcos": "FCOSEvaluator"}[family]
    args = (model, anchors, hyp) if family in ("yolov5", "yolov7") else (model, hyp)
    ev = getattr(trainer, cls)(*args, compute_metric=True)
    # distinguishable picture content so that a wrong flip/scale of the INPUT would not go unnoticed by a real model;
    # the stand-in model ignores it
    dummy = torch.rand(batch, 3, img_h, img_w, generator=torch.Generator().manual_seed(seed))
    merged, per_pass = ev.test_time_augmentation(dummy)
    assert calls["n"] == 3
    outs = ev.numba_nms(merged.clone())
    rows, cnt = _pack_outputs(outs)
    # __call__ must agree with the two-step form
    again = ev(dummy)
    for a, b in zip(again, outs):
        assert (a is None and b is None) or np.array_equal(a.numpy(), b)
    store = {"merged": merged.numpy().astype(np.float32), "rows": rows, "counts": cnt}
    assert calls["n"] % 3 == 0
    for k in range(3):
        _flat_np(f"p{k}_head", sets[k], store)
        # the pass's decoded tensor BEFORE the scale/flip undo (the stand-in model hands out set k again)
        store[f"p{k}_decoded"] = ev.do_inference(dummy).numpy().astype(np.float32)
    meta = dict(family=family, dist="dense", img=img_h, img_h=img_h, img_w=img_w, batch=batch, seed=seed, num_class=C)
    meta.update({k: v for k, v in hyp.items() if isinstance(v, (int, float, bool, str))})
    store["meta"] = np.array(repr(meta))
    return store


def utils_case(utils):
    """utils/nms.py + utils/bbox_tools.py on random and hand-made boxes."""
    rng = np.random.default_rng(7)
    store = {}
    for tag, m, span in (("a", 400, 120.0), ("b", 1500, 300.0)):
        xy = rng.uniform(0, span, size=(m, 2)).astype(np.float32)
        wh = rng.uniform(4, 60, size=(m, 2)).astype(np.float32)
        boxes = np.concatenate((xy, xy + wh), axis=1).astype(np.float32)
        cls = rng.integers(0, 80, size=m).astype(np.float32)
        boxes_off = (boxes + (cls * np.float32(4096))[:, None]).astype(np.float32)
        scores = rng.uniform(0, 1, size=m).astype(np.float32)
        scores[rng.integers(0, m, size=m // 20)] = 0.0          # zero scores are never kept
        scores[rng.integers(0, m, size=m // 20)] = np.float32(0.5)  # ties -> lower index first
        store[f"nms_{tag}_boxes"] = boxes_off
        store[f"nms_{tag}_scores"] = scores
        for thr in (0.2, 0.5, 0.65):
            keep = utils.numba_nms(boxes_off, scores, thr)
            store[f"nms_{tag}_keep_{thr}"] = np.asarray(keep, dtype=np.int32)
        store[f"iou_{tag}"] = utils.numba_iou(boxes_off[:64], boxes_off[:256])
    # micro known-answer cases
    kat_boxes = np.array([[0, 0, 2, 1], [0, 0, 1, 1], [5, 5, 5, 5], [10, 10, 12, 12]], dtype=np.float32)
    store["kat_boxes"] = kat_boxes
    store["kat_iou"] = utils.numba_iou(kat_boxes, kat_boxes)
    store["kat_keep_0.5"] = np.asarray(utils.numba_nms(kat_boxes, np.array([.9, .8, .7, .6], np.float32), 0.5), np.int32)
    # torch IoU family (float32)
    xy = rng.uniform(0, 100, size=(200, 2)).astype(np.float32)
    wh = rng.uniform(1, 50, size=(200, 2)).astype(np.float32)
    b1 = torch.from_numpy(np.concatenate((xy, xy + wh), axis=1))
    xy2 = xy + rng.normal(0, 8, size=(200, 2)).astype(np.float32)
    wh2 = wh * np.exp(rng.normal(0, 0.3, size=(200, 2))).astype(np.float32)
    b2 = torch.from_numpy(np.concatenate((xy2, xy2 + wh2), axis=1).astype(np.float32))
    store["tiou_b1"], store["tiou_b2"] = b1.numpy(), b2.numpy()
    store["tiou_iou"] = utils.gpu_iou(b1[:50], b2).numpy()
    store["tiou_giou"] = utils.gpu_Giou(b1, b2).numpy()
    store["tiou_diou"] = utils.gpu_DIoU(b1, b2).numpy()
    store["tiou_ciou"] = utils.gpu_CIoU(b1, b2).numpy()
    store["tiou_giou_row"] = utils.gpu_Giou(b1[:1], b2).numpy()
    store["tiou_diou_row"] = utils.gpu_DIoU(b1[:1], b2).numpy()
    sc = torch.from_numpy(rng.uniform(0.01, 1, size=200).astype(np.float32))
    for kind in ("giou", "diou"):
        store[f"tnms_{kind}"] = np.asarray(utils.gpu_nms(b2, sc, kind, 0.45), dtype=np.int32)
    store["tnms_scores"] = sc.numpy()
    store["anchors_64x96"] = utils.GPUAnchor([64, 96])().cpu().numpy()
    # autograd through the row-wise IoU flavours (the loss-side callers, loss/yolov5_loss.py:110): d(sum(w * iou))/d boxes
    ga = b1[:64].clone()
    gb = b2[:64].clone()
    gb[0, 0] = ga[0, 0]                      # tie in max(x1, x1'): torch splits the gradient evenly
    gb[1] = ga[1]                            # identical boxes: every max/min ties
    gb[2, 2] = ga[2, 2]
    gb[3] = ga[3] + 500.0                    # disjoint: intersection clamped at 0
    gb[4, 0] = ga[4, 2]                      # touching: clamp input exactly 0 (gradient still passes)
    gb[4, 2] = gb[4, 0] + 10.0
    wts = torch.from_numpy(rng.uniform(-1, 1, size=64).astype(np.float32))
    store["grad_b1"], store["grad_b2"], store["grad_w"] = ga.numpy().copy(), gb.numpy().copy(), wts.numpy()
    for kind, fn in (("giou", utils.gpu_Giou), ("diou", utils.gpu_DIoU), ("ciou", utils.gpu_CIoU)):
        x1, x2 = ga.clone().requires_grad_(True), gb.clone().requires_grad_(True)
        (fn(x1, x2) * wts).sum().backward()
        store[f"grad_{kind}_d1"], store[f"grad_{kind}_d2"] = x1.grad.numpy(), x2.grad.numpy()
    for kind, fn in (("giou", utils.gpu_Giou), ("diou", utils.gpu_DIoU)):      # box1 broadcast over the rows
        x1, x2 = ga[:1].clone().requires_grad_(True), gb.clone().requires_grad_(True)
        (fn(x1, x2) * wts).sum().backward()
        store[f"grad_{kind}_row_d1"], store[f"grad_{kind}_row_d2"] = x1.grad.numpy(), x2.grad.numpy()
    # soft-NMS (utils/nms.py:68-140): dead code in the reference, runs for the giou/diou/ciou flavours with (M,1) scores
    m = 48
    sb = b2[:m].clone()
    ss = torch.from_numpy(rng.uniform(0.05, 1, size=(m, 1)).astype(np.float32))
    store["soft_boxes"], store["soft_scores"] = sb.numpy(), ss.numpy()
    for kind in ("giou", "diou", "ciou"):
        store[f"soft_linear_{kind}"] = utils.gpu_linear_soft_nms(sb, ss, kind, iou_threshold=0.1, thresh=0.4).numpy()
    store["soft_exp_diou"] = utils.gpu_exponential_soft_nms(sb[:12], ss[:12], "diou", 0.3, sigmma=0.5, thresh=0.001).numpy()
    # letterbox undo, val_yolov5.py:166-172 (the torch expressions of preds_postprocess on a (K,6) float32 tensor)
    pred = torch.from_numpy(rng.uniform(-20, 700, size=(40, 6)).astype(np.float32))
    pred[:, 2:4] = pred[:, 0:2] + torch.from_numpy(rng.uniform(1, 200, size=(40, 2)).astype(np.float32))
    store["lb_in"] = pred.numpy().copy()
    scale, pad_top, pad_left, org_h, org_w = 0.7339449541284404, 0, 85, 545, 640
    pred[:, [0, 2]] -= pad_left
    pred[:, [1, 3]] -= pad_top
    pred[:, [0, 1, 2, 3]] /= scale
    pred[:, [0, 2]] = pred[:, [0, 2]].clamp(1, org_w - 1)
    pred[:, [1, 3]] = pred[:, [1, 3]].clamp(1, org_h - 1)
    store["lb_out"] = pred.numpy()
    store["lb_info"] = np.array([scale, pad_top, pad_left, org_h, org_w], dtype=np.float64)
    return store


def main():
    utils, trainer = refharness.import_reference()
    os.makedirs(OUT, exist_ok=True)
    if len(sys.argv) <= 1 or "utils_nms_iou" in sys.argv[1:]:
        np.savez_compressed(os.path.join(OUT, "utils_nms_iou.npz"), **utils_case(utils))
    fcos_thr = {"compute_metric_cls_threshold": 0.2, "compute_metric_iou_threshold": 0.35, "max_predictions_per_img": 100}
    cases = [
        # name, family, dist, img, batch, seed, C, hyp overrides.  Few classes on the small images so that
        # same-class overlaps (suppression, the count filter) actually occur.
        ("yolov5_dense_c80_nopp", "yolov5", "dense", 64, 1, 10, 80, {"postprocess_bbox": False}),
        ("yolov5_dense", "yolov5", "dense", 64, 2, 11, 4, {}),
        ("yolov5_sparse", "yolov5", "sparse", 128, 2, 12, 6, {}),
        ("yolov5_crowd", "yolov5", "crowd", 128, 2, 13, 6, {}),
        ("yolov5_agnostic_off", "yolov5", "dense", 64, 1, 14, 6, {"agnostic": False, "postprocess_bbox": False}),
        ("yolov5_deploy_thr", "yolov5", "dense", 64, 2, 15, 6,
         {"compute_metric_conf_threshold": 0.3, "compute_metric_cls_threshold": 0.3, "compute_metric_iou_threshold": 0.2}),
        ("yolov5_maxdet", "yolov5", "dense", 96, 1, 16, 6, {"max_predictions_per_img": 20, "postprocess_bbox": False}),
        ("yolov7_dense", "yolov7", "dense", 64, 2, 21, 4, {}),
        ("yolov7_crowd", "yolov7", "crowd", 128, 1, 22, 6, {}),
        ("yolox_dense", "yolox", "dense", 64, 2, 31, 4, {}),
        ("yolox_dense_nopp", "yolox", "dense", 128, 1, 32, 6, {"postprocess_bbox": False}),
        ("yolov8_dense", "yolov8", "dense", 64, 2, 41, 4, {}),
        ("yolov8_sparse", "yolov8", "sparse", 64, 1, 42, 6, {}),
        ("retinanet_dense", "retinanet", "dense", 64, 2, 51, 4, {}),
        ("retinanet_sparse", "retinanet", "sparse", 96, 1, 52, 6, {}),
        ("retinanet_exp_dense", "retinanet_exp", "dense", 64, 1, 53, 4, {}),
        ("fcos_dense", "fcos", "dense", 128, 2, 61, 4, fcos_thr),
        ("fcos_sparse", "fcos", "sparse", 256, 1, 62, 6, dict(fcos_thr, compute_metric_cls_threshold=0.05)),
        # mutil_label: one record per (candidate, class) above the class threshold (eval_yolov5.py:276-279)
        ("yolov5_multilabel", "yolov5", "crowd", 128, 2, 71, 6, {"mutil_label": True, "compute_metric_cls_threshold": 0.05}),
        ("yolov5_multilabel_dense", "yolov5", "dense", 64, 1, 72, 4, {"mutil_label": True, "compute_metric_cls_threshold": 0.2}),
        ("yolov7_multilabel", "yolov7", "crowd", 128, 1, 73, 6, {"mutil_label": True, "compute_metric_cls_threshold": 0.05}),
        ("yolox_multilabel", "yolox", "dense", 64, 2, 74, 4, {"mutil_label": True, "compute_metric_cls_threshold": 0.3}),
        ("yolov8_multilabel", "yolov8", "dense", 64, 1, 75, 4, {"mutil_label": True, "compute_metric_cls_threshold": 0.5}),
        ("fcos_multilabel", "fcos", "dense", 128, 1, 76, 4, dict(fcos_thr, mutil_label=True)),
        # cpu_*: pinned on the oracle side only so far (not yet part of the GPU parametrisations, see tests/conftest.py)
        ("cpu_fcos_no_ctr", "fcos", "dense", 128, 2, 91, 4, dict(fcos_thr, thresh_with_ctr=False)),
        ("cpu_retinanet_exp_sparse", "retinanet_exp", "sparse", 96, 2, 92, 6, {}),
        ("cpu_yolov5_agnostic_off_multilabel", "yolov5", "crowd", 128, 1, 93, 6,
         {"agnostic": False, "mutil_label": True, "compute_metric_cls_threshold": 0.05}),
    ]
    only = set(sys.argv[1:])
    tta_cases = [
        ("tta_yolov5", "yolov5", 64, 96, 2, 81, 4, {}),
        ("tta_yolov7", "yolov7", 64, 96, 1, 82, 4, {}),
        ("tta_yolox", "yolox", 64, 96, 2, 83, 4, {}),
        ("tta_yolov8", "yolov8", 64, 96, 1, 84, 4, {}),
        ("tta_retinanet", "retinanet", 64, 96, 1, 85, 4, {}),
        ("tta_fcos", "fcos", 128, 128, 1, 86, 4, fcos_thr),
    ]
    for name, family, img_h, img_w, batch, seed, C, over in tta_cases:
        if only and name not in only:
            continue
        store = tta_case(trainer, family, img_h, img_w, batch, seed, C, **over)
        path = os.path.join(OUT, f"{name}.npz")
        np.savez_compressed(path, **store)
        print(f"{name:24s} rows={store['merged'].shape[1]:6d} counts={store['counts'].tolist()} "
              f"{os.path.getsize(path) / 1024:.0f} KiB")
    for name, family, dist, img, batch, seed, C, over in cases:
        if only and name not in only:
            continue
        store = evaluator_case(trainer, family, dist, img, batch, seed, C, **over)
        path = os.path.join(OUT, f"{name}.npz")
        np.savez_compressed(path, **store)
        print(f"{name:24s} N={store['decoded'].shape[1]:6d} counts={store['counts'].tolist()} "
              f"{os.path.getsize(path) / 1024:.0f} KiB")


if __name__ == "__main__":
    main()
