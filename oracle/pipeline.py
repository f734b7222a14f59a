"""Restatement of the ``numba_nms`` METHOD of every evaluator (filter -> class pick ->
class offset -> greedy NMS -> max_det -> box post-filter -> rows), one rule table instead
of seven near-identical copies.  TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

Reference text per family:
  yolov5        trainer/eval_yolov5.py:261-316
  yolov7        trainer/eval_yolov7.py:203-282
  yolox         trainer/eval_yolox.py:201-258
  yolov8        trainer/eval_yolov8.py:168-226
  retinanet     trainer/eval_retinanet.py:297-353
  retinanet_exp trainer/eval_retinanet_experiment.py:307-363
  fcos          trainer/eval_fcos.py:225-307

numpy >= 2 compares a float32 array with a Python float in float32 (NEP 50), so every
score threshold is rounded to float32 first; the IoU threshold stays a float64 because
numba compares float64 IoUs with it (utils/nms.py:22) and so does the post-filter.
"""
from dataclasses import dataclass
from typing import List, Optional

import numpy as np

from . import cnms

F32 = np.float32


@dataclass(frozen=True)
class FamilyRules:
    name: str
    box_col: int           # first of the 4 box columns in a decoded row
    box_is_xywh: bool      # centre/size rows need numba_xywh2xyxy (utils/bbox_tools.py:137-148)
    obj_col: Optional[int]  # objectness / centerness / conf column, or None (negative = from the end)
    cls_col: int           # first class column
    pre_mask: str          # 'obj' | 'obj_x_maxcls' | 'maxcls' | 'none' | 'any_cls_gt_pre'
    pre_thr: str           # hyp key compared against in the pre-mask
    post_strict: bool      # score > cls_thr (True) or >= (False)
    small_box_filter: bool  # remove_small_boxes after NMS (eval_yolov7.py:203-213, eval_fcos.py:225-234)
    merge_boxes: bool      # RetinaNet writes the weighted-mean boxes into the output rows
    none_when_empty: bool  # v7/FCOS: an empty result is None, others return a (0, 6) array
    topk_sqrt: bool        # FCOS: argsort desc -> first pre_nms_topk -> score = sqrt(score)
    window_hi_inclusive_300: bool  # FCOS post-filter window is 1 < M <= 300, others 1 < M < 3000


FAMILY_RULES = {
    "yolov5": FamilyRules("yolov5", 0, True, 4, 5, "obj", "conf", True, False, False, False, False, False),
    "yolov7": FamilyRules("yolov7", 0, True, 4, 5, "obj_x_maxcls", "conf", False, True, False, True, False, False),
    "yolox": FamilyRules("yolox", 0, True, 4, 5, "obj_x_maxcls", "conf", False, False, False, False, False, False),
    "yolov8": FamilyRules("yolov8", 0, False, None, 4, "maxcls", "cls", False, False, False, False, False, False),
    "retinanet": FamilyRules("retinanet", -4, False, None, 0, "none", "cls", True, False, True, False, False, False),
    "retinanet_exp": FamilyRules("retinanet_exp", -5, False, -1, 0, "obj", "conf", True, False, True, False, False,
                                 False),
    "fcos": FamilyRules("fcos", 0, False, 4, 5, "any_cls_gt_pre", "pre", True, True, False, True, True, True),
}


def default_hyp(**over):
    """The mAP-profile thresholds (config/train_yolov5.yaml:85-111 'compute_metric_*', validation.yaml:3-20)."""
    hyp = dict(
        num_class=80, conf_threshold=0.001, cls_threshold=0.001, iou_threshold=0.65,
        max_predictions_per_img=300, min_prediction_box_wh=2, mutil_label=False, agnostic=True,
        postprocess_bbox=True, pre_nms_topk=1000, pre_nms_thresh=0.05, thresh_with_ctr=True,
    )
    hyp.update(over)
    return hyp


@dataclass
class ImageResult:
    rows: Optional[np.ndarray]      # (K, 6) float32 [x1, y1, x2, y2, score, cls] or None
    cand_index: np.ndarray          # (K,) original candidate index n of every output row
    survivors: np.ndarray           # (M,) candidate indices entering NMS, in NMS array order
    nms_boxes: np.ndarray           # (M, 4) float32 offset boxes handed to numba_nms
    nms_scores: np.ndarray          # (M,) float32
    keep: List[int]                 # indices into the NMS arrays after max_det truncation
    keep_after_filter: np.ndarray   # after the postprocess_bbox count filter


def _xywh2xyxy(b):
    out = np.zeros_like(b)
    out[:, 0] = b[:, 0] - b[:, 2] / F32(2)
    out[:, 1] = b[:, 1] - b[:, 3] / F32(2)
    out[:, 2] = b[:, 0] + b[:, 2] / F32(2)
    out[:, 3] = b[:, 1] + b[:, 3] / F32(2)
    return out


def evaluator_nms(family, decoded, hyp, full_nms=False):
    """Per-image results for a decoded ``(b, N, C')`` float32 array.

    ``full_nms=True`` runs the greedy loop to exhaustion like the reference; the default
    stops after ``max_predictions_per_img`` keeps, which yields the same truncated list
    (prefix stability, tests/test_oracle.py::test_nms_prefix_stable).
    """
    rules = FAMILY_RULES[family]
    decoded = np.asarray(decoded, dtype=F32)
    C = hyp["num_class"]
    width = decoded.shape[-1]
    box0 = rules.box_col if rules.box_col >= 0 else C  # retinanet rows: [cls..., box4(, conf)]
    obj_col = None if rules.obj_col is None else (rules.obj_col if rules.obj_col >= 0 else width + rules.obj_col)
    conf_thr = F32(hyp.get("conf_threshold", 0.0))
    cls_thr = F32(hyp["cls_threshold"])
    iou_thr = float(hyp["iou_threshold"])
    max_det = int(hyp["max_predictions_per_img"])
    results = []
    for img in decoded:
        cls_all = img[:, rules.cls_col: rules.cls_col + C]
        obj_all = img[:, obj_col] if obj_col is not None else None
        if rules.pre_mask == "obj":
            pre = obj_all >= conf_thr
        elif rules.pre_mask == "obj_x_maxcls":
            pre = (obj_all * cls_all.max(axis=1)) >= conf_thr
        elif rules.pre_mask == "maxcls":
            pre = cls_all.max(axis=1) >= cls_thr
        elif rules.pre_mask == "any_cls_gt_pre":
            pre = (cls_all > F32(hyp["pre_nms_thresh"])).any(axis=1)
        else:
            pre = np.ones(img.shape[0], dtype=bool)
        cand = np.nonzero(pre)[0]
        empty = ImageResult(None, np.zeros(0, np.int64), np.zeros(0, np.int64), np.zeros((0, 4), F32),
                            np.zeros(0, F32), [], np.zeros(0, np.int64))
        if cand.size == 0:
            results.append(empty)
            continue
        x = img[cand]
        cls = x[:, rules.cls_col: rules.cls_col + C].copy()
        if obj_col is not None and (family != "fcos" or hyp["thresh_with_ctr"]):
            cls = cls * x[:, obj_col: obj_col + 1]
        box = x[:, box0: box0 + 4]
        box = _xywh2xyxy(box) if rules.box_is_xywh else box.copy()
        if hyp["mutil_label"]:
            if family.startswith("retinanet"):
                raise NotImplementedError("mutil_label is broken in the reference RetinaNet path (SURVEY 8a-2)")
            sel = (cls > cls_thr) if family == "fcos" else (cls >= cls_thr)
            rows_i, cols_i = np.nonzero(sel)
            score, cid, box, cand = cls[rows_i, cols_i], cols_i.astype(F32), box[rows_i], cand[rows_i]
            n_pre = min(int((cls_all > F32(hyp["pre_nms_thresh"])).sum()), hyp["pre_nms_topk"]) if family == "fcos" else 0
        else:
            score = cls.max(axis=1)
            cid = cls.argmax(axis=1).astype(F32)
            ok = (score > cls_thr) if rules.post_strict else (score >= cls_thr)
            score, cid, box, cand = score[ok], cid[ok], box[ok], cand[ok]
            n_pre = min(int(pre.sum()), hyp["pre_nms_topk"]) if family == "fcos" else 0
        m = score.shape[0]
        if m == 0:
            results.append(empty)
            continue
        if rules.topk_sqrt:
            # eval_fcos.py:272-281 -- np.argsort()[::-1] is not stable; ties are resolved here as
            # (score desc, index asc), which the reference does not pin.
            order = np.lexsort((np.arange(m), -score.astype(np.float64)))[:n_pre]
            score, cid, box, cand = score[order], cid[order], box[order], cand[order]
        offset = (cid * F32(4096)) if hyp["agnostic"] else (cid * F32(0))
        boxes_off = (box + offset[:, None]).astype(F32)
        if rules.topk_sqrt:
            score = np.sqrt(score)
        keep = cnms.numba_nms(boxes_off, score, iou_thr, 0 if full_nms else max_det)
        keep = keep[:max_det]
        keep_f = np.asarray(keep, dtype=np.int64)
        out_box = box
        in_window = (1 < m <= 300) if rules.window_hi_inclusive_300 else (1 < m < 3000)
        if hyp["postprocess_bbox"] and in_window and len(keep):
            flags, merged = cnms.postprocess_count(boxes_off, score, keep_f, iou_thr, raw_boxes=box,
                                                   merge=rules.merge_boxes)
            if rules.merge_boxes:
                out_box = box.copy()
                out_box[keep_f] = merged
            keep_f = keep_f[flags]
        rows = np.concatenate((out_box[keep_f], score[keep_f, None], cid[keep_f, None]), axis=1).astype(F32)
        cidx = cand[keep_f]
        if rules.small_box_filter:
            min_wh = F32(hyp["min_prediction_box_wh"])
            good = ((rows[:, 2] - rows[:, 0]) > min_wh) & ((rows[:, 3] - rows[:, 1]) > min_wh)
            rows, cidx = rows[good], cidx[good]
        if rows.shape[0] == 0 and rules.none_when_empty:
            rows = None
        results.append(ImageResult(rows, cidx, cand, boxes_off, score, list(keep), keep_f))
    return results
