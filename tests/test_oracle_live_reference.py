"""Oracle vs the UNMODIFIED reference executed live, on seeds that are not in tests/golden/ (build container only:
skipped where /root/reference does not exist, e.g. on the GPU box).  Runs in a subprocess so that the reference's
top-level ``utils`` / ``trainer`` packages never enter this process."""
import json
import os
import subprocess
import sys

import pytest

from oracle import refharness

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

CHILD = r'''
import json, sys
import numpy as np
import torch
sys.path.insert(0, sys.argv[1])
import oracle
from oracle import refharness
from oracle.gen_golden import evaluator_case
utils, trainer = refharness.import_reference()
out = {}
# utils.numba_nms / numba_iou (utils/nms.py:10-27, utils/bbox_tools.py:12-35) on fresh seeds, ties and zero scores included
ok = True
for seed in (101, 102, 103):
    rng = np.random.default_rng(seed)
    m = int(rng.integers(50, 400))
    xy = rng.uniform(0, 150, size=(m, 2)).astype(np.float32)
    wh = rng.uniform(3, 70, size=(m, 2)).astype(np.float32)
    boxes = np.concatenate((xy, xy + wh), 1)
    boxes += (rng.integers(0, 5, size=m).astype(np.float32) * np.float32(4096))[:, None]
    scores = np.round(rng.uniform(0, 1, size=m), 2).astype(np.float32)       # two decimals: many ties
    scores[rng.integers(0, m, size=m // 10)] = 0.0
    for thr in (0.3, 0.65):
        ok &= list(utils.numba_nms(boxes, scores, thr)) == oracle.numba_nms(boxes, scores, thr)
    ok &= np.array_equal(utils.numba_iou(boxes[:40], boxes), oracle.numba_iou(boxes[:40], boxes), equal_nan=True)
out["utils"] = bool(ok)
# evaluators: the oracle's numba_nms restatement on the reference's own decoded tensor == the reference's rows
fcos_thr = {"compute_metric_cls_threshold": 0.2, "compute_metric_iou_threshold": 0.35, "max_predictions_per_img": 100}
for fam, dist, img, seed, over in (("yolov5", "dense", 64, 901, {}), ("yolov7", "crowd", 128, 902, {}),
                                   ("yolox", "sparse", 128, 903, {}), ("fcos", "dense", 128, 904, fcos_thr)):
    store = evaluator_case(trainer, fam, dist, img, 2, seed, 4, **over)
    import ast
    meta = ast.literal_eval(str(store["meta"]))
    hyp = oracle.default_hyp(
        num_class=meta["num_class"], conf_threshold=meta["compute_metric_conf_threshold"],
        cls_threshold=meta["compute_metric_cls_threshold"], iou_threshold=meta["compute_metric_iou_threshold"],
        max_predictions_per_img=meta["max_predictions_per_img"], min_prediction_box_wh=meta["min_prediction_box_wh"],
        mutil_label=meta["mutil_label"], agnostic=meta["agnostic"], postprocess_bbox=meta["postprocess_bbox"],
        pre_nms_topk=meta["pre_nms_topk"], pre_nms_thresh=meta["pre_nms_thresh"], thresh_with_ctr=meta["thresh_with_ctr"])
    res = oracle.evaluator_nms(fam, store["decoded"], hyp, full_nms=True)
    good = True
    for i, r in enumerate(res):
        cnt = int(store["counts"][i])
        if cnt < 0:
            good &= r.rows is None
        else:
            good &= r.rows is not None and np.array_equal(r.rows, store["rows"][i, :cnt])
    out[fam] = bool(good)
print("RESULT " + json.dumps(out))
'''


@pytest.mark.skipif(not refharness.available(), reason="the reference checkout exists only in the build container")
def test_oracle_equals_live_reference_on_fresh_seeds():
    r = subprocess.run([sys.executable, "-c", CHILD, ROOT], cwd=ROOT, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-3000:]
    line = [ln for ln in r.stdout.splitlines() if ln.startswith("RESULT ")][-1]
    res = json.loads(line[len("RESULT "):])
    assert res and all(res.values()), res
