"""Child process of tests/test_dropin_reference.py: import the UNMODIFIED reference (checkout or staged copy), run
dropin.install() on it and report what happened as one JSON line.  ``run`` additionally drives the reference's own
call shape (val_yolov5.py:106-107, 166-172, 388: evaluator -> preds rows -> mAP_v2.compute_tp) on a CUDA device."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from oracle import refharness  # noqa: E402


def main():
    mode = sys.argv[1]
    utils, trainer = refharness.import_reference()
    import trainer.eval_yolov5 as ev5          # the reference's modules, by their own names
    import trainer.eval_yolox as evx
    import utils.bbox_tools as ubt
    import utils.nms as unms
    before = dict(numba_nms=utils.numba_nms, cls=trainer.YOLOV5Evaluator, tp=utils.mAP_v2.compute_tp, giou=utils.gpu_Giou)
    assert ev5.YOLOV5Evaluator is trainer.YOLOV5Evaluator
    from yoloseries_b200 import dropin, trainer as ours, utils as our_utils
    done = dropin.install(patch_iou=(mode == "run"))
    rep = {"done": done, "root": refharness.REFERENCE_ROOT}
    rep["utils_rebound"] = all(getattr(utils, n) is getattr(our_utils, n) for n in
                               ("numba_nms", "gpu_nms", "numba_iou", "gpu_linear_soft_nms", "gpu_exponential_soft_nms"))
    rep["submodules_rebound"] = unms.numba_nms is our_utils.numba_nms and ubt.numba_iou is our_utils.numba_iou
    rep["eval_modules_rebound"] = all(getattr(m, n) is getattr(our_utils, n) for m in (ev5, evx)
                                      for n in ("numba_nms", "numba_iou") if hasattr(m, n))
    rep["evaluators_rebound"] = (trainer.YOLOV5Evaluator is ours.YOLOV5Evaluator and ev5.YOLOV5Evaluator is ours.YOLOV5Evaluator
                                 and trainer.YOLOXEvaluator is ours.YOLOXEvaluator and evx.YOLOXEvaluator is ours.YOLOXEvaluator)
    rep["compute_tp_rebound"] = utils.mAP_v2.compute_tp is not before["tp"]
    rep["old_still_bound"] = utils.numba_nms is before["numba_nms"] or trainer.YOLOV5Evaluator is before["cls"]
    if mode == "run":
        import numpy as np
        import torch
        from conftest import golden_heads, load_golden
        from yoloseries_b200.synth import V5_ANCHORS_PX
        rep["iou_family_rebound"] = utils.gpu_Giou is our_utils.gpu_Giou and ubt.gpu_iou is our_utils.gpu_iou
        g = load_golden("yolov5_crowd")
        meta = g["meta"]
        hyp = refharness.reference_hyp((meta["img"], meta["img"]), num_class=meta["num_class"], device="cuda")
        for k in ("compute_metric_conf_threshold", "compute_metric_cls_threshold", "compute_metric_iou_threshold",
                  "max_predictions_per_img", "postprocess_bbox", "mutil_label", "agnostic"):
            hyp[k] = meta[k]
        heads = [torch.from_numpy(h).cuda() for h in golden_heads(g)]
        # val_yolov5.py:106-107: the evaluator is built through the reference's package attribute
        validater = trainer.YOLOV5Evaluator(lambda x: [h.clone() for h in heads], torch.tensor(V5_ANCHORS_PX), hyp,
                                            compute_metric=True)
        x = torch.zeros(meta["batch"], 3, meta["img"], meta["img"], device="cuda")
        outs = validater(x)
        ok = True
        rep["ref_counts"] = [int(c) for c in g["counts"]]
        rep["max_rel_err"] = 0.0
        for i, o in enumerate(outs):
            c = int(g["counts"][i])
            ok &= (o is None) == (c < 0)
            if o is not None:
                # fused CUDA decode vs the golden's CPU ATen decode: north-star tolerance on the values, same rows kept
                ref = g["rows"][i, :c]
                ok &= isinstance(o, torch.Tensor) and o.device.type == "cpu" and tuple(o.shape) == ref.shape
                if tuple(o.shape) == ref.shape and ref.size:
                    # corners against the un-cancelled xywh operands (oracle/parity.py: x1 = cx - w/2)
                    den = np.maximum(np.abs(ref.astype(np.float64)), 1.0)
                    den[:, [0, 2]] = np.maximum(den[:, [0]], den[:, [2]])
                    den[:, [1, 3]] = np.maximum(den[:, [1]], den[:, [3]])
                    err = float((np.abs(o.numpy().astype(np.float64) - ref) / den).max())
                    rep["max_rel_err"] = max(rep["max_rel_err"], err)
                    ok &= err <= 1e-5
        rep["rows_equal_reference"] = bool(ok)
        rep["kept"] = [(-1 if o is None else int(o.shape[0])) for o in outs]
        # val_yolov5.py:388: the reference's own mAP_v2 class, whose compute_tp now runs on the device
        e = load_golden("utils_extra")
        import tempfile
        mv2 = utils.mAP_v2([e["tp_gt_0"], e["tp_gt_3"]], [e["tp_pred_0"], e["tp_pred_3"]], tempfile.mkdtemp(prefix="ysb_map_"))
        rep["compute_tp_equal_reference"] = bool(np.array_equal(mv2.compute_tp(mv2.gt[0], mv2.pred[0]), e["tp_out_0"])
                                                 and np.array_equal(mv2.compute_tp(mv2.gt[1], mv2.pred[1]), e["tp_out_3"]))
        # the reference's NMS entry point by its own name
        u = load_golden("utils_nms_iou")
        rep["numba_nms_equal_reference"] = utils.numba_nms(u["nms_a_boxes"], u["nms_a_scores"], 0.5) == u["nms_a_keep_0.5"].tolist()
    print(json.dumps(rep))


if __name__ == "__main__":
    main()
