"""CPU restatement of the consumers right after the kept rows (TEST INFRASTRUCTURE ONLY; see oracle/__init__.py):

    map_iou(box1, box2)        utils/mAP.py:18-42
    compute_tp(gt, pred)       utils/mAP.py:70-100   (mAP_v2.compute_tp)

Pinned by tests/golden/utils_extra.npz (outputs of the reference's own functions, oracle/gen_golden.py::utils_extra_case).
The restatement keeps numpy's promoted precision (float32 stays float32) and replaces the reference's unstable
``argsort()[::-1]`` by a STABLE ascending sort read backwards -- what numpy's default sort does up to 16 elements and the
documented tie rule of ysb_compute_tp beyond.
"""
import numpy as np

IOU_THRESHOLDS = np.linspace(0.5, 0.95, 10)


def map_iou(box1, box2):
    """utils/mAP.py:18-42: (M,4), (N,4) -> (M,N) in the operands' promoted dtype."""
    box1 = np.expand_dims(np.asarray(box1), axis=1)
    box2 = np.asarray(box2)
    a1 = np.prod(box1[..., [2, 3]] - box1[..., [0, 1]], axis=-1)
    a2 = np.prod(box2[:, [2, 3]] - box2[:, [0, 1]], axis=-1)
    w = np.maximum(0., np.minimum(box1[..., 2], box2[:, 2]) - np.maximum(box1[..., 0], box2[:, 0]))
    h = np.maximum(0., np.minimum(box1[..., 3], box2[:, 3]) - np.maximum(box1[..., 1], box2[:, 1]))
    inter = w * h
    return inter / np.clip(a1 + a2 - inter, a_min=1e-6, a_max=10000000)


def compute_tp(gt, pred, iou_thr=IOU_THRESHOLDS):
    """utils/mAP.py:70-100, step by step, with the stable tie rule."""
    gt, pred = np.asarray(gt), np.asarray(pred)
    tp = np.zeros((pred.shape[0], len(iou_thr)), dtype=bool)
    if gt.shape[0] == 0 or pred.shape[0] == 0:
        return tp
    ious = map_iou(gt[:, :4], pred[:, :4])
    mask = (ious >= iou_thr[0]) & (gt[:, [4]] == pred[:, 5])
    if mask.sum() > 0:
        gt_i, pred_i = np.nonzero(mask)
        match = np.concatenate((np.stack((gt_i, pred_i), axis=1), ious[mask][:, None]), axis=1)
        if mask.sum() > 1:
            match = match[match[:, 2].argsort(kind="stable")[::-1]]
            match = match[np.unique(match[:, 1], return_index=True)[1]]
            match = match[np.unique(match[:, 0], return_index=True)[1]]
        tp[match[:, 1].astype(np.int32)] = match[:, [2]] >= iou_thr
    return tp
