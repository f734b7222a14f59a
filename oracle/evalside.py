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


# ---- weighted-box-fusion --------------------------------------------------------------------------------------------
def cpu_iou(bbox1, bbox2):
    """utils/bbox_tools.py:63-84 (bbox1 (1,4) float32, bbox2 (F,4) float64 in the fusion loop)."""
    a1 = np.prod(bbox1[:, [2, 3]] - bbox1[:, [0, 1]], axis=-1)
    a2 = np.prod(bbox2[:, [2, 3]] - bbox2[:, [0, 1]], axis=-1)
    ymax = np.minimum(bbox1[:, 3], bbox2[:, 3])
    xmax = np.minimum(bbox1[:, 2], bbox2[:, 2])
    ymin = np.maximum(bbox1[:, 1], bbox2[:, 1])
    xmin = np.maximum(bbox1[:, 0], bbox2[:, 0])
    inter = np.maximum(0., xmax - xmin) * np.maximum(0., ymax - ymin)
    return inter / np.clip(a1 + a2 - inter, a_min=1e-6, a_max=None)


def _update_fusion(cluster_bbox):
    """utils/weighted_fusion_bbox.py:41-60, literally (mean of box*score/sum(score): divided by the member count again)."""
    out = []
    for members in cluster_bbox:
        arr = np.array(members)
        bbox, score, lab, w = arr[:, :4], arr[:, 4], arr[:, 5], arr[:, 6]
        wb = bbox * score.reshape(-1, 1)
        wb /= np.sum(score)
        wb = np.mean(wb, axis=0)
        ws = np.sum(score * w) / np.sum(w)
        out.append(np.append(wb, [ws, lab[0]]))
    return out


def weighted_fusion_bbox(bbox_list, iou_thr=0.5):
    """utils/weighted_fusion_bbox.py:63-96 -> (Cluster, Fusion); ``argsort()[::-1]`` made deterministic with a stable sort
    (numpy's default is stable up to 16 elements; beyond that the reference's tie order is unspecified)."""
    bbox_list = np.asarray(bbox_list)
    Cluster, Fusion = [], []
    for lab in np.unique(bbox_list[:, 5]):
        bbox = bbox_list[bbox_list[:, 5] == lab]
        sort_index = np.argsort(bbox[:, 4], kind="stable")[::-1]
        fusion_bbox = [bbox[sort_index[0]][:7].tolist()]
        cluster_bbox = [[]]
        for i in sort_index:
            cur = bbox[i]
            ious = cpu_iou(np.array(cur)[:4][None, :], np.array(fusion_bbox)[:, :4])
            hit = np.greater_equal(ious, iou_thr).nonzero()[0]
            if len(hit) == 0:
                fusion_bbox.append(cur.tolist())
                cluster_bbox.append([cur.tolist()])
            else:
                for j in hit:
                    cluster_bbox[j].append(cur.tolist())
            if any(len(c) == 0 for c in cluster_bbox):
                raise IndexError("too many indices for array")   # what np.array([])[:, :4] raises in the reference
            fusion_bbox = _update_fusion(cluster_bbox)
        Cluster.append(cluster_bbox)
        Fusion.append(fusion_bbox)
    return Cluster, Fusion


def do_wfb(preds_out, weights, skip_thr, iou_thr, multi_label=False):
    """trainer/eval_yolov5.py:44-92 (the same method in eval_yolov7.py / eval_yolox.py) on decoded numpy tensors
    [(b, X, 5 + C), ...] -> list over images of the reference's Fusion nesting, or None."""
    F32 = np.float32
    bs = preds_out[0].shape[0]
    per_img = [[] for _ in range(bs)]
    for preds, weight in zip(preds_out, weights):
        preds = np.asarray(preds, dtype=F32)
        for j in range(bs):
            x = preds[j][preds[j][:, 4] > F32(skip_thr)].copy()
            if x.shape[0] == 0:
                continue
            x[:, 5:] *= x[:, 4:5]
            box = np.stack((x[:, 0] - x[:, 2] / F32(2), x[:, 1] - x[:, 3] / F32(2), x[:, 0] + x[:, 2] / F32(2),
                            x[:, 1] + x[:, 3] / F32(2)), axis=1)
            if multi_label:
                r, c = (x[:, 5:] > F32(skip_thr)).nonzero()
                rows = np.concatenate((box[r], x[r, c + 5][:, None], c[:, None].astype(F32)), axis=1)
            else:
                c = x[:, 5:].argmax(axis=1)
                conf = x[np.arange(x.shape[0]), c + 5]
                rows = np.concatenate((box, conf[:, None], c[:, None].astype(F32)), axis=1)[conf > F32(skip_thr)]
            if rows.shape[0] == 0:
                continue
            per_img[j].append(np.concatenate((rows, np.full((rows.shape[0], 1), weight, dtype=F32)), axis=1).astype(F32))
    out = []
    for j in range(bs):
        out.append(weighted_fusion_bbox(np.vstack(per_img[j]), iou_thr)[1] if per_img[j] else None)
    return out
