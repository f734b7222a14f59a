"""TEST INFRASTRUCTURE ONLY -- numpy restatement of the test-time-augmentation merge of the reference's evaluators
(``test_time_augmentation``: trainer/eval_yolov5.py:152-179 for the xywh families v5 / v7 / YOLOX,
trainer/eval_yolov8.py:40-73, eval_retinanet.py:147-182 and eval_fcos.py:90-123 for the xyxy families).
Never imported by the product path.  Pinned by tests/golden/tta_*.npz (outputs of the reference itself).
"""
import numpy as np

F = np.float32
TTA_SCALES = (1, 0.83, 0.67)
TTA_FLIPS = (None, 2, 3)
XYWH_FAMILIES = ("yolov5", "yolov7", "yolox")


def box_col(family, num_class):
    return num_class if family.startswith("retinanet") else 0


def undo_pass(family, decoded, scale, flip, img_h, img_w, num_class):
    """One pass: ``preds[..., box:box+4] /= s`` (float32 true division) and the flip undo.  Returns a copy."""
    out = np.array(decoded, dtype=F, copy=True)
    b0 = box_col(family, num_class)
    out[..., b0:b0 + 4] = out[..., b0:b0 + 4] / F(scale)
    H, W = F(img_h), F(img_w)
    if family in XYWH_FAMILIES:
        if flip == 2:
            out[..., b0 + 1] = H - out[..., b0 + 1]
        if flip == 3:
            out[..., b0 + 0] = W - out[..., b0 + 0]
    else:
        if flip == 2:
            ymin, ymax = H - out[..., b0 + 3], H - out[..., b0 + 1]
            out[..., b0 + 1], out[..., b0 + 3] = ymin, ymax
        if flip == 3:
            xmin, xmax = W - out[..., b0 + 2], W - out[..., b0 + 0]
            out[..., b0 + 0], out[..., b0 + 2] = xmin, xmax
    return out


def tta_merge(family, decoded_passes, img_h, img_w, num_class, scales=TTA_SCALES, flips=TTA_FLIPS):
    """[(b, N_i, C') decoded tensor of pass i] -> (b, sum N_i, C'), torch.cat(aug_preds, dim=1)."""
    return np.concatenate([undo_pass(family, d, s, f, img_h, img_w, num_class)
                           for d, s, f in zip(decoded_passes, scales, flips)], axis=1)
