"""Import the UNMODIFIED reference: from /root/reference in the build container, else from the byte-for-byte
staged copy baseline/_ref/ (git-ignored, shipped with the gpurun snapshot; baseline/stage_reference.py).

TEST INFRASTRUCTURE ONLY.  Used by oracle/gen_golden.py to produce tests/golden/*.npz, by the parity tests that run
the reference itself (CPU here, ``hyp['device']='cuda'`` on the GPU box = SURVEY.md 8c mode B), and by
``bench.py --impl reference`` (the reference arm).  Never imported by the product package.

The reference star-imports plotting modules that are not installed (matplotlib, seaborn,
emoji); they are irrelevant to the hot path and are replaced by MagicMock stubs
(SURVEY.md section 8c).
"""
import os
import sys
from unittest.mock import MagicMock

_REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
STAGED_ROOT = os.path.join(_REPO, "baseline", "_ref")


def _resolve_root():
    env = os.environ.get("YSB_REFERENCE_ROOT")
    if env:
        return env
    if os.path.isdir("/root/reference/trainer"):
        return "/root/reference"
    return STAGED_ROOT


REFERENCE_ROOT = _resolve_root()
_STUBS = ["matplotlib", "matplotlib.pyplot", "matplotlib.patches", "matplotlib.font_manager",
          "matplotlib.colors", "seaborn", "emoji"]


def available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "trainer"))


def import_reference():
    """Returns the reference's (utils, trainer) packages."""
    if not available():
        raise RuntimeError(f"reference checkout not found at {REFERENCE_ROOT}")
    for name in _STUBS:
        sys.modules.setdefault(name, MagicMock())
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    import trainer  # noqa: E402  (reference package)
    import utils  # noqa: E402  (reference package)
    return utils, trainer


def make_evaluator(trainer, family, model, hyp, compute_metric=True):
    """The reference's evaluator of ``family`` around ``model`` (any callable returning head tensors)."""
    import torch

    from yoloseries_b200 import synth
    cls = {"yolov5": "YOLOV5Evaluator", "yolov7": "YOLOV7Evaluator", "yolox": "YOLOXEvaluator", "yolov8": "YOLOV8Evaluator",
           "retinanet": "RetinaNetEvaluator", "retinanet_exp": "RetinaNetEvaluatorExperiment", "fcos": "FCOSEvaluator"}[family]
    if family in ("yolov5", "yolov7"):
        anchors = torch.tensor(synth.V5_ANCHORS_PX, device=hyp["device"])
        return getattr(trainer, cls)(model, anchors, hyp, compute_metric=compute_metric)
    return getattr(trainer, cls)(model, hyp, compute_metric=compute_metric)


def clone_heads(x):
    import torch
    if isinstance(x, torch.Tensor):
        return x.clone()
    if isinstance(x, (list, tuple)):
        return type(x)(clone_heads(v) for v in x)
    if isinstance(x, dict):
        return type(x)((k, clone_heads(v)) for k, v in x.items())
    return x


def head_model(family, heads):
    """A stand-in model: every forward hands out fresh clones of ``heads`` in the container type the reference's
    models use (the reference mutates model outputs in place for YOLOv7 / RetinaNet)."""
    from collections import OrderedDict

    def model(_x):
        h = clone_heads(heads)
        if family in ("yolov7", "yolox", "yolov8"):
            return OrderedDict((f"p{i}", t) for i, t in enumerate(h))
        return h
    return model


def reference_hyp(input_hw, **over):
    """A flat ``hyp`` dict with the hot-path keys, mAP-profile thresholds."""
    hyp = dict(
        device="cpu", num_class=80, input_img_size=list(input_hw), use_tta=False, wfb=False, half=False,
        iou_threshold=0.65, conf_threshold=0.001, cls_threshold=0.001,
        compute_metric_iou_threshold=0.65, compute_metric_conf_threshold=0.001,
        compute_metric_cls_threshold=0.001, max_predictions_per_img=300, min_prediction_box_wh=2,
        iou_type="iou", mutil_label=False, agnostic=True, postprocess_bbox=True, num_anchors=1, reg=16,
        tar_box_scale_factor=[0.1, 0.1, 0.2, 0.2], pre_nms_topk=1000, pre_nms_thresh=0.05, thresh_with_ctr=True,
    )
    hyp.update(over)
    return hyp
