"""Import the UNMODIFIED reference from /root/reference (build container only).

TEST INFRASTRUCTURE ONLY.  Used by oracle/gen_golden.py to produce tests/golden/*.npz and
by tests that are skipped when /root/reference is absent (it does not exist on the GPU
box; nothing in the ``-m gpu`` tests, smoke() or bench.py touches this module).

The reference star-imports plotting modules that are not installed (matplotlib, seaborn,
emoji); they are irrelevant to the hot path and are replaced by MagicMock stubs
(SURVEY.md section 8c).
"""
import os
import sys
from unittest.mock import MagicMock

REFERENCE_ROOT = os.environ.get("YSB_REFERENCE_ROOT", "/root/reference")
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
