"""oracle/ -- TEST INFRASTRUCTURE ONLY.

A CPU restatement (plain C for the NMS/IoU loops, numpy for the decode and filter
arithmetic) of the detection post-processing hot path of yl-jiang/YOLOSeries:
``trainer/eval_*.py`` (``do_inference`` + ``numba_nms`` methods), ``utils/nms.py`` and the IoU
routines of ``utils/bbox_tools.py``.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import this package, and only as the checker / reported CPU
baseline.  The product package ``yoloseries_b200`` never imports it and has no CPU fallback.

Parity status: PINNED against the unmodified reference executed in the build container
(``oracle/gen_golden.py`` -> ``tests/golden/*.npz``).  The reference itself has no tests or
golden vectors, and its arithmetic rests on unpinned third-party
versions (torch>=1.8.1, numba>=0.54, numpy>=1.20); the fixtures were generated with
torch 2.11.0, numba 0.65.0, numpy 2.3.5.
"""
from .cnms import (  # noqa: F401
    numba_iou,
    numba_nms,
    gpu_iou_f32,
    gpu_nms_iou,
    postprocess_count,
    load_library,
)
from .decode import (  # noqa: F401
    sigmoid_f32,
    decode_yolov5,
    decode_yolov7,
    decode_yolox,
    decode_yolov8,
    decode_retinanet,
    decode_fcos,
    retinanet_anchors,
    retinanet_base_anchors,
    V5_ANCHORS,
)
from .pipeline import FAMILY_RULES, FamilyRules, default_hyp, evaluator_nms, ImageResult  # noqa: F401
from .softnms import (giou, diou, ciou, linear_soft_nms, exponential_soft_nms, undo_letterbox,  # noqa: F401
                      iou_backward, pairwise_iou_backward)
from . import evalside  # noqa: F401
from .evalside import map_iou, compute_tp, weighted_fusion_bbox, do_wfb  # noqa: F401
from .tta import tta_merge, undo_pass, TTA_SCALES, TTA_FLIPS  # noqa: F401
