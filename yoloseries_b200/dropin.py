"""Install the B200 engine behind the reference's own names (for a checkout of yl-jiang/YOLOSeries on sys.path).

    import yoloseries_b200.dropin as dropin
    dropin.install()          # after `import utils, trainer` of the reference

Evaluators bind ``numba_nms`` & co. at import time (``from utils import ...``, trainer/eval_yolov5.py:4-7), so both the
``utils`` attributes and the already-bound names inside ``trainer.eval_*`` are replaced.
``mAP_v2.compute_tp`` (utils/mAP.py:70-100) is rebound as well.  Loss modules keep the reference's torch IoUs unless
``patch_iou=True``: then gpu_iou / gpu_Giou / gpu_DIoU / gpu_CIoU are replaced too, in ``utils``, ``utils.bbox_tools`` and
every imported ``loss.*`` module that bound them (the mirrors are differentiable: ysb_*_iou_backward).
"""
import sys

from . import trainer as _trainer
from .utils import bbox_tools as _bbox
from .utils import mAP as _map
from .utils import nms as _nms

_EVALUATORS = {
    "eval_yolov5": "YOLOV5Evaluator", "eval_yolov7": "YOLOV7Evaluator", "eval_yolox": "YOLOXEvaluator",
    "eval_yolov8": "YOLOV8Evaluator", "eval_retinanet": "RetinaNetEvaluator",
    "eval_retinanet_experiment": "RetinaNetEvaluatorExperiment", "eval_fcos": "FCOSEvaluator",
}


def install(patch_evaluators=True, patch_utils=True, patch_iou=False):
    """Returns the list of names that were replaced."""
    done = []
    ref_utils, ref_trainer = sys.modules.get("utils"), sys.modules.get("trainer")
    if ref_utils is None or ref_trainer is None:
        raise RuntimeError("import the reference's `utils` and `trainer` packages before dropin.install()")
    if patch_utils:
        for name, fn in (("numba_nms", _nms.numba_nms), ("gpu_nms", _nms.gpu_nms), ("numba_iou", _bbox.numba_iou),
                         ("gpu_linear_soft_nms", _nms.gpu_linear_soft_nms),
                         ("gpu_exponential_soft_nms", _nms.gpu_exponential_soft_nms)):
            setattr(ref_utils, name, fn)
            for sub in ("utils.nms", "utils.bbox_tools"):
                mod = sys.modules.get(sub)
                if mod is not None and hasattr(mod, name):
                    setattr(mod, name, fn)
            for modname in _EVALUATORS:
                mod = sys.modules.get(f"trainer.{modname}")
                if mod is not None and hasattr(mod, name):
                    setattr(mod, name, fn)
            done.append(f"utils.{name}")
        ref_map = getattr(ref_utils, "mAP_v2", None)
        if ref_map is not None and hasattr(ref_map, "compute_tp"):
            ref_map.compute_tp = _map._compute_tp_method
            done.append("utils.mAP_v2.compute_tp")
    if patch_iou:
        for name in ("gpu_iou", "gpu_Giou", "gpu_DIoU", "gpu_CIoU"):
            fn = getattr(_bbox, name)
            holders = [ref_utils, sys.modules.get("utils.bbox_tools")]
            holders += [m for k, m in list(sys.modules.items()) if k == "loss" or k.startswith("loss.")]
            for mod in holders:
                if mod is not None and hasattr(mod, name):
                    setattr(mod, name, fn)
            done.append(f"utils.{name}")
    if patch_evaluators:
        for modname, cls in _EVALUATORS.items():
            new = getattr(_trainer, cls)
            setattr(ref_trainer, cls, new)
            mod = sys.modules.get(f"trainer.{modname}")
            if mod is not None:
                setattr(mod, cls, new)
            done.append(f"trainer.{cls}")
    return done
