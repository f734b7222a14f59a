"""dropin.install(): the names the reference binds (utils/__init__.py:1-21, trainer/__init__.py:1-8, and the
``from utils import numba_nms`` style imports inside trainer/eval_*.py) are replaced in every module that holds them."""
import sys
import types

import pytest


def _stub_reference(monkeypatch):
    def old(*a, **k):
        raise AssertionError("reference implementation still bound")

    utils = types.ModuleType("utils")
    utils_nms = types.ModuleType("utils.nms")
    utils_bbox = types.ModuleType("utils.bbox_tools")
    for name in ("numba_nms", "gpu_nms", "gpu_linear_soft_nms", "gpu_exponential_soft_nms"):
        setattr(utils, name, old)
        setattr(utils_nms, name, old)
    utils.numba_iou = utils_bbox.numba_iou = old
    trainer = types.ModuleType("trainer")
    eval_v5 = types.ModuleType("trainer.eval_yolov5")
    eval_v5.numba_nms = eval_v5.numba_iou = old          # `from utils import numba_nms, numba_iou`
    eval_v5.YOLOV5Evaluator = trainer.YOLOV5Evaluator = type("YOLOV5Evaluator", (), {})
    for name, mod in (("utils", utils), ("utils.nms", utils_nms), ("utils.bbox_tools", utils_bbox), ("trainer", trainer),
                      ("trainer.eval_yolov5", eval_v5)):
        monkeypatch.setitem(sys.modules, name, mod)
    return utils, utils_nms, utils_bbox, trainer, eval_v5, old


def test_install_rebinds_every_holder(monkeypatch):
    from yoloseries_b200 import dropin, trainer as ours, utils as our_utils
    utils, utils_nms, utils_bbox, trainer, eval_v5, old = _stub_reference(monkeypatch)
    done = dropin.install()
    for name in ("numba_nms", "gpu_nms", "gpu_linear_soft_nms", "gpu_exponential_soft_nms"):
        assert getattr(utils, name) is getattr(our_utils, name)
        assert getattr(utils_nms, name) is getattr(our_utils, name)
        assert f"utils.{name}" in done
    assert utils.numba_iou is our_utils.numba_iou and utils_bbox.numba_iou is our_utils.numba_iou
    assert eval_v5.numba_nms is our_utils.numba_nms and eval_v5.numba_iou is our_utils.numba_iou
    assert trainer.YOLOV5Evaluator is ours.YOLOV5Evaluator and eval_v5.YOLOV5Evaluator is ours.YOLOV5Evaluator
    assert "trainer.YOLOV5Evaluator" in done


def test_install_selective(monkeypatch):
    from yoloseries_b200 import dropin
    utils, _, _, trainer, eval_v5, old = _stub_reference(monkeypatch)
    ref_cls = trainer.YOLOV5Evaluator
    done = dropin.install(patch_evaluators=False)
    assert trainer.YOLOV5Evaluator is ref_cls and utils.numba_nms is not old
    assert not [d for d in done if d.startswith("trainer.")]


def test_install_needs_reference_imported(monkeypatch):
    from yoloseries_b200 import dropin
    monkeypatch.delitem(sys.modules, "utils", raising=False)
    monkeypatch.delitem(sys.modules, "trainer", raising=False)
    with pytest.raises(RuntimeError):
        dropin.install()
