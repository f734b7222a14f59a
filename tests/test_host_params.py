"""Host-side translation of the reference's hyp dict and head shapes into ysb_params (engine.make_params) -- no GPU."""
import numpy as np
import pytest
import torch

import oracle
from yoloseries_b200 import _lib, engine, synth


def _anchors():
    return torch.tensor(synth.V5_ANCHORS_PX)


@pytest.mark.parametrize("family,strides,apc", [
    ("yolov5", (8, 16, 32), 3), ("yolov7", (8, 16, 32), 3), ("yolox", (8, 16, 32), 1), ("yolov8", (4, 8, 16, 32), 1),
    ("fcos", (8, 16, 32, 64, 128), 1)])
def test_grid_families(family, strides, apc):
    hyp = oracle.default_hyp(num_class=7)
    shapes = [(640 // s, 640 // s) for s in strides]
    p = engine.make_params(family, hyp, 5, 640, 640, shapes, _anchors() if family in ("yolov5", "yolov7") else None)
    assert (p.family, p.batch, p.num_classes, p.num_levels) == (_lib.FAMILY_IDS[family], 5, 7, len(strides))
    assert [p.level_h[i] for i in range(p.num_levels)] == [h for h, _ in shapes]
    assert [p.level_stride[i] for i in range(p.num_levels)] == [float(s) for s in strides]
    assert p.anchors_per_cell == apc
    assert p.input_kind == _lib.INPUT_RAW_HEADS and p.decoded_rows == 0
    assert (p.tta_scale, p.tta_flip) == (0.0, 0)


def test_thresholds_follow_the_compute_metric_profile():
    hyp = oracle.default_hyp(num_class=3)
    hyp.update(compute_metric_iou_threshold=0.5, compute_metric_conf_threshold=0.25, compute_metric_cls_threshold=0.3)
    shapes = [(8, 8), (4, 4), (2, 2)]
    a = engine.make_params("yolov5", hyp, 1, 64, 64, shapes, _anchors(), compute_metric=False)
    b = engine.make_params("yolov5", hyp, 1, 64, 64, shapes, _anchors(), compute_metric=True)
    assert (a.iou_thr, b.iou_thr) == (hyp["iou_threshold"], 0.5)
    assert np.float32(b.conf_thr) == np.float32(0.25) and np.float32(b.cls_thr) == np.float32(0.3)
    # thresholds are float32: numpy compares float32 arrays with Python floats in float32 (numpy >= 2)
    assert np.float32(a.cls_thr) == np.float32(hyp["cls_threshold"])
    assert a.class_aware == int(bool(hyp["agnostic"])) and a.max_det == hyp["max_predictions_per_img"]


def test_anchor_scaling_matches_the_reference_expression():
    """(self.anchors[i] / self.ds_scales[i]).type_as(inputs), eval_yolov5.py:192: true divide, then float32."""
    p = engine.make_params("yolov5", oracle.default_hyp(), 1, 640, 640, [(80, 80), (40, 40), (20, 20)], _anchors())
    anc = _anchors()
    for i, s in enumerate((8, 16, 32)):
        want = (anc[i] / s).to(torch.float32)
        for a in range(3):
            assert p.anchor[i][a][0] == float(want[a, 0]) and p.anchor[i][a][1] == float(want[a, 1])


def test_retinanet_levels_and_base_anchors():
    p = engine.make_params("retinanet", oracle.default_hyp(), 2, 640, 512)
    assert p.num_levels == 5 and p.anchors_per_cell == 9
    assert [p.level_h[i] for i in range(5)] == [80, 40, 20, 10, 5] and [p.level_w[i] for i in range(5)] == [64, 32, 16, 8, 4]
    base = oracle.retinanet_base_anchors(32)  # level 3: size 2^(3+2)
    got = np.array([[p.anchor[0][a][c] for c in range(4)] for a in range(9)], dtype=np.float32)
    np.testing.assert_array_equal(got, base.astype(np.float32))
    assert [p.reg_scale[i] for i in range(4)] == [np.float32(x) for x in (0.1, 0.1, 0.2, 0.2)]
    # odd sizes: ceil division like the reference's pyramid ((x - 1) // 2**l + 1)
    q = engine.make_params("retinanet", oracle.default_hyp(), 1, 100, 100)
    assert [q.level_h[i] for i in range(5)] == [13, 7, 4, 2, 1]


def test_tta_and_decoded_rows_fields():
    hyp = oracle.default_hyp()
    p = engine.make_params("yolox", hyp, 1, 64, 96, [(8, 12), (4, 6), (2, 3)], tta=(0.83, 2, 60, 90))
    assert np.float32(p.tta_scale) == np.float32(0.83) and (p.tta_flip, p.tta_img_h, p.tta_img_w) == (2, 60, 90)
    p = engine.make_params("yolox", hyp, 1, 64, 96, [(8, 12), (4, 6), (2, 3)], tta=(1, None, 64, 96))
    assert (p.tta_scale, p.tta_flip) == (1.0, 0)
    d = engine.make_params("yolov5", hyp, 1, 64, 64, [(8, 8), (4, 4), (2, 2)], _anchors(),
                           input_kind=_lib.INPUT_DECODED_ROWS, decoded_rows=756)
    assert d.input_kind == _lib.INPUT_DECODED_ROWS and d.decoded_rows == 756


def test_letterbox_table_and_argument_errors():
    info = [dict(scale=0.5, pad_top=3, pad_left=7, org_shape=(200, 240)), dict(scale=1.25, pad_top=0, pad_left=0, org_shape=[64, 48])]
    assert engine.letterbox_table(info) == [[0.5, 3.0, 7.0, 200.0, 240.0], [1.25, 0.0, 0.0, 64.0, 48.0]]
    with pytest.raises(ValueError):
        engine.make_params("yolov5", oracle.default_hyp(), 1, 64, 64, [(8, 8)], None)       # anchors required
    with pytest.raises(ValueError):
        engine.make_params("yolox", oracle.default_hyp(), 1, 64, 64, None)                  # level shapes required
    with pytest.raises(ValueError):
        engine.preds_postprocess([None], [])


def test_normalise_heads_is_a_no_op_for_reference_outputs_and_fixes_the_rest():
    from collections import OrderedDict
    a = torch.zeros(2, 255, 8, 8)
    assert engine.normalise_heads(a) is a                                    # float32 + contiguous: untouched, no copy
    lst = [a, torch.zeros(2, 255, 4, 4)]
    out = engine.normalise_heads(lst)
    assert isinstance(out, list) and out[0] is a and out[1] is lst[1]
    half = torch.zeros(2, 3, 8, 8, 85, dtype=torch.float16)
    view = torch.zeros(2, 85, 3, 8, 8).permute(0, 2, 3, 4, 1)                # channels-last view, not contiguous
    d = engine.normalise_heads(OrderedDict(p0=half, p1=view))
    assert isinstance(d, OrderedDict) and list(d) == ["p0", "p1"]
    assert d["p0"].dtype == torch.float32 and d["p1"].is_contiguous() and d["p1"].shape == view.shape
    fcos = ([a], [torch.zeros(2, 4, 8, 8)], [torch.zeros(2, 1, 8, 8).half()])
    f = engine.normalise_heads(fcos)
    assert isinstance(f, tuple) and f[0][0] is a and f[2][0].dtype == torch.float32
