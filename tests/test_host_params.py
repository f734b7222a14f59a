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


# ---- full shape validation of the head tensors (the C ABI only sees bare pointers) -----------------------------------
def _heads(family, b=2, img=64, C=5):
    return engine.flatten_heads(family, synth.make_heads(family, b, img, img, C, "dense", 3, "cpu"))


@pytest.mark.parametrize("family", ["yolov5", "yolov7", "yolox", "yolov8", "retinanet", "retinanet_exp", "fcos"])
def test_validate_heads_accepts_the_reference_layouts(family):
    anchors = _anchors() if family in ("yolov5", "yolov7") else None
    engine.validate_heads(family, _heads(family), 5, anchors, 16)


def test_validate_heads_rejects_wrong_geometry():
    a = _anchors()
    good = _heads("yolov5")
    with pytest.raises(ValueError, match="level 0 must be"):     # channel count is not A*(5+C)
        engine.validate_heads("yolov5", good, 6, a)
    with pytest.raises(ValueError, match="anchor groups"):       # fewer levels than anchor groups
        engine.validate_heads("yolov5", good[:2], 5, a)
    with pytest.raises(ValueError, match="batch"):               # batch differs across levels
        engine.validate_heads("yolov5", [good[0], good[1][:1], good[2]], 5, a)
    with pytest.raises(ValueError, match="contiguous float32"):
        engine.validate_heads("yolov5", [good[0].double(), good[1], good[2]], 5, a)
    with pytest.raises(ValueError, match="contiguous float32"):
        engine.validate_heads("yolov5", [good[0].transpose(2, 3), good[1], good[2]], 5, a)
    reg, cls = _heads("retinanet")
    with pytest.raises(ValueError, match="expected reg"):        # reg / cls row counts differ
        engine.validate_heads("retinanet", [reg[:, :-9].contiguous(), cls], 5)
    with pytest.raises(ValueError, match="expected reg"):        # the 5-column reg tensor belongs to retinanet_exp
        engine.validate_heads("retinanet_exp", [reg, cls], 5)
    f = _heads("fcos", img=128)
    with pytest.raises(ValueError, match="level 1 must be"):
        bad = list(f)
        bad[5 + 1] = bad[5 + 1][:, :3].contiguous()
        engine.validate_heads("fcos", bad, 5)
    with pytest.raises(ValueError, match=r"\(b, N, 10\)"):
        engine.validate_heads("yolov5", [torch.zeros(2, 7, 9)], 5, decoded_row_w=10)
    with pytest.raises(ValueError, match="level 2 must be"):
        v8 = _heads("yolov8")
        engine.validate_heads("yolov8", v8[:2] + [v8[2][:, :-1].contiguous()] + v8[3:], 5, None, 16)


def test_fcos_stride_follows_hyp_input_size_like_the_reference():
    """trainer/eval_fcos.py:137 -- level stride = hyp['input_img_size'][0] / fm_h, not actual input height / fm_h."""
    hyp = oracle.default_hyp(num_class=3)
    hyp["input_img_size"] = [256, 256]
    shapes = [(16, 16), (8, 8), (4, 4), (2, 2), (1, 1)]
    p = engine.make_params("fcos", hyp, 1, 128, 128, shapes)
    assert [p.level_stride[i] for i in range(5)] == [16.0, 32.0, 64.0, 128.0, 256.0]
    del hyp["input_img_size"]
    p = engine.make_params("fcos", hyp, 1, 128, 128, shapes)
    assert [p.level_stride[i] for i in range(5)] == [8.0, 16.0, 32.0, 64.0, 128.0]
