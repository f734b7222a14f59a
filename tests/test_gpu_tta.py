"""Test-time augmentation (hyp['use_tta'], the default of every reference yaml) through ysb_decode_into and the fused
ysb_postprocess_tta, against the reference's own merged tensors / rows (tests/golden/tta_*.npz) and the oracle."""
from collections import OrderedDict

import numpy as np
import pytest
import torch

import oracle
from conftest import close_rel, golden_names, hyp_from_meta, load_golden, tta_pass_heads

pytestmark = pytest.mark.gpu


def _to_dev(x):
    if isinstance(x, np.ndarray):
        return torch.from_numpy(np.ascontiguousarray(x)).cuda()
    return type(x)(_to_dev(v) for v in x)


def _processor(meta, **over):
    from yoloseries_b200.engine import PostProcessor
    from yoloseries_b200.synth import V5_ANCHORS_PX
    hyp = hyp_from_meta(meta)
    hyp.update(over)
    anchors = torch.tensor(V5_ANCHORS_PX) if meta["family"] in ("yolov5", "yolov7") else None
    return PostProcessor(meta["family"], hyp, anchors=anchors), hyp


def _assert_rows_equal(fam, got, ref):
    assert got.shape == ref.shape
    if fam.startswith("retinanet"):
        np.testing.assert_array_equal(got[:, 4:], ref[:, 4:])
        assert close_rel(got[:, :4], ref[:, :4], 1e-5).all()
    else:
        np.testing.assert_array_equal(got, ref)


def _passes(g, h, w):
    return [(_to_dev(tta_pass_heads(g, k)), h, w, s, f) for k, (s, f) in enumerate(zip(oracle.TTA_SCALES, oracle.TTA_FLIPS))]


@pytest.mark.parametrize("name", golden_names("tta_"))
def test_decode_into_builds_the_reference_merged_tensor(name):
    g = load_golden(name)
    meta = g["meta"]
    fam, C, h, w = meta["family"], meta["num_class"], meta["img_h"], meta["img_w"]
    pp, _ = _processor(meta)
    passes = _passes(g, h, w)
    merged, views = pp.decode_tta(passes, (h, w))
    got, ref = merged.cpu().numpy(), g["merged"]
    assert got.shape == ref.shape and len(views) == 3
    # (i) the undo + slotting is exact: equal to the oracle's undo of this library's own per-pass decode
    own = [pp.decode(p[0], h, w).cpu().numpy() for p in passes]
    np.testing.assert_array_equal(got, oracle.tta_merge(fam, own, h, w, C))
    off = 0
    for v, d in zip(views, own):
        assert v.shape == d.shape and v.data_ptr() == merged[:, off:].data_ptr()
        off += d.shape[1]
    # (ii) against the reference's merged tensor at the decode tolerance
    b0 = oracle.tta.box_col(fam, C)
    other = [c for c in range(ref.shape[2]) if not b0 <= c < b0 + 4]
    assert close_rel(got[..., other], ref[..., other], 1e-5).all()
    box_g, box_r = got[..., b0:b0 + 4], ref[..., b0:b0 + 4]
    if fam == "yolov8":
        assert np.abs(box_g - box_r).max() <= 1e-5 * (max(h, w) + 16 * 32) / min(oracle.TTA_SCALES)
    elif fam.startswith("retinanet"):
        bad = ~close_rel(box_g, box_r, 1e-5)
        assert bad.sum() <= 4 and np.all(np.abs(box_g[bad] - box_r[bad]) <= 1.0 / 0.67 + 1e-3)
    else:
        assert close_rel(box_g, box_r, 1e-5).all()


@pytest.mark.parametrize("name", golden_names("tta_"))
def test_fused_tta_vs_oracle_and_decoded_rows_path(name):
    """ysb_postprocess_tta (no merged tensor) == oracle on the merged tensor == the decoded-rows kernels on it,
    rows and merged candidate indices bit-exact."""
    g = load_golden(name)
    meta = g["meta"]
    fam, h, w = meta["family"], meta["img_h"], meta["img_w"]
    pp, hyp = _processor(meta)
    passes = _passes(g, h, w)
    merged, _ = pp.decode_tta(passes, (h, w))
    want = oracle.evaluator_nms(fam, merged.cpu().numpy(), hyp)
    rows, idx = pp.to_list(pp.run_tta(passes, (h, w)), as_numpy=True, with_index=True)
    staged, sidx = pp.to_list(pp.run(merged, h, w, decoded=True), as_numpy=True, with_index=True)
    for i, wnt in enumerate(want):
        if wnt.rows is None:
            assert rows[i] is None and staged[i] is None
            continue
        _assert_rows_equal(fam, rows[i], wnt.rows)
        np.testing.assert_array_equal(idx[i], wnt.cand_index)
        _assert_rows_equal(fam, staged[i], wnt.rows)
        np.testing.assert_array_equal(sidx[i], wnt.cand_index)


@pytest.mark.parametrize("name", golden_names("tta_"))
def test_reference_merged_tensor_gives_reference_rows(name):
    """The reference's merged tensor through the decoded-rows kernels == the reference's rows (bit-exact)."""
    g = load_golden(name)
    meta = g["meta"]
    pp, _ = _processor(meta)
    rows = pp.to_list(pp.run(_to_dev(g["merged"]), meta["img_h"], meta["img_w"], decoded=True), as_numpy=True)
    for i, r in enumerate(rows):
        cnt = int(g["counts"][i])
        if cnt < 0:
            assert r is None
        else:
            _assert_rows_equal(meta["family"], r, g["rows"][i, :cnt])


@pytest.mark.parametrize("name", ["tta_yolov5", "tta_yolov8", "tta_fcos"])
def test_evaluator_call_with_tta(name):
    """XEvaluator(use_tta=True).__call__ / test_time_augmentation with a stand-in model cycling through the pass heads."""
    from yoloseries_b200 import trainer
    from yoloseries_b200.synth import V5_ANCHORS_PX
    from test_gpu_api import _ref_hyp
    g = load_golden(name)
    meta = g["meta"]
    fam, h, w = meta["family"], meta["img_h"], meta["img_w"]
    sets = [_to_dev(tta_pass_heads(g, k)) for k in range(3)]
    calls = {"n": 0, "shapes": []}

    def model(x):
        calls["shapes"].append(tuple(x.shape[2:]))
        heads = sets[calls["n"] % 3]
        calls["n"] += 1
        return OrderedDict((f"p{i}", t) for i, t in enumerate(heads)) if fam == "yolov8" else heads

    hyp = _ref_hyp(meta, use_tta=True, input_img_size=[h, w])
    if fam == "yolov5":
        ev = trainer.YOLOV5Evaluator(model, torch.tensor(V5_ANCHORS_PX), hyp, compute_metric=True)
    elif fam == "yolov8":
        ev = trainer.YOLOV8Evaluator(model, hyp, True)
    else:
        ev = trainer.FCOSEvaluator(model, hyp, True)
    x = torch.rand(meta["batch"], 3, h, w, device="cuda")
    outs = ev(x)
    assert calls["n"] == 3 and calls["shapes"] == [(h, w)] * 3
    merged, views = ev.test_time_augmentation(x)
    assert merged.shape == g["merged"].shape and len(views) == 3
    staged = ev.numba_nms(merged)
    for o, s in zip(outs, staged):
        if s is None:
            assert o is None
            continue
        assert isinstance(o, torch.Tensor) and o.device.type == "cpu" and o.dtype == torch.float32
        np.testing.assert_array_equal(o.numpy(), s)
    # and against the reference's rows wherever the kept set agrees (scores differ by float32 sigmoid ulps at most)
    for i, o in enumerate(outs):
        cnt = int(g["counts"][i])
        if cnt >= 0 and o is not None and o.shape[0] == cnt:
            assert close_rel(o.numpy(), g["rows"][i, :cnt], 1e-5).all()


def test_passes_with_different_geometry():
    """Inputs that are not a multiple of 32 give passes of different sizes (scale_img pads to ceil(h/32)*32): every
    pass has its own level shapes, candidate count and stride table."""
    from yoloseries_b200 import synth
    from yoloseries_b200.engine import PostProcessor
    C = 6
    for fam in ("yolov5", "yolox", "fcos", "retinanet"):
        hyp = oracle.default_hyp(num_class=C)
        if fam == "fcos":
            hyp.update(cls_threshold=0.2, iou_threshold=0.35, max_predictions_per_img=100)
        anchors = torch.tensor(synth.V5_ANCHORS_PX) if fam == "yolov5" else None
        pp = PostProcessor(fam, hyp, anchors=anchors)
        sizes = [(128, 160), (160, 192), (128, 128)]
        passes = [(synth.make_heads(fam, 2, ph, pw, C, "dense", seed=900 + k, device="cuda"), ph, pw, s, f)
                  for k, ((ph, pw), s, f) in enumerate(zip(sizes, oracle.TTA_SCALES, oracle.TTA_FLIPS))]
        org = (150, 170)
        own = [pp.decode(p[0], p[1], p[2]).cpu().numpy() for p in passes]
        assert len({d.shape[1] for d in own}) == 3
        merged, _ = pp.decode_tta(passes, org)
        np.testing.assert_array_equal(merged.cpu().numpy(), oracle.tta_merge(fam, own, org[0], org[1], C))
        want = oracle.evaluator_nms(fam, merged.cpu().numpy(), hyp)
        rows, idx = pp.to_list(pp.run_tta(passes, org), as_numpy=True, with_index=True)
        for i, wnt in enumerate(want):
            if wnt.rows is None:
                assert rows[i] is None
                continue
            _assert_rows_equal(fam, rows[i], wnt.rows)
            np.testing.assert_array_equal(idx[i], wnt.cand_index)


@pytest.mark.parametrize("fam", ["retinanet", "retinanet_exp", "fcos"])
def test_tta_with_the_post_filter_window_active(fam):
    """Few enough merged survivors (1 < M < 3000; FCOS <= 300) that postprocess_bbox runs over boxes decoded from all three
    passes -- for RetinaNet including the score-weighted box merge (eval_retinanet.py:342-352)."""
    from yoloseries_b200 import synth
    from yoloseries_b200.engine import PostProcessor
    C = 5
    hyp = oracle.default_hyp(num_class=C)
    img = 64
    if fam == "fcos":
        hyp.update(cls_threshold=0.55, iou_threshold=0.35, max_predictions_per_img=100)
        img = 128
    pp = PostProcessor(fam, hyp)
    passes = [(synth.make_heads(fam, 3, img, img, C, "dense", seed=700 + k, device="cuda"), img, img, s, f)
              for k, (s, f) in enumerate(zip(oracle.TTA_SCALES, oracle.TTA_FLIPS))]
    merged, _ = pp.decode_tta(passes, (img, img))
    want = oracle.evaluator_nms(fam, merged.cpu().numpy(), hyp)
    rows, idx = pp.to_list(pp.run_tta(passes, (img, img)), as_numpy=True, with_index=True)
    checked = 0
    for i, wnt in enumerate(want):
        if wnt.rows is None:
            assert rows[i] is None
            continue
        _assert_rows_equal(fam, rows[i], wnt.rows)
        np.testing.assert_array_equal(idx[i], wnt.cand_index)
        checked += wnt.rows.shape[0]
    assert checked > 0


def test_full_size_tta_yolov5():
    """3 x 25 200 candidates per image (the reference's default validation path at 640x640)."""
    from yoloseries_b200 import synth
    from yoloseries_b200.engine import PostProcessor
    hyp = oracle.default_hyp(num_class=80)
    pp = PostProcessor("yolov5", hyp, anchors=torch.tensor(synth.V5_ANCHORS_PX))
    for dist in ("dense", "crowd"):
        passes = [(synth.make_heads("yolov5", 2, 640, 640, 80, dist, seed=300 + k, device="cuda"), 640, 640, s, f)
                  for k, (s, f) in enumerate(zip(oracle.TTA_SCALES, oracle.TTA_FLIPS))]
        merged, _ = pp.decode_tta(passes, (640, 640))
        assert merged.shape == (2, 75600, 85)
        want = oracle.evaluator_nms("yolov5", merged.cpu().numpy(), hyp)
        rows, idx = pp.to_list(pp.run_tta(passes, (640, 640)), as_numpy=True, with_index=True)
        for i, wnt in enumerate(want):
            _assert_rows_equal("yolov5", rows[i], wnt.rows)
            np.testing.assert_array_equal(idx[i], wnt.cand_index)
            assert idx[i].max() >= 25200  # rows from the augmented passes take part


def test_tta_argument_errors():
    from yoloseries_b200.engine import PostProcessor
    from yoloseries_b200 import synth
    pp = PostProcessor("yolox", oracle.default_hyp(num_class=4))
    heads = synth.make_heads("yolox", 1, 64, 64, 4, "dense", seed=1, device="cuda")
    with pytest.raises(ValueError):
        pp.run_tta([], (64, 64))
    with pytest.raises(ValueError):
        pp.run_tta([(heads, 64, 64, 1, None)] * 5, (64, 64))
    with pytest.raises(ValueError):
        pp.run_tta([(heads, 64, 64, 1, 1)], (64, 64))          # flip axis must be None / 2 / 3
    with pytest.raises(ValueError):
        two = synth.make_heads("yolox", 2, 64, 64, 4, "dense", seed=1, device="cuda")
        pp.run_tta([(heads, 64, 64, 1, None), (two, 64, 64, 0.83, 2)], (64, 64))
    one = pp.to_list(pp.run_tta([(heads, 64, 64, 1, None)], (64, 64)), as_numpy=True)
    ref = pp.to_list(pp.run(heads, 64, 64), as_numpy=True)
    for a, b in zip(one, ref):
        assert (a is None and b is None) or np.array_equal(a, b)
