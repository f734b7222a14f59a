"""The reference-facing Python mirrors (yoloseries_b200.utils / .trainer) on the GPU, checked against the golden
vectors produced by the unmodified reference and against the oracle."""
from collections import OrderedDict

import numpy as np
import pytest
import torch

import oracle
from conftest import close_rel, golden_heads, golden_names, load_golden

pytestmark = pytest.mark.gpu


def _ref_hyp(meta, **over):
    hyp = dict(
        device="cuda", num_class=meta["num_class"], input_img_size=[meta["img"], meta["img"]], use_tta=False, wfb=False,
        iou_threshold=meta["iou_threshold"], conf_threshold=meta["conf_threshold"], cls_threshold=meta["cls_threshold"],
        compute_metric_iou_threshold=meta["compute_metric_iou_threshold"],
        compute_metric_conf_threshold=meta["compute_metric_conf_threshold"],
        compute_metric_cls_threshold=meta["compute_metric_cls_threshold"],
        max_predictions_per_img=meta["max_predictions_per_img"], min_prediction_box_wh=meta["min_prediction_box_wh"],
        iou_type="iou", mutil_label=meta["mutil_label"], agnostic=meta["agnostic"],
        postprocess_bbox=meta["postprocess_bbox"], num_anchors=1, reg=16, tar_box_scale_factor=[0.1, 0.1, 0.2, 0.2],
        pre_nms_topk=meta["pre_nms_topk"], pre_nms_thresh=meta["pre_nms_thresh"], thresh_with_ctr=meta["thresh_with_ctr"])
    hyp.update(over)
    return hyp


def _cuda(x):
    if isinstance(x, np.ndarray):
        return torch.from_numpy(np.ascontiguousarray(x)).cuda()
    return type(x)(_cuda(v) for v in x)


def _make_evaluator(g, **over):
    from yoloseries_b200 import trainer
    from yoloseries_b200.synth import V5_ANCHORS_PX
    meta = g["meta"]
    fam = meta["family"]
    hyp = _ref_hyp(meta, **over)
    heads = _cuda(golden_heads(g))
    anchors = torch.tensor(V5_ANCHORS_PX)
    if fam == "yolov5":
        return trainer.YOLOV5Evaluator(lambda x: heads, anchors, hyp, compute_metric=True)
    if fam == "yolov7":
        return trainer.YOLOV7Evaluator(lambda x: OrderedDict((f"p{i}", h) for i, h in enumerate(heads)), anchors, hyp, True)
    if fam == "yolox":
        return trainer.YOLOXEvaluator(lambda x: OrderedDict((f"p{i}", h) for i, h in enumerate(heads)), hyp, True)
    if fam == "yolov8":
        return trainer.YOLOV8Evaluator(lambda x: OrderedDict((f"p{i}", h) for i, h in enumerate(heads)), hyp, True)
    if fam == "retinanet":
        return trainer.RetinaNetEvaluator(lambda x: heads, hyp, True)
    if fam == "retinanet_exp":
        return trainer.RetinaNetEvaluatorExperiment(lambda x: heads, hyp, True)
    return trainer.FCOSEvaluator(lambda x: heads, hyp, True)


@pytest.mark.parametrize("name", golden_names())
def test_evaluator_numba_nms_method_matches_reference(name):
    """XEvaluator.numba_nms(reference decoded tensor) == the reference's own rows (bit-exact; RetinaNet boxes 1e-5)."""
    g = load_golden(name)
    ev = _make_evaluator(g)
    outs = ev.numba_nms(torch.from_numpy(g["decoded"]))
    assert isinstance(outs, list) and len(outs) == g["counts"].shape[0]
    for i, o in enumerate(outs):
        cnt = int(g["counts"][i])
        if cnt < 0:
            assert o is None
            continue
        assert isinstance(o, np.ndarray) and o.dtype == np.float32 and o.shape == (cnt, 6)
        ref = g["rows"][i, :cnt]
        if g["meta"]["family"].startswith("retinanet"):
            np.testing.assert_array_equal(o[:, 4:], ref[:, 4:])
            assert close_rel(o[:, :4], ref[:, :4], 1e-5).all()
        else:
            np.testing.assert_array_equal(o, ref)


@pytest.mark.parametrize("name", golden_names())
def test_evaluator_call_contract(name):
    """__call__ returns list[CPU float32 Tensor(K,6) | None] and agrees with numba_nms(do_inference(x))."""
    g = load_golden(name)
    ev = _make_evaluator(g)
    dummy = torch.zeros(g["meta"]["batch"], 3, g["meta"]["img"], g["meta"]["img"], device="cuda")
    outs = ev(dummy)
    staged = ev.numba_nms(ev.do_inference(dummy))
    assert len(outs) == len(staged)
    for o, s in zip(outs, staged):
        if s is None:
            assert o is None
            continue
        assert isinstance(o, torch.Tensor) and o.device.type == "cpu" and o.dtype == torch.float32
        if g["meta"]["family"].startswith("retinanet"):
            assert close_rel(o.numpy(), s, 1e-5).all()
        else:
            np.testing.assert_array_equal(o.numpy(), s)


def test_utils_numba_nms_and_iou_match_reference():
    from yoloseries_b200.utils import numba_iou, numba_nms
    g = load_golden("utils_nms_iou")
    for tag in ("a", "b"):
        boxes, scores = g[f"nms_{tag}_boxes"], g[f"nms_{tag}_scores"]
        keep_before = boxes.copy()
        for thr in (0.2, 0.5, 0.65):
            ref = g[f"nms_{tag}_keep_{thr}"].tolist()
            assert numba_nms(boxes, scores, thr) == ref            # full keep list, reference order
            assert numba_nms(boxes, scores, thr, max_keep=37) == ref[:37]
        np.testing.assert_array_equal(boxes, keep_before)            # inputs never mutated
        got = numba_iou(boxes[:64], boxes[:256])
        assert got.dtype == np.float64
        np.testing.assert_array_equal(got, g[f"iou_{tag}"])
    kat = numba_iou(g["kat_boxes"], g["kat_boxes"])
    np.testing.assert_array_equal(kat, g["kat_iou"])                 # NaN self-IoU of the degenerate box included
    assert numba_nms(g["kat_boxes"], np.array([.9, .8, .7, .6], np.float32), 0.5) == g["kat_keep_0.5"].tolist()
    disjoint = np.array([[0, 0, 1, 1], [10, 0, 11, 1], [20, 0, 21, 1], [30, 0, 31, 1]], np.float32)
    assert numba_nms(disjoint, np.array([.5, .9, .9, .1], np.float32), 0.5) == [1, 2, 0, 3]
    assert numba_nms(disjoint[:3], np.array([.5, 0, .9], np.float32), 0.5) == [2, 0]
    assert numba_nms(np.zeros((0, 4), np.float32), np.zeros(0, np.float32), 0.5) == []
    with pytest.raises(AssertionError):
        numba_nms(disjoint, np.ones(3, np.float32), 0.5)


def test_utils_large_unbounded_keep_list():
    """numba_nms returns the FULL keep list (no max_det) -- far beyond the 1024-entry shared-memory list of the fused path."""
    from yoloseries_b200.utils import numba_nms
    rng = np.random.default_rng(3)
    m = 6000
    xy = rng.uniform(0, 2000, size=(m, 2)).astype(np.float32)
    wh = rng.uniform(8, 80, size=(m, 2)).astype(np.float32)
    boxes = np.concatenate((xy, xy + wh), axis=1)
    scores = rng.uniform(0.01, 1, size=m).astype(np.float32)
    ref = oracle.numba_nms(boxes, scores, 0.5)
    assert len(ref) > 2000
    assert numba_nms(boxes, scores, 0.5) == ref


def test_torch_iou_family_matches_reference():
    from yoloseries_b200.utils import gpu_CIoU, gpu_DIoU, gpu_Giou, gpu_iou, gpu_nms
    g = load_golden("utils_nms_iou")
    b1, b2 = torch.from_numpy(g["tiou_b1"]).cuda(), torch.from_numpy(g["tiou_b2"]).cuda()
    assert np.abs(gpu_iou(b1[:50], b2).cpu().numpy() - g["tiou_iou"]).max() <= 1e-6
    for fn, key in ((gpu_Giou, "tiou_giou"), (gpu_DIoU, "tiou_diou"), (gpu_CIoU, "tiou_ciou")):
        got = fn(b1, b2).cpu().numpy()
        assert got.shape == g[key].shape
        assert np.abs(got - g[key]).max() <= 2e-6, key
    assert np.abs(gpu_Giou(b1[:1], b2).cpu().numpy() - g["tiou_giou_row"]).max() <= 2e-6
    assert np.abs(gpu_DIoU(b1[:1], b2).cpu().numpy() - g["tiou_diou_row"]).max() <= 2e-6
    sc = torch.from_numpy(g["tnms_scores"]).cuda()
    for kind in ("giou", "diou"):
        assert gpu_nms(b2, sc, kind, 0.45) == g[f"tnms_{kind}"].tolist()
    # 'iou' raises IndexError in the reference as shipped; defined by intent == oracle
    assert gpu_nms(b2, sc, "iou", 0.45) == oracle.gpu_nms_iou(g["tiou_b2"], g["tnms_scores"], 0.45)
    with pytest.raises(ValueError):
        gpu_nms(b2, sc, "siou", 0.45)
    with pytest.raises(AssertionError):
        gpu_nms(g["tiou_b2"], sc, "iou", 0.45)


def test_rowwise_iou_autograd_matches_reference():
    """gpu_Giou / gpu_DIoU / gpu_CIoU under autograd (ysb_elementwise_iou_backward) vs torch autograd through the
    reference's own functions: |d - ref| <= 1e-5 * max(|ref|, 1) (float32 chain; the golden gradients are O(0.1))."""
    from yoloseries_b200.utils import gpu_CIoU, gpu_DIoU, gpu_Giou, gpu_iou
    g = load_golden("utils_nms_iou")
    w = torch.from_numpy(g["grad_w"]).cuda()
    worst = 0.0
    for kind, fn in (("giou", gpu_Giou), ("diou", gpu_DIoU), ("ciou", gpu_CIoU)):
        x1 = torch.from_numpy(g["grad_b1"]).cuda().requires_grad_(True)
        x2 = torch.from_numpy(g["grad_b2"]).cuda().requires_grad_(True)
        out = fn(x1, x2)
        assert out.requires_grad and out.shape == (64,)
        (out * w).sum().backward()
        for got, key in ((x1.grad, f"grad_{kind}_d1"), (x2.grad, f"grad_{kind}_d2")):
            err = np.abs(got.cpu().numpy() - g[key])
            worst = max(worst, float(err.max()))
            assert close_rel(got.cpu().numpy(), g[key], 1e-5).all(), (key, err.max())
    for kind, fn in (("giou", gpu_Giou), ("diou", gpu_DIoU)):      # broadcast box1: gradient summed over the rows
        x1 = torch.from_numpy(g["grad_b1"][:1]).cuda().requires_grad_(True)
        x2 = torch.from_numpy(g["grad_b2"]).cuda().requires_grad_(True)
        (fn(x1, x2) * w).sum().backward()
        assert x1.grad.shape == (1, 4)
        assert close_rel(x1.grad.cpu().numpy(), g[f"grad_{kind}_row_d1"], 1e-5).all()
        assert close_rel(x2.grad.cpu().numpy(), g[f"grad_{kind}_row_d2"], 1e-5).all()
    assert worst <= 2e-6, worst                                       # in practice far inside the stated tolerance
    # only one side needs a gradient (the loss case: targets are constants); CPU leaves get CPU gradients
    x1 = torch.from_numpy(g["grad_b1"]).requires_grad_(True)
    x2 = torch.from_numpy(g["grad_b2"])
    gpu_CIoU(x1, x2).sum().backward()
    assert x1.grad.device.type == "cpu" and x2.grad is None
    ref1, _ = oracle.iou_backward("ciou", g["grad_b1"], g["grad_b2"], np.ones(64))
    assert close_rel(x1.grad.numpy(), ref1, 1e-5).all()
    # a larger random batch against the analytic oracle
    rng = np.random.default_rng(3)
    xy = rng.uniform(0, 600, size=(5000, 2)).astype(np.float32)
    wh = rng.uniform(2, 200, size=(5000, 2)).astype(np.float32)
    a = np.concatenate((xy, xy + wh), axis=1)
    xy2 = xy + rng.normal(0, 20, size=(5000, 2)).astype(np.float32)
    wh2 = (wh * np.exp(rng.normal(0, 0.4, size=(5000, 2)))).astype(np.float32)
    b = np.concatenate((xy2, xy2 + wh2), axis=1).astype(np.float32)
    go = rng.uniform(-1, 1, size=5000).astype(np.float32)
    for kind, fn in (("giou", gpu_Giou), ("diou", gpu_DIoU), ("ciou", gpu_CIoU)):
        x1, x2 = torch.from_numpy(a).cuda().requires_grad_(True), torch.from_numpy(b).cuda().requires_grad_(True)
        fn(x1, x2).backward(torch.from_numpy(go).cuda())
        r1, r2 = oracle.iou_backward(kind, a, b, go)
        assert close_rel(x1.grad.cpu().numpy(), r1, 1e-5).all() and close_rel(x2.grad.cpu().numpy(), r2, 1e-5).all()
    # no_grad / detached inputs take the plain forward
    with torch.no_grad():
        assert not gpu_CIoU(x1, x2).requires_grad
        assert not gpu_iou(x1, x2).requires_grad


def test_pairwise_iou_autograd_matches_reference():
    """utils.gpu_iou under autograd (ysb_pairwise_iou_backward) vs torch autograd through the reference's own gpu_iou
    (tests/golden/utils_extra.npz; the call shape of loss/yolox_loss.py:133: targets (N,4) x predictions (M,4) that
    require grad): |d - ref| <= 1e-5 * max(|ref|, 1)."""
    from yoloseries_b200.utils import gpu_iou
    g = load_golden("utils_extra")
    w = torch.from_numpy(g["pair_w"]).cuda()
    x1 = torch.from_numpy(g["pair_b1"]).cuda().requires_grad_(True)
    x2 = torch.from_numpy(g["pair_b2"]).cuda().requires_grad_(True)
    out = gpu_iou(x1, x2)
    assert out.requires_grad and out.shape == (37, 150)
    assert close_rel(out.detach().cpu().numpy(), g["pair_iou"], 1e-5).all()
    (out * w).sum().backward()
    for got, key in ((x1.grad, "pair_d1"), (x2.grad, "pair_d2")):
        err = np.abs(got.cpu().numpy() - g[key]).max()
        assert close_rel(got.cpu().numpy(), g[key], 1e-5).all(), (key, err)
        assert err <= 2e-6, (key, err)
    # the loss case: only the predictions need a gradient, targets are constants; CPU leaves get CPU gradients
    p = torch.from_numpy(g["pair_b2"]).requires_grad_(True)
    gpu_iou(torch.from_numpy(g["pair_b1"]), p).sum().backward()
    assert p.grad.device.type == "cpu"
    _, r2 = oracle.pairwise_iou_backward(g["pair_b1"], g["pair_b2"], np.ones((37, 150)))
    assert close_rel(p.grad.numpy(), r2, 1e-5).all()
    # larger, non-multiple-of-tile shapes against the analytic oracle; empty sides give zero gradients
    rng = np.random.default_rng(5)
    for n, m in ((1, 1), (3, 1000), (700, 45), (257, 300)):
        xy = rng.uniform(0, 300, size=(m, 2)).astype(np.float32)
        b = np.concatenate((xy, xy + rng.uniform(2, 120, size=(m, 2)).astype(np.float32)), axis=1)
        xy = rng.uniform(0, 300, size=(n, 2)).astype(np.float32)
        a = np.concatenate((xy, xy + rng.uniform(2, 120, size=(n, 2)).astype(np.float32)), axis=1)
        go = rng.uniform(-1, 1, size=(n, m)).astype(np.float32)
        x1, x2 = torch.from_numpy(a).cuda().requires_grad_(True), torch.from_numpy(b).cuda().requires_grad_(True)
        gpu_iou(x1, x2).backward(torch.from_numpy(go).cuda())
        r1, r2 = oracle.pairwise_iou_backward(a, b, go)
        assert close_rel(x1.grad.cpu().numpy(), r1, 1e-5).all() and close_rel(x2.grad.cpu().numpy(), r2, 1e-5).all(), (n, m)
    x1 = torch.zeros((0, 4), device="cuda", requires_grad=True)
    x2 = torch.from_numpy(b).cuda().requires_grad_(True)
    gpu_iou(x1, x2).sum().backward()
    assert x1.grad.shape == (0, 4) and torch.count_nonzero(x2.grad).item() == 0


def test_soft_nms_matches_reference():
    """utils/nms.py:68-140 through ysb_soft_nms: the reference's own keep masks, and the oracle on a larger set."""
    from yoloseries_b200.utils import gpu_exponential_soft_nms, gpu_linear_soft_nms
    g = load_golden("utils_nms_iou")
    sb, ss = torch.from_numpy(g["soft_boxes"]).cuda(), torch.from_numpy(g["soft_scores"]).cuda()
    for kind in ("giou", "diou", "ciou"):
        got = gpu_linear_soft_nms(sb, ss, kind, iou_threshold=0.1, thresh=0.4)
        assert got.dtype == torch.bool and got.shape == (48,) and got.device == ss.device
        np.testing.assert_array_equal(got.cpu().numpy(), g[f"soft_linear_{kind}"], err_msg=kind)
    got = gpu_exponential_soft_nms(sb[:12], ss[:12], "diou", 0.3, sigmma=0.5, thresh=0.001)
    np.testing.assert_array_equal(got.cpu().numpy(), g["soft_exp_diou"])
    # CPU tensors in -> CPU mask out, inputs untouched (the reference clones)
    keep = ss.cpu().clone()
    got = gpu_linear_soft_nms(sb.cpu(), keep, "giou", iou_threshold=0.1, thresh=0.4)
    assert got.device.type == "cpu" and torch.equal(keep, ss.cpu())
    # larger random set against the oracle restatement (more than one 1024-thread sweep per pick)
    rng = np.random.default_rng(11)
    m = 2500
    xy = rng.uniform(0, 400, size=(m, 2)).astype(np.float32)
    wh = rng.uniform(8, 90, size=(m, 2)).astype(np.float32)
    boxes = np.concatenate((xy, xy + wh), axis=1)
    scores = rng.uniform(0.01, 1, size=(m, 1)).astype(np.float32)
    for kind, thr in (("giou", 0.2), ("diou", 0.3)):
        ref = oracle.linear_soft_nms(boxes, scores, kind, iou_threshold=thr, thresh=0.3)
        got = gpu_linear_soft_nms(torch.from_numpy(boxes).cuda(), torch.from_numpy(scores).cuda(), kind, thr, 0.3)
        assert 0 < ref.sum() < m
        np.testing.assert_array_equal(got.cpu().numpy(), ref, err_msg=kind)
    assert gpu_linear_soft_nms(sb[:0], ss[:0], "giou").shape == (0,)
    with pytest.raises(ValueError):
        gpu_linear_soft_nms(sb, ss, "siou")
    with pytest.raises(AssertionError):
        gpu_linear_soft_nms(sb, ss[:5], "giou")


def test_undo_letterbox_bit_exact():
    """val_yolov5.py:166-172 through ysb_undo_letterbox: list-level mirror and the in-place device hook."""
    from yoloseries_b200.engine import DetectionBuffers, PostProcessor, preds_postprocess
    g = load_golden("utils_nms_iou")
    scale, pad_top, pad_left, org_h, org_w = g["lb_info"].tolist()
    info = dict(scale=scale, pad_top=int(pad_top), pad_left=int(pad_left), pad_bottom=0, pad_right=0,
                org_shape=(int(org_h), int(org_w)))
    other = dict(scale=0.5, pad_top=12, pad_left=0, pad_bottom=12, pad_right=0, org_shape=(720, 1280))
    rows = torch.from_numpy(g["lb_in"])
    out = preds_postprocess([rows, None, rows[:7], rows[:0]], [info, info, other, info])
    np.testing.assert_array_equal(out[0], g["lb_out"])
    assert out[1] is None and out[3].shape == (0, 6)
    np.testing.assert_array_equal(out[2], oracle.undo_letterbox(g["lb_in"][:7], 0.5, 12, 0, 720, 1280))
    np.testing.assert_array_equal(rows.numpy(), g["lb_in"])  # caller's rows untouched
    # device hook on DetectionBuffers: rows beyond the count are left alone
    buf = DetectionBuffers(2, 64, torch.device("cuda"))
    buf.dets.zero_()
    buf.dets[0, :40] = rows.cuda()
    buf.dets[1, :40] = rows.cuda()
    buf.det_cnt.copy_(torch.tensor([40, 5], dtype=torch.int32))
    pp = PostProcessor("yolov5", oracle.default_hyp())
    pp.undo_letterbox(buf, [info, info])
    torch.cuda.synchronize()
    np.testing.assert_array_equal(buf.dets[0, :40].cpu().numpy(), g["lb_out"])
    np.testing.assert_array_equal(buf.dets[1, :5].cpu().numpy(), g["lb_out"][:5])
    np.testing.assert_array_equal(buf.dets[1, 5:40].cpu().numpy(), g["lb_in"][5:])
    with pytest.raises(ValueError):
        pp.undo_letterbox(buf, [info])


def test_fused_letterbox_undo_equals_separate_pass():
    """ysb_params.d_letterbox: the undo of val_yolov5.py:166-172 inside the NMS kernel's row write == NMS rows mapped
    afterwards by the oracle's literal restatement (bit-exact), per image, incl. TTA; counts and indices unchanged; and
    the evaluator mirror's DetectionList lets preds_postprocess reuse the device rows."""
    from yoloseries_b200 import synth
    from yoloseries_b200.engine import DetectionList, PostProcessor, preds_postprocess
    from yoloseries_b200.trainer import YOLOV5Evaluator
    hyp = oracle.default_hyp(num_class=6, postprocess_bbox=False)
    anchors = torch.tensor(synth.V5_ANCHORS_PX)
    heads = synth.make_heads("yolov5", 3, 128, 160, 6, "crowd", seed=9, device="cuda")
    infos = [dict(scale=0.731, pad_top=3, pad_left=0, org_shape=(167, 219)),
             dict(scale=0.25, pad_top=0, pad_left=21, org_shape=(512, 470)),
             dict(scale=1.0, pad_top=0, pad_left=0, org_shape=(128, 160))]
    pp = PostProcessor("yolov5", hyp, anchors=anchors)
    plain, plain_idx = pp.to_list(pp.run(heads, 128, 160), as_numpy=True, with_index=True)
    fused, fused_idx = pp.to_list(pp.run(heads, 128, 160, info=infos), as_numpy=True, with_index=True)
    again = pp.to_list(pp.run(heads, 128, 160), as_numpy=True)      # info=None switches it off again
    assert sum(r.shape[0] for r in plain) > 30
    for i, d in enumerate(infos):
        want = oracle.undo_letterbox(plain[i], d["scale"], d["pad_top"], d["pad_left"], *d["org_shape"])
        np.testing.assert_array_equal(fused[i], want)
        np.testing.assert_array_equal(fused_idx[i], plain_idx[i])
        np.testing.assert_array_equal(again[i], plain[i])
    with pytest.raises(ValueError):
        pp.run(heads, 128, 160, info=infos[:2])
    # evaluator mirror: __call__(inputs, info=...) fused; __call__(inputs) + preds_postprocess(outputs, info) on device rows
    hyp_e = dict(hyp, device="cuda", use_tta=False, input_img_size=[128, 160], wfb=False)
    ev = YOLOV5Evaluator(lambda x: [h.clone() for h in heads], anchors, hyp_e)
    x = torch.zeros(3, 3, 128, 160, device="cuda")
    outs = ev(x)
    assert isinstance(outs, DetectionList) and outs.device_rows is not None
    mapped = preds_postprocess(outs, infos)
    direct = ev(x, info=infos)
    for i in range(3):
        np.testing.assert_array_equal(outs[i].numpy(), plain[i])
        np.testing.assert_array_equal(mapped[i], fused[i])
        np.testing.assert_array_equal(direct[i].numpy(), fused[i])
    # TTA: the undo applies to the merged result
    hyp_t = dict(hyp_e, use_tta=True)
    ev_t = YOLOV5Evaluator(lambda x: synth.make_heads("yolov5", 3, x.shape[2], x.shape[3], 6, "crowd", seed=9, device="cuda"),
                           anchors, hyp_t)
    base = ev_t(x)
    got = ev_t(x, info=infos)
    for i, d in enumerate(infos):
        want = oracle.undo_letterbox(base[i].numpy(), d["scale"], d["pad_top"], d["pad_left"], *d["org_shape"])
        np.testing.assert_array_equal(got[i].numpy(), want)


def test_compute_tp_matches_reference():
    """utils/mAP.py:18-42, 70-100 through ysb_map_iou / ysb_compute_tp: the reference's own tp matrices and IoUs
    (utils_extra.npz, float32 and float64 images), one launch for the whole list, and random images vs the oracle."""
    from yoloseries_b200.utils import compute_tp, compute_tp_batch
    from yoloseries_b200.utils.mAP import iou as map_iou
    g = load_golden("utils_extra")
    gts, preds = [g[f"tp_gt_{i}"] for i in range(6)], [g[f"tp_pred_{i}"] for i in range(6)]
    for i in range(6):
        got = compute_tp(gts[i], preds[i])
        assert got.dtype == bool and got.shape == (preds[i].shape[0], 10)
        np.testing.assert_array_equal(got, g[f"tp_out_{i}"], err_msg=f"image {i}")
        io = map_iou(gts[i][:, :4], preds[i][:, :4])
        assert io.dtype == g[f"tp_iou_{i}"].dtype and io.shape == g[f"tp_iou_{i}"].shape
        np.testing.assert_array_equal(io, g[f"tp_iou_{i}"])
    # images of one precision batched into one launch
    for sel in ([0, 1, 2, 4], [3, 5]):
        outs = compute_tp_batch([gts[i] for i in sel], [preds[i] for i in sel])
        for i, o in zip(sel, outs):
            np.testing.assert_array_equal(o, g[f"tp_out_{i}"])
    # mixed precisions promote to float64 like numpy: compare with the oracle on the promoted arrays
    outs = compute_tp_batch(gts, preds)
    for i, o in enumerate(outs):
        np.testing.assert_array_equal(o, oracle.compute_tp(gts[i].astype(np.float64), preds[i].astype(np.float64)))
    # a validation-sized run: 200 images x up to 300 kept rows
    rng = np.random.default_rng(21)
    big_g, big_p = [], []
    for _ in range(200):
        n = int(rng.integers(0, 40))
        xy = rng.uniform(0, 600, size=(n, 2))
        gt = np.concatenate((xy, xy + rng.uniform(10, 200, size=(n, 2)), rng.integers(0, 5, size=(n, 1))), axis=1).astype(np.float32)
        m = int(rng.integers(0, 300))
        src = gt[rng.integers(0, n, size=m)] if n else np.zeros((m, 5), np.float32)
        box = src[:, :4] + rng.normal(0, 8, size=(m, 4)).astype(np.float32)
        lab = np.where(rng.random(m) < 0.9, src[:, 4], rng.integers(0, 5, size=m)).astype(np.float32)
        big_g.append(gt)
        big_p.append(np.concatenate((box, rng.uniform(0, 1, size=(m, 1)).astype(np.float32), lab[:, None]), axis=1).astype(np.float32))
    outs = compute_tp_batch(big_g, big_p)
    total = 0
    for gt, pr, o in zip(big_g, big_p, outs):
        want = oracle.compute_tp(gt, pr)
        np.testing.assert_array_equal(o, want)
        total += int(want.sum())
    assert total > 5000
    assert compute_tp_batch([], []) == []
    with pytest.raises(ValueError):
        compute_tp_batch(gts, preds[:2])


def test_gather_detections_single_process_identity():
    from yoloseries_b200.dist import gather_detections
    d = torch.rand(4, 10, 6, device="cuda")
    c = torch.tensor([1, -1, 0, 10], dtype=torch.int32, device="cuda")
    gd, gc = gather_detections(d, c, 4)
    assert gd is d and gc is c
