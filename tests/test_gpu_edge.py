"""Edge cases of the CUDA path against the oracle: empty / ragged batches, ties, saturation, degenerate boxes, limits,
non-square inputs, and size-independent properties at the full BASELINE sizes."""
import numpy as np
import pytest
import torch

import oracle

pytestmark = pytest.mark.gpu


def _pp(family, hyp, anchors=True):
    from yoloseries_b200 import synth
    from yoloseries_b200.engine import PostProcessor
    a = torch.tensor(synth.V5_ANCHORS_PX) if (anchors and family in ("yolov5", "yolov7")) else None
    return PostProcessor(family, hyp, anchors=a)


def _check_against_oracle(family, heads, img_h, img_w, hyp):
    pp = _pp(family, hyp)
    decoded = pp.decode(heads, img_h, img_w).cpu().numpy()
    want = oracle.evaluator_nms(family, decoded, hyp)
    rows, idx = pp.to_list(pp.run(heads, img_h, img_w), as_numpy=True, with_index=True)
    assert len(rows) == len(want)
    for i, w in enumerate(want):
        if w.rows is None:
            assert rows[i] is None, f"image {i}: expected None"
            continue
        assert rows[i] is not None, f"image {i}: unexpected None"
        np.testing.assert_array_equal(rows[i], w.rows)
        np.testing.assert_array_equal(idx[i], w.cand_index)
    return rows, want


def test_ragged_batch_with_empty_images():
    """Images without any survivor give None; neighbours in the same batch are unaffected."""
    from yoloseries_b200 import synth
    hyp = oracle.default_hyp()
    heads = synth.make_heads("yolov5", 5, 320, 320, 80, "crowd", seed=3, device="cuda")
    for h in heads:                       # images 1 and 3: objectness far below conf_threshold everywhere
        h.view(5, 3, 85, h.shape[2], h.shape[3])[[1, 3], :, 4] = -30.0
    rows, want = _check_against_oracle("yolov5", heads, 320, 320, hyp)
    assert rows[1] is None and rows[3] is None and rows[0] is not None and len(rows[0]) > 0


def test_single_survivor_and_empty_after_count_filter():
    from yoloseries_b200 import synth
    hyp = oracle.default_hyp()
    heads = synth.make_heads("yolov5", 2, 128, 128, 80, "dense", seed=4, device="cuda")
    for h in heads:
        h.view(2, 3, 85, h.shape[2], h.shape[3])[:, :, 4] = -30.0
    heads[0].view(2, 3, 85, 16, 16)[0, 1, 4, 5, 7] = 4.0            # exactly one candidate of image 0 survives
    heads[1].view(2, 3, 85, 8, 8)[1, 0, 4, 2, 2] = 4.0              # image 1: two isolated survivors -> count filter
    heads[1].view(2, 3, 85, 8, 8)[1, 2, 4, 6, 6] = 4.0              #          leaves an empty (0, 6) array, not None
    rows, want = _check_against_oracle("yolov5", heads, 128, 128, hyp)
    assert rows[0].shape == (1, 6)                                    # M == 1: postprocess_bbox window not entered
    assert rows[1] is not None and rows[1].shape == (0, 6)


@pytest.mark.parametrize("value", [0.0, 50.0, -3.0])
def test_all_equal_logits_tie_breaking(value):
    """Every candidate has the same score: order is by candidate index, class ties resolve to class 0."""
    hyp = oracle.default_hyp(postprocess_bbox=False)
    heads = [torch.full((1, 255, s, s), value, device="cuda") for s in (16, 8, 4)]
    rows, want = _check_against_oracle("yolov5", heads, 128, 128, hyp)
    if rows[0] is not None and len(rows[0]):
        assert np.all(rows[0][:, 5] == 0.0)


def test_saturated_and_denormal_logits():
    from yoloseries_b200 import synth
    hyp = oracle.default_hyp()
    heads = synth.make_heads("yolov5", 2, 128, 128, 80, "dense", seed=6, device="cuda")
    for h in heads:
        h.mul_(30.0)                      # sigmoid saturates to exactly 1.0 / underflows to denormals and 0
    _check_against_oracle("yolov5", heads, 128, 128, hyp)


@pytest.mark.parametrize("family,h,w", [("yolov5", 384, 640), ("yolov7", 256, 448), ("yolox", 320, 192)])
def test_non_square_inputs(family, h, w):
    from yoloseries_b200 import synth
    hyp = oracle.default_hyp(num_class=12)
    heads = synth.make_heads(family, 2, h, w, 12, "dense", seed=8, device="cuda")
    pp = _pp(family, hyp)
    decoded = pp.decode(heads, h, w).cpu().numpy()
    hn = [t.cpu().numpy() for t in heads]
    ref = {"yolov5": lambda: oracle.decode_yolov5(hn, num_class=12), "yolov7": lambda: oracle.decode_yolov7(hn, num_class=12),
           "yolox": lambda: oracle.decode_yolox(hn, h, num_class=12)}[family]()
    assert np.all(np.abs(decoded - ref) <= 1e-5 * np.maximum(np.abs(ref), 1.0))
    _check_against_oracle(family, heads, h, w, hyp)


@pytest.mark.parametrize("max_det", [1, 7, 1024])
def test_max_det_limits(max_det):
    from yoloseries_b200 import synth
    hyp = oracle.default_hyp(max_predictions_per_img=max_det, postprocess_bbox=False)
    heads = synth.make_heads("yolov5", 1, 320, 320, 80, "dense", seed=9, device="cuda")
    rows, _ = _check_against_oracle("yolov5", heads, 320, 320, hyp)
    assert len(rows[0]) == max_det


def test_limits_raise():
    from yoloseries_b200 import synth
    heads = synth.make_heads("yolov5", 1, 64, 64, 80, "dense", seed=1, device="cuda")
    with pytest.raises(ValueError):
        _pp("yolov5", oracle.default_hyp(max_predictions_per_img=5000)).run(heads, 64, 64)
    from yoloseries_b200 import synth as _s
    with pytest.raises(NotImplementedError):  # the reference's multi-label branch is broken for RetinaNet
        _pp("retinanet", oracle.default_hyp(mutil_label=True)).run(_s.make_heads("retinanet", 1, 64, 64, 80, "dense", 1, "cuda"), 64, 64)
    with pytest.raises(ValueError):
        _pp("yolov5", oracle.default_hyp()).run([h.double() for h in heads], 64, 64)


def test_class_agnostic_mode_and_deploy_thresholds():
    from yoloseries_b200 import synth
    heads = synth.make_heads("yolov5", 2, 320, 320, 80, "crowd", seed=10, device="cuda")
    _check_against_oracle("yolov5", heads, 320, 320, oracle.default_hyp(agnostic=False))
    _check_against_oracle("yolov5", heads, 320, 320, oracle.default_hyp(conf_threshold=0.3, cls_threshold=0.3, iou_threshold=0.2))


def test_degenerate_and_duplicate_boxes_in_array_nms():
    from yoloseries_b200.utils import numba_nms
    rng = np.random.default_rng(5)
    base = rng.uniform(0, 50, size=(40, 2)).astype(np.float32)
    boxes = np.concatenate((base, base + rng.uniform(0, 20, size=(40, 2)).astype(np.float32)), axis=1)
    boxes[5] = [7, 7, 7, 7]                 # zero area: self-IoU is NaN, never suppresses, is kept if scored
    boxes[6] = [9, 3, 9, 12]                # zero width
    boxes[10] = boxes[11]                   # exact duplicates: IoU == 1 >= thr
    boxes[20] = [30, 30, 10, 10]            # inverted box: negative sides
    scores = rng.uniform(0.1, 1.0, size=40).astype(np.float32)
    scores[12] = scores[13]                 # equal scores: lower index first
    for thr in (0.3, 0.5, 1.0):
        assert numba_nms(boxes, scores, thr) == oracle.numba_nms(boxes, scores, thr)


@pytest.mark.parametrize("family,img,batch,dist", [("yolov5", 640, 8, "dense"), ("yolov5", 640, 8, "crowd"),
                                                   ("yolox", 640, 16, "dense"), ("retinanet", 640, 2, "dense")])
def test_full_size_properties(family, img, batch, dist):
    """Size-independent properties on BASELINE-sized inputs (no oracle needed): rows sorted by score, unique candidate
    indices, counts within max_det, kept boxes of one class never overlap at or above the threshold, and the run is
    deterministic."""
    from yoloseries_b200 import synth
    hyp = oracle.default_hyp(postprocess_bbox=False)
    heads = synth.make_heads(family, batch, img, img, 80, dist, seed=21, device="cuda")
    pp = _pp(family, hyp)
    out = pp.run(heads, img, img)
    rows, idx = pp.to_list(out, as_numpy=True, with_index=True)
    rows2, idx2 = pp.to_list(pp.run(heads, img, img), as_numpy=True, with_index=True)
    for r, i, r2, i2 in zip(rows, idx, rows2, idx2):
        assert r is not None and 0 < len(r) <= 300
        np.testing.assert_array_equal(r, r2)
        np.testing.assert_array_equal(i, i2)
        assert np.all(np.diff(r[:, 4]) <= 0)
        assert len(np.unique(i)) == len(i)
        off = (r[:, :4] + (r[:, 5] * np.float32(4096))[:, None]).astype(np.float32)
        iou = oracle.numba_iou(off, off)
        np.fill_diagonal(iou, 0.0)
        assert not np.any(iou >= hyp["iou_threshold"])


@pytest.mark.parametrize("family,img,dist", [("yolov5", 640, "dense"), ("yolov5", 640, "sparse"), ("yolov5", 640, "crowd"),
                                             ("yolov7", 320, "crowd"), ("yolox", 640, "dense"), ("yolox", 640, "sparse"),
                                             ("fcos", 512, "dense"), ("retinanet", 320, "dense")])
def test_nms_cta_flavours_agree_with_the_oracle(family, img, dist):
    """The NMS kernel has a 512-thread and a 1024-thread CTA flavour (picked by batch size, csrc/nms_kernel.cu
    launch_any) whose first tranche, sort width and post-filter path differ: both are forced here on the same inputs and
    each must reproduce the oracle's rows and candidate indices bit for bit."""
    from yoloseries_b200 import _lib, synth
    hyp = oracle.default_hyp()
    if family == "fcos":
        hyp.update(cls_threshold=0.2, iou_threshold=0.35, max_predictions_per_img=100)
    heads = synth.make_heads(family, 2, img, img, 80, dist, seed=77, device="cuda")
    lib = _lib.load()
    try:
        for threads in (512, 1024):
            assert lib.ysb_set_nms_cta_threads(threads) == 0
            _check_against_oracle(family, heads, img, img, hyp)
    finally:
        lib.ysb_set_nms_cta_threads(0)


def test_empty_batch_returns_empty_list():
    pp = _pp("yolov5", oracle.default_hyp())
    heads = [torch.zeros((0, 255, s, s), device="cuda") for s in (8, 4, 2)]
    assert pp.to_list(pp.run(heads, 64, 64)) == []


@pytest.mark.parametrize("family,img,dist", [("yolov5", 320, "crowd"), ("yolov7", 320, "crowd"), ("yolox", 320, "dense"),
                                             ("yolov8", 160, "dense"), ("fcos", 256, "dense")])
def test_multi_label_mode(family, img, dist):
    """mutil_label: one record per (candidate, class) -- rows and indices equal the oracle (itself pinned to the
    reference's multi-label goldens)."""
    from yoloseries_b200 import synth
    hyp = oracle.default_hyp(num_class=20, mutil_label=True, cls_threshold=0.25)
    if family == "fcos":
        hyp.update(iou_threshold=0.35, max_predictions_per_img=100)
    heads = synth.make_heads(family, 2, img, img, 20, dist, seed=31, device="cuda")
    _check_against_oracle(family, heads, img, img, hyp)


def test_pipelined_processor_matches_serial():
    from yoloseries_b200 import synth
    from yoloseries_b200.engine import PipelinedPostProcessor
    hyp = oracle.default_hyp()
    anchors = torch.tensor(synth.V5_ANCHORS_PX)
    ppp = PipelinedPostProcessor("yolov5", hyp, anchors=anchors, lanes=3)
    serial = _pp("yolov5", hyp)
    batches = [synth.make_heads("yolov5", 3, 320, 320, 80, d, seed=40 + i, device="cuda")
               for i, d in enumerate(["dense", "crowd", "sparse"])]
    tickets = [ppp.submit(h, 320, 320) for h in batches]
    for h, t in zip(batches, tickets):
        got = ppp.result(t, as_numpy=True)
        want = serial.to_list(serial.run(h, 320, 320), as_numpy=True)
        for g, w in zip(got, want):
            assert (g is None) == (w is None)
            if g is not None:
                np.testing.assert_array_equal(g, w)


@pytest.mark.parametrize("thr", [0.5, 0.25, 1.0 / 3.0, 0.2, 0.75])
def test_iou_exactly_at_threshold(thr):
    """Integer-coordinate boxes give exactly representable IoUs (1/2, 1/4, 1/3 rounded, ...): the '>=' of numba_nms
    versus the '>' of the count filter must both be reproduced on the boundary, with many tied scores."""
    from yoloseries_b200.utils import numba_nms
    rng = np.random.default_rng(int(thr * 1000))
    m = 600
    xy = rng.integers(0, 24, size=(m, 2)).astype(np.float32)
    wh = rng.integers(1, 9, size=(m, 2)).astype(np.float32)
    boxes = np.concatenate((xy, xy + wh), axis=1)
    scores = (rng.integers(1, 12, size=m) / 12.0).astype(np.float32)   # heavy ties
    ref = oracle.numba_nms(boxes, scores, thr)
    iou = oracle.numba_iou(boxes, boxes)
    assert np.sum(iou == thr) > 0 or thr in (1.0 / 3.0, 0.2)             # the boundary really is exercised
    assert numba_nms(boxes, scores, thr) == ref


def test_count_filter_exactly_at_threshold():
    """postprocess_bbox uses a strict '>' on the same IoU: a neighbour at exactly the threshold must not count."""
    hyp = oracle.default_hyp(num_class=1, iou_threshold=0.5, conf_threshold=0.0, cls_threshold=0.0)
    # decoded rows [cx, cy, w, h, obj, cls0] for yolov5 -> two boxes with IoU exactly 0.5, one isolated box
    dec = np.zeros((1, 3, 6), dtype=np.float32)
    dec[0, 0] = [1.0, 0.5, 2.0, 1.0, 0.9, 1.0]     # [0,0,2,1]
    dec[0, 1] = [0.5, 0.5, 1.0, 1.0, 0.8, 1.0]     # [0,0,1,1]  IoU with the first = 0.5
    dec[0, 2] = [10.0, 10.0, 2.0, 2.0, 0.7, 1.0]
    want = oracle.evaluator_nms("yolov5", dec, hyp)
    pp = _pp("yolov5", hyp)
    out = pp.run(torch.from_numpy(dec).cuda(), 640, 640, decoded=True)
    rows = pp.to_list(out, as_numpy=True)
    assert want[0].rows is not None and rows[0] is not None
    np.testing.assert_array_equal(rows[0], want[0].rows)


def test_equal_scores_beyond_one_tranche():
    """25 200 candidates with IDENTICAL scores: the radix select has to resolve the tranche on the candidate-index bits
    of the key alone (score range is zero), over several levels."""
    hyp = oracle.default_hyp(postprocess_bbox=False)
    heads = [torch.zeros((1, 255, s, s), device="cuda") for s in (80, 40, 20)]
    rows, want = _check_against_oracle("yolov5", heads, 640, 640, hyp)
    assert len(rows[0]) == 300


def test_two_score_levels_huge_ties():
    """Half of the candidates share one score, half another: every histogram digit but one is empty."""
    hyp = oracle.default_hyp(postprocess_bbox=False, max_predictions_per_img=1024)
    heads = [torch.zeros((2, 255, s, s), device="cuda") for s in (80, 40, 20)]
    for h in heads:
        v = h.view(2, 3, 85, h.shape[2], h.shape[3])
        v[:, :, 4, ::2, :] = 1.5          # objectness of every other row of cells
        v[:, :, 5 + 7] = 0.75             # one class stands out everywhere
    _check_against_oracle("yolov5", heads, 640, 640, hyp)


def test_cuda_graph_capture_replays_with_refilled_heads():
    from yoloseries_b200 import synth
    hyp = oracle.default_hyp()
    pp = _pp("yolov5", hyp)
    heads = synth.make_heads("yolov5", 2, 320, 320, 80, "crowd", seed=50, device="cuda")
    replay = pp.capture(heads, 320, 320)
    for seed, dist_ in ((51, "dense"), (52, "crowd")):
        fresh = synth.make_heads("yolov5", 2, 320, 320, 80, dist_, seed=seed, device="cuda")
        for dst, src in zip(heads, fresh):
            dst.copy_(src)                                   # refill in place: the graph holds the addresses
        got = pp.to_list(replay(), as_numpy=True)
        ref = _pp("yolov5", hyp)
        want = ref.to_list(ref.run(fresh, 320, 320), as_numpy=True)
        for g, w in zip(got, want):
            assert (g is None) == (w is None)
            if g is not None:
                np.testing.assert_array_equal(g, w)


def test_filter_kernel_variants_emit_the_same_survivors():
    """The profiling build (python -m yoloseries_b200.build --variants -> libysb_postproc_variants.so, loaded through
    YSB_LIBRARY) carries the tuning points of the direct-load kernels, the cp.async ring, the 1-D bulk-copy TMA ring and
    the 2-D tensor-map TMA ring, selected by environment variables read at library load: every one of them must emit the
    product kernel's survivor sets.  The product library itself has no such switches (the same variables are ignored)."""
    import os
    import subprocess
    import sys
    from yoloseries_b200 import build as ysb_build
    have_variants = ysb_build.variants_fresh()   # a stale profiling build (older ABI) is not used
    child = r'''
import hashlib, sys, torch
sys.path.insert(0, ".")
import oracle
from yoloseries_b200 import synth
from yoloseries_b200.engine import PostProcessor
for fam, img in (("yolov5", 320), ("yolox", 256), ("yolov8", 128)):
    pp = PostProcessor(fam, oracle.default_hyp(num_class=80),
                       anchors=torch.tensor(synth.V5_ANCHORS_PX) if fam == "yolov5" else None)
    heads = synth.make_heads(fam, 3, img, img, 80, "dense", seed=9, device="cuda")
    keys, counts = pp.filter_only(heads, img, img)
    torch.cuda.synchronize()
    h = hashlib.sha1(counts.cpu().numpy().tobytes())
    for i in range(3):
        h.update(torch.sort(keys[i, : int(counts[i, 0])]).values.cpu().numpy().tobytes())
    print(fam, int(counts[:, 0].sum()), h.hexdigest())
'''
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    outs = {}
    combos = (("1", "0"), ("1", "50"), ("1", "37"), ("1", "41"), ("3", "8"), ("3", "43"), ("2", "4"), ("0", "0"))
    for variant, ppt in combos if have_variants else (("1", "0"), ("3", "8"), ("0", "0")):
        env = dict(os.environ, YSB_FILTER_VARIANT=variant, YSB_BULK_PPT=ppt)
        if have_variants and (variant, ppt) != ("1", "0"):
            env["YSB_LIBRARY"] = ysb_build.VARIANTS_LIB
        r = subprocess.run([sys.executable, "-c", child], cwd=root, env=env, capture_output=True, text=True, timeout=300)
        assert r.returncode == 0, r.stderr[-1500:]
        outs[(variant, ppt)] = r.stdout.strip().splitlines()[-3:]
    base = outs[("1", "0")]
    assert len(base) == 3 and all(int(line.split()[1]) > 0 for line in base)
    for k, v in outs.items():
        assert v == base, (k, v, base)


def test_boxes_wider_than_the_class_offset():
    """The reference separates classes by adding cls * 4096 to the boxes (eval_yolov5.py:293-297), which only works while
    boxes stay inside [0, 4096]: wider or negative boxes of DIFFERENT classes do overlap in its arithmetic, in the NMS
    and in the postprocess_bbox count.  The class-bucketed count filter must notice and fall back to all pairs."""
    rng = np.random.default_rng(5)
    n, C = 600, 6
    hyp = oracle.default_hyp(num_class=C, conf_threshold=0.0, cls_threshold=0.0, iou_threshold=0.3)
    for widen in (False, True):
        dec = np.zeros((2, n, 5 + C), dtype=np.float32)
        dec[..., 0:2] = rng.uniform(100, 500, size=(2, n, 2))
        dec[..., 2:4] = rng.uniform(20, 120, size=(2, n, 2))
        if widen:  # a few boxes several class offsets wide and some with negative x1: cross-class overlaps exist
            dec[:, :40, 2] = rng.uniform(3000, 15000, size=(2, 40))
            dec[:, 40:60, 0] = rng.uniform(-300, 10, size=(2, 20))
        dec[..., 4] = rng.uniform(0.3, 1.0, size=(2, n))
        dec[..., 5:] = rng.uniform(0.0, 1.0, size=(2, n, C))
        want = oracle.evaluator_nms("yolov5", dec, hyp)
        pp = _pp("yolov5", hyp)
        rows, idx = pp.to_list(pp.run(torch.from_numpy(dec).cuda(), 640, 640, decoded=True), as_numpy=True, with_index=True)
        for i, w in enumerate(want):
            assert (w.rows is None) == (rows[i] is None)
            if w.rows is not None:
                np.testing.assert_array_equal(rows[i], w.rows)
                np.testing.assert_array_equal(idx[i], w.cand_index)


@pytest.mark.parametrize("threads", [0, 512, 1024])
def test_class_bucket_walk_switches_to_all_pairs_mid_image(threads):
    """The greedy walk tests a candidate only against kept / chunk boxes of its own class bucket (cls & 127) while every
    decoded box is finite and the boxes span <= 4095 px in x; the first wider box -- here they only appear in the LATER
    tranches, among the low scores -- switches the image to all pairs, where classes interact through the reference's
    cls * 4096 offset.  150 classes: several classes share a bucket.  Heavy crowding: the walk needs several tranches."""
    from yoloseries_b200 import _lib
    rng = np.random.default_rng(11)
    n, C, b = 2900, 150, 2                                             # < 3000: the postprocess_bbox count filter runs too
    hyp = oracle.default_hyp(num_class=C, conf_threshold=0.0, cls_threshold=0.0, iou_threshold=0.45)
    centres = rng.uniform(80, 560, size=(b, 40, 2))
    dec = np.zeros((b, n, 5 + C), dtype=np.float32)
    pick = rng.integers(0, 40, size=(b, n))
    dec[..., 0:2] = np.take_along_axis(centres, pick[..., None].repeat(2, -1), 1) + rng.normal(0, 4, size=(b, n, 2))
    dec[..., 2:4] = rng.uniform(40, 90, size=(b, n, 2))
    dec[..., 4] = rng.uniform(0.35, 1.0, size=(b, n))
    cls = (pick * 7 + rng.integers(0, 2, size=(b, n)) * 128) % C      # classes c and c + 128 land in one bucket
    dec[..., 5:] = rng.uniform(0.0, 0.2, size=(b, n, C))
    np.put_along_axis(dec[..., 5:], cls[..., None], rng.uniform(0.8, 1.0, size=(b, n, 1)).astype(np.float32), 2)
    low = np.argsort(dec[..., 4], axis=1)[:, :900]                    # the 900 lowest objectness values of each image
    for i in range(b):
        wide = low[i, rng.choice(900, size=60, replace=False)]
        dec[i, wide, 2] = rng.uniform(4500, 14000, size=60)
        dec[i, wide[:20], 0] = rng.uniform(-400, 0, size=20)
    want = oracle.evaluator_nms("yolov5", dec, hyp)
    lib = _lib.load()
    try:
        assert lib.ysb_set_nms_cta_threads(threads) == 0
        pp = _pp("yolov5", hyp)
        rows, idx = pp.to_list(pp.run(torch.from_numpy(dec).cuda(), 640, 640, decoded=True), as_numpy=True, with_index=True)
    finally:
        lib.ysb_set_nms_cta_threads(0)
    for i, w in enumerate(want):
        assert w.rows is not None and rows[i] is not None
        np.testing.assert_array_equal(rows[i], w.rows)
        np.testing.assert_array_equal(idx[i], w.cand_index)
        assert w.cand_index.max() >= 0


def test_sigmoid_reciprocal_is_frcp_rn():
    """sigmoid_ref's spelled-out reciprocal == __frcp_rn for every float in [1, +inf] (1.07e9 bit patterns), and the
    branch-free batch variant of the decode kernel agrees below 2^126 and flags everything at or above it."""
    import ctypes

    from yoloseries_b200 import _lib
    bad = torch.full((2,), -1, dtype=torch.int64, device="cuda")
    _lib.check(_lib.load().ysb_selftest_reciprocal(bad.data_ptr(), ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)),
               "ysb_selftest_reciprocal")
    assert bad.tolist() == [0, 0]
