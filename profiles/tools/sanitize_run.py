"""Every C-ABI entry point once at small sizes, for compute-sanitizer (memcheck / racecheck / initcheck):
    compute-sanitizer --tool memcheck python profiles/tools/sanitize_run.py"""
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
from yoloseries_b200 import synth  # noqa: E402
from yoloseries_b200.engine import PostProcessor, preds_postprocess  # noqa: E402
from yoloseries_b200.utils import (gpu_CIoU, gpu_DIoU, gpu_Giou, gpu_exponential_soft_nms, gpu_iou,  # noqa: E402
                                   gpu_linear_soft_nms, gpu_nms, numba_iou, numba_nms)

C = 8
for fam, img in (("yolov5", 128), ("yolov7", 128), ("yolox", 96), ("yolov8", 64), ("retinanet", 64), ("retinanet_exp", 64),
                 ("fcos", 160)):
    hyp = synth.map_profile_hyp(num_class=C)
    if fam == "fcos":
        hyp.update(cls_threshold=0.2, iou_threshold=0.35, max_predictions_per_img=100)
    anchors = torch.tensor(synth.V5_ANCHORS_PX) if fam in ("yolov5", "yolov7") else None
    for multi in (False, True):
        if multi and fam.startswith("retinanet"):
            continue
        pp = PostProcessor(fam, dict(hyp, mutil_label=multi, cls_threshold=0.3 if multi else hyp["cls_threshold"]), anchors=anchors)
        for dist in ("dense", "sparse"):
            heads = synth.make_heads(fam, 3, img, img, C, dist, seed=3, device="cuda")
            dec = pp.decode(heads, img, img)
            pp.to_list(pp.run(heads, img, img))
            pp.to_list(pp.run(dec, img, img, decoded=True))
        passes = [(synth.make_heads(fam, 2, img, img, C, "dense", seed=5 + k, device="cuda"), img, img, s, f)
                  for k, (s, f) in enumerate(zip((1, 0.83, 0.67), (None, 2, 3)))]
        out = pp.run_tta(passes, (img, img))
        pp.undo_letterbox(out, [dict(scale=0.5, pad_top=3, pad_left=0, org_shape=(200, 240))] * 2)
        pp.to_list(out)
        pp.decode_tta(passes, (img, img))
    print(fam, "ok", flush=True)

# both CTA flavours of the NMS kernel on multi-tranche inputs: a crowd (several tranches, class-bucket walk) and decoded rows
# whose wide boxes switch the walk to all pairs mid-image (count filter active)
from yoloseries_b200 import _lib  # noqa: E402
lib = _lib.load()
rng = np.random.default_rng(11)
n, CC = 2900, 150
dec = np.zeros((2, n, 5 + CC), dtype=np.float32)
dec[..., 0:2] = rng.uniform(80, 560, size=(2, 40, 2))[:, rng.integers(0, 40, size=n)] + rng.normal(0, 4, size=(2, n, 2))
dec[..., 2:4] = rng.uniform(40, 90, size=(2, n, 2))
dec[..., 4] = rng.uniform(0.35, 1.0, size=(2, n))
dec[..., 5:] = rng.uniform(0.0, 1.0, size=(2, n, CC))
dec[:, -60:, 2] = rng.uniform(4500, 14000, size=(2, 60))
dec[:, -60:, 4] *= 0.3
for threads in (512, 1024):
    lib.ysb_set_nms_cta_threads(threads)
    hyp = synth.map_profile_hyp(num_class=80)
    pp = PostProcessor("yolov5", hyp, anchors=torch.tensor(synth.V5_ANCHORS_PX))
    for dist in ("crowd", "dense", "sparse"):
        heads = synth.make_heads("yolov5", 2, 320, 320, 80, dist, seed=9, device="cuda")
        pp.to_list(pp.run(heads, 320, 320))
    hyp2 = synth.map_profile_hyp(num_class=CC)
    hyp2.update(conf_threshold=0.0, cls_threshold=0.0, iou_threshold=0.45)
    pp2 = PostProcessor("yolov5", hyp2, anchors=torch.tensor(synth.V5_ANCHORS_PX))
    pp2.to_list(pp2.run(torch.from_numpy(dec).cuda(), 640, 640, decoded=True))
    print("nms flavour", threads, "ok", flush=True)
lib.ysb_set_nms_cta_threads(0)

rng = np.random.default_rng(0)
xy = rng.uniform(0, 300, size=(3000, 2)).astype(np.float32)
wh = rng.uniform(4, 80, size=(3000, 2)).astype(np.float32)
boxes = np.concatenate((xy, xy + wh), axis=1)
scores = rng.uniform(0, 1, size=3000).astype(np.float32)
numba_nms(boxes, scores, 0.5)
numba_nms(boxes[:0], scores[:0], 0.5)
numba_iou(boxes[:70], boxes[:300])
tb, ts = torch.from_numpy(boxes).cuda(), torch.from_numpy(scores).cuda()
for kind in ("iou", "giou", "diou", "ciou"):
    gpu_nms(tb, ts, kind, 0.45)
gpu_iou(tb[:50], tb[:333])
gpu_linear_soft_nms(tb[:700], ts[:700, None], "giou", 0.3, 0.2)
gpu_exponential_soft_nms(tb[:20], ts[:20, None], "diou", 0.3)
for fn in (gpu_Giou, gpu_DIoU, gpu_CIoU):
    a, b = tb[:1000].clone().requires_grad_(True), tb[1000:2000].clone().requires_grad_(True)
    fn(a, b).sum().backward()
a = tb[:1].clone().requires_grad_(True)
gpu_Giou(a, tb[:777]).sum().backward()
preds_postprocess([tb[:5].new_zeros(5, 6).cpu(), None], [dict(scale=0.5, pad_top=3, pad_left=0, org_shape=(200, 240))] * 2)
torch.cuda.synchronize()
print("all entry points ok")
