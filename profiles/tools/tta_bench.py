"""use_tta path: fused ysb_postprocess_tta vs the staged form (3 x ysb_decode_into -> merged tensor -> decoded-rows
kernels) vs the reference-shaped torch expressions (decode, /= s, flips, torch.cat) feeding the decoded-rows kernels.
YOLOv5s 640x640, 80 classes, dense logits.  Usage: python profiles/tools/tta_bench.py [batch]"""
import sys

import torch

sys.path.insert(0, ".")
from yoloseries_b200 import synth  # noqa: E402
from yoloseries_b200.engine import PostProcessor  # noqa: E402

batch = int(sys.argv[1]) if len(sys.argv) > 1 else 16
hyp = synth.map_profile_hyp(num_class=80)
pp = PostProcessor("yolov5", hyp, anchors=torch.tensor(synth.V5_ANCHORS_PX))
S, Fl = (1, 0.83, 0.67), (None, 2, 3)
passes = [(synth.make_heads("yolov5", batch, 640, 640, 80, "dense", seed=10 + k, device="cuda"), 640, 640, s, f)
          for k, (s, f) in enumerate(zip(S, Fl))]


def fused():
    return pp.run_tta(passes, (640, 640))


def staged():
    merged, _ = pp.decode_tta(passes, (640, 640))
    return pp.run(merged, 640, 640, decoded=True)


def torch_style():
    out = []
    for heads, h, w, s, f in passes:
        p = pp.decode(heads, h, w)
        p[..., :4] /= s
        if f == 2:
            p[..., 1] = 640 - p[..., 1]
        if f == 3:
            p[..., 0] = 640 - p[..., 0]
        out.append(p)
    return pp.run(torch.cat(out, dim=1).contiguous(), 640, 640, decoded=True)


for name, fn in (("fused ysb_postprocess_tta", fused), ("decode_into x3 + decoded-rows kernels", staged),
                 ("decode + torch /=,flip,cat + decoded-rows kernels", torch_style)):
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(30):
        fn()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 30
    print(f"{name:52s} {ms:8.3f} ms / {batch} images  = {batch / ms * 1e3:10.0f} images/s")
