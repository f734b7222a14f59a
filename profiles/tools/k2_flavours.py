"""Throughput of the select/sort/NMS kernel alone for a family and batch, with the CTA flavour forced (profiling build:
YSB_LIBRARY=.../libysb_k2timing.so).  Usage: python profiles/tools/k2_flavours.py family batch [dist]"""
import ctypes
import sys

import torch

sys.path.insert(0, ".")
from yoloseries_b200 import _lib, synth  # noqa: E402
from yoloseries_b200.engine import PostProcessor, flatten_heads  # noqa: E402

fam = sys.argv[1] if len(sys.argv) > 1 else "yolox"
batch = int(sys.argv[2]) if len(sys.argv) > 2 else 256
dist = sys.argv[3] if len(sys.argv) > 3 else "dense"
hyp = synth.map_profile_hyp(num_class=80)
if fam == "fcos":
    hyp.update(cls_threshold=0.2, iou_threshold=0.35, max_predictions_per_img=100)
anchors = torch.tensor(synth.V5_ANCHORS_PX) if fam in ("yolov5", "yolov7") else None
pp = PostProcessor(fam, hyp, anchors=anchors)
heads = synth.make_heads(fam, batch, 640, 640, 80, dist, seed=7, device="cuda")
flat = flatten_heads(fam, heads)
keys, counts = pp.filter_only(heads, 640, 640)
ent = pp._prepare(flat, batch, 640, 640, _lib.INPUT_RAW_HEADS)
lib = _lib.load()
raw = ctypes.CDLL(_lib.LIB_PATH)
out = ent["out"]
ptrs = _lib.head_pointer_array(flat)
st = torch.cuda.current_stream()


def nms():
    _lib.check(lib.ysb_select_nms(ctypes.byref(ent["params"]), ptrs, len(flat), keys.data_ptr(), keys.shape[1], counts.data_ptr(),
                                  out.dets.data_ptr(), out.det_idx.data_ptr(), out.det_cnt.data_ptr(),
                                  ctypes.c_void_p(st.cuda_stream)), "ysb_select_nms")


for threads in (1024, 512):
    if hasattr(raw, "ysb_debug_set_nms_threads"):
        raw.ysb_debug_set_nms_threads(threads)
    for _ in range(3):
        nms()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(20):
        nms()
    b.record()
    torch.cuda.synchronize()
    print(f"{fam} b={batch} {dist} threads={threads}: {a.elapsed_time(b) / 20 * 1000:.1f} us per launch, "
          f"kept/img {out.det_cnt.clamp(min=0).float().mean().item():.1f}, M/img {counts[:, 0].float().mean().item():.0f}")
