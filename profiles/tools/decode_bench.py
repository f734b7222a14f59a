"""Throughput of ysb_decode (the do_inference API path: raw heads -> (b, N, C') rows) per family."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from yoloseries_b200 import synth
from yoloseries_b200.engine import PostProcessor

for family, img, batch in (("yolov5", 640, 64), ("yolov7", 640, 64), ("yolox", 640, 128), ("yolov8", 640, 32),
                           ("retinanet", 640, 32), ("fcos", 640, 128)):
    heads = synth.make_heads(family, batch, img, img, 80, "dense", 1, "cuda")
    anchors = torch.tensor(synth.V5_ANCHORS_PX) if family in ("yolov5", "yolov7") else None
    pp = PostProcessor(family, synth.map_profile_hyp(), anchors=anchors)
    for _ in range(3):
        out = pp.decode(heads, img, img)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        out = pp.decode(heads, img, img)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 20
    flat = heads if isinstance(heads, (list,)) else [t for g in heads for t in (g if isinstance(g, (list, tuple)) else [g])]
    in_bytes = sum(t.numel() * 4 for t in flat)
    out_bytes = out.numel() * 4
    print(f"{family:10s} b={batch:3d} N={out.shape[1]:6d} {ms:.3f} ms  {batch / ms * 1e3:10.0f} img/s  "
          f"{(in_bytes + out_bytes) / ms / 1e6:7.0f} GB/s (read+write)")
