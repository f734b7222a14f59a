import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from yoloseries_b200 import synth
from yoloseries_b200.engine import PostProcessor
family = sys.argv[1] if len(sys.argv) > 1 else "yolov5"
heads = synth.make_heads(family, 32, 640, 640, 80, "dense", 1, "cuda")
pp = PostProcessor(family, synth.map_profile_hyp(), anchors=torch.tensor(synth.V5_ANCHORS_PX) if family in ("yolov5", "yolov7") else None)
for _ in range(4):
    pp.decode(heads, 640, 640)
torch.cuda.synchronize()
