"""A few filter-kernel launches for profiling under ncu (variant chosen by YSB_FILTER_VARIANT / YSB_BULK_PPT).
Usage: python profiles/tools/filter_once.py [family] [batch]"""
import sys

import torch

sys.path.insert(0, ".")
from yoloseries_b200 import synth  # noqa: E402
from yoloseries_b200.engine import PostProcessor  # noqa: E402

fam = sys.argv[1] if len(sys.argv) > 1 else "yolov5"
batch = int(sys.argv[2]) if len(sys.argv) > 2 else 64
hyp = synth.map_profile_hyp(num_class=80)
anchors = torch.tensor(synth.V5_ANCHORS_PX) if fam in ("yolov5", "yolov7") else None
pp = PostProcessor(fam, hyp, anchors=anchors)
sets = [synth.make_heads(fam, batch, 640, 640, 80, "dense", seed=40 + k, device="cuda") for k in range(3)]
for i in range(9):
    pp.filter_only(sets[i % 3], 640, 640)
torch.cuda.synchronize()
