"""One PostProcessor.run on synthetic heads (for compute-sanitizer / ncu): python profiles/tools/run_once.py family img batch dist [seed]"""
import sys

import torch

sys.path.insert(0, ".")
from yoloseries_b200 import synth  # noqa: E402
from yoloseries_b200.engine import PostProcessor  # noqa: E402

fam, img, batch, dist = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), sys.argv[4]
seed = int(sys.argv[5]) if len(sys.argv) > 5 else 2234
hyp = synth.map_profile_hyp(num_class=80)
anchors = torch.tensor(synth.V5_ANCHORS_PX) if fam in ("yolov5", "yolov7") else None
pp = PostProcessor(fam, hyp, anchors=anchors)
heads = synth.make_heads(fam, batch, img, img, 80, dist, seed=seed, device="cuda")
keys, counts = pp.filter_only(heads, img, img)
torch.cuda.synchronize()
print("survivors/img", counts[:, 0].tolist())
out = pp.run(heads, img, img)
torch.cuda.synchronize()
print("kept/img", out.det_cnt.tolist())
