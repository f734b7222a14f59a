"""Time the filter kernel variants alone (CUDA events, 64 YOLOv5s images, rotating inputs > L2) and check that every
variant emits the same survivor set.  Usage: python profiles/tools/filter_variants.py "1:0" "3:8" "3:4" ...
(each argument = YSB_FILTER_VARIANT:YSB_BULK_PPT; the library reads them at load time, so one subprocess each)."""
import hashlib
import os
import subprocess
import sys

CHILD = r'''
import hashlib, sys, torch
sys.path.insert(0, ".")
from yoloseries_b200 import synth
from yoloseries_b200.engine import PostProcessor
fam, batch = sys.argv[1], int(sys.argv[2])
hyp = synth.map_profile_hyp(num_class=80)
anchors = torch.tensor(synth.V5_ANCHORS_PX) if fam in ("yolov5", "yolov7") else None
pp = PostProcessor(fam, hyp, anchors=anchors)
sets = [synth.make_heads(fam, batch, 640, 640, 80, "dense", seed=40 + k, device="cuda") for k in range(3)]
keys, counts = pp.filter_only(sets[0], 640, 640)
torch.cuda.synchronize()
m = counts[:, 0].cpu()
h = hashlib.sha1()
for i in range(batch):
    h.update(torch.sort(keys[i, : int(m[i])]).values.cpu().numpy().tobytes())
h.update(counts.cpu().numpy().tobytes())
for _ in range(5):
    for s in sets:
        pp.filter_only(s, 640, 640)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
n = 60
e0.record()
for i in range(n):
    pp.filter_only(sets[i % 3], 640, 640)
e1.record()
torch.cuda.synchronize()
print(f"{e0.elapsed_time(e1) / n:.4f} ms  M={int(m.sum())}  sha1={h.hexdigest()[:12]}")
'''

fam = os.environ.get("FAMILY", "yolov5")
batch = os.environ.get("BATCH", "64")
for spec in sys.argv[1:]:
    v, ppt = spec.split(":")
    env = dict(os.environ, YSB_FILTER_VARIANT=v, YSB_BULK_PPT=ppt)
    r = subprocess.run([sys.executable, "-c", CHILD, fam, batch], env=env, capture_output=True, text=True, timeout=300)
    out = r.stdout.strip().splitlines()[-1] if r.stdout.strip() else ("FAILED: " + r.stderr.strip()[-400:])
    print(f"variant {v} ppt {ppt:>3s}: {out}", flush=True)
