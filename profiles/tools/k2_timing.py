import ctypes, sys, torch, numpy as np
sys.path.insert(0, __import__("os").path.dirname(__import__("os").path.dirname(__import__("os").path.dirname(__import__("os").path.abspath(__file__)))))
from yoloseries_b200 import synth, _lib
from yoloseries_b200.engine import PostProcessor
lib = _lib.load()
fam = sys.argv[1] if len(sys.argv) > 1 else "yolov5"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 64
dists = ("dense", "sparse", "crowd") if fam == "yolov5" else ("dense", "sparse")
for dist in dists:
    heads = synth.make_heads(fam, B, 640, 640, 80, dist, 1234, "cuda")
    hyp = synth.map_profile_hyp()
    if fam == "fcos":
        hyp.update(cls_threshold=0.2, iou_threshold=0.35, max_predictions_per_img=100)
    pp = PostProcessor(fam, hyp, anchors=torch.tensor(synth.V5_ANCHORS_PX) if fam in ("yolov5", "yolov7") else None)
    for _ in range(3):
        pp.run(heads, 640, 640)
    torch.cuda.synchronize()
    buf = (ctypes.c_longlong * (64 * 16))()
    f = ctypes.CDLL(_lib.LIB_PATH).ysb_debug_k2_timing
    f(buf)
    t = np.array(buf[:]).reshape(64, 16).astype(np.float64)
    d = np.diff(t[:, :7], axis=1) / 1.9  # ns at ~1.9 GHz
    names = ["select", "gather", "sort", "(setup)", "decode+nms", "postfilter", "emit"]
    names = ["select(hist)", "gather", "sort", "nms(+decode)", "postfilter", "emit"]
    acc = t[:, 8:13].mean(axis=0) / 1.9 / 1000
    print(fam, B, dist, "nms breakdown: decode=%.1f phaseA=%.1f phaseB=%.1f phaseC=%.1f append=%.1f us" % tuple(acc))
    print(dist, " ".join(f"{n}={v/1000:.1f}us" for n, v in zip(names, d.mean(axis=0))), "total=%.1fus" % ((t[:, 6] - t[:, 0]).mean() / 1.9 / 1000))
