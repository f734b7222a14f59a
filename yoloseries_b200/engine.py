"""Host-side engine: turns (family, hyp, head tensors) into ysb_params + device buffers and calls the C ABI.

PyTorch is plumbing here (device memory, current stream, pinned host buffers); all compute happens in
libysb_postproc.so.  Nothing in this module computes on the CPU and nothing falls back to torch ops.
"""
import ctypes
from collections import OrderedDict

import numpy as np
import torch

from . import _lib, _ops
from ._lib import YsbParams

# strides fixed by the reference evaluators (self.ds_scales), trainer/eval_yolov5.py:21, eval_yolov7.py, eval_fcos.py:18
_FIXED_STRIDES = {"yolov5": (8, 16, 32), "yolov7": (8, 16, 32), "fcos": (8, 16, 32, 64, 128)}


def retinanet_base_anchors(size):
    """The 9 base anchors of one pyramid level, float32, ratio-major (utils/anchor.py:176-191).

    Every step is a single correctly-rounded float32 operation (mul, div, sqrt, sub), so this numpy restatement is
    bit-identical to the torch code of the reference; the device kernel adds the per-cell shift analytically.
    """
    f = np.float32
    scales = np.array([1, 2 ** (1 / 3), 2 ** (2 / 3)], dtype=f)
    ratios = np.array([0.5, 1, 2], dtype=f)
    side = f(size) * np.tile(scales, 3)
    areas = side * side
    rr = np.repeat(ratios, 3)
    w = np.sqrt(areas / rr)
    h = w * rr
    out = np.zeros((9, 4), dtype=f)
    out[:, 0] = f(0) - w / f(2)
    out[:, 1] = f(0) - h / f(2)
    out[:, 2] = w - w / f(2)
    out[:, 3] = h - h / f(2)
    return out


def flatten_heads(family, heads):
    """Head tensors in the order the C ABI expects (include/ysb_postproc.h, enum ysb_family)."""
    if family in ("retinanet", "retinanet_exp"):
        reg, cls = heads
        return [reg, cls]
    if family == "fcos":
        cls_l, reg_l, ctr_l = heads
        return list(cls_l) + list(reg_l) + list(ctr_l)
    if isinstance(heads, (dict, OrderedDict)):
        return list(heads.values())
    return list(heads)


def normalise_heads(heads):
    """Model outputs as the engine needs them: float32 and contiguous, same nesting.  A no-op (no copy) for what the
    reference's models emit by default; half-precision or permuted-view outputs are converted like the reference's own
    ``.float()`` / ``.contiguous()`` calls (trainer/eval_yolov5.py:194, 265)."""
    if isinstance(heads, torch.Tensor):
        if heads.dtype == torch.float32 and heads.is_contiguous():
            return heads
        return heads.detach().to(torch.float32).contiguous()
    if isinstance(heads, (dict, OrderedDict)):
        return type(heads)((k, normalise_heads(v)) for k, v in heads.items())
    return type(heads)(normalise_heads(v) for v in heads)


def _level_shapes(family, flat, num_class):
    if family in ("yolov5", "yolox", "yolov8"):
        return [(t.shape[-2], t.shape[-1]) for t in flat]
    if family == "yolov7":
        return [(t.shape[2], t.shape[3]) for t in flat]
    if family == "fcos":
        n = len(flat) // 3
        return [(t.shape[2], t.shape[3]) for t in flat[:n]]
    raise ValueError(family)


def make_params(family, hyp, batch, img_h, img_w, level_shapes=None, anchors=None, input_kind=_lib.INPUT_RAW_HEADS,
                compute_metric=False, num_anchors=None, decoded_rows=0, tta=None):
    """Build ``ysb_params`` from the reference's flat ``hyp`` dict (SURVEY.md section 5 lists the keys).

    ``tta`` = (scale, flip_axis, org_h, org_w) describes one test-time-augmentation pass (trainer/eval_yolov5.py:152-179):
    the boxes decoded from this pass's heads are divided by ``scale`` and un-flipped (axis 2 or 3) against the size of
    the un-augmented input."""
    p = YsbParams()
    p.family = _lib.FAMILY_IDS[family]
    p.input_kind = input_kind
    p.batch = int(batch)
    p.num_classes = int(hyp["num_class"])
    p.img_h, p.img_w = int(img_h), int(img_w)
    pre = "compute_metric_" if compute_metric else ""
    p.iou_thr = float(hyp[pre + "iou_threshold"])
    p.cls_thr = float(hyp[pre + "cls_threshold"])
    p.conf_thr = float(hyp.get(pre + "conf_threshold", 0.0))
    p.pre_nms_thr = float(hyp.get("pre_nms_thresh", 0.0))
    p.max_det = int(hyp["max_predictions_per_img"])
    p.class_aware = int(bool(hyp["agnostic"]))
    p.multi_label = int(bool(hyp["mutil_label"]))
    p.postprocess_bbox = int(bool(hyp["postprocess_bbox"]))
    p.min_box_wh = float(hyp.get("min_prediction_box_wh", 0))
    p.pre_nms_topk = int(hyp.get("pre_nms_topk", 1000))
    p.thresh_with_ctr = int(bool(hyp.get("thresh_with_ctr", True)))
    p.dfl_bins = int(hyp.get("reg", 16))
    p.decoded_rows = int(decoded_rows)
    if tta is not None:
        p.tta_scale = float(tta[0])
        p.tta_flip = int(tta[1] or 0)
        p.tta_img_h, p.tta_img_w = int(tta[2]), int(tta[3])
    scale = hyp.get("tar_box_scale_factor", [0.1, 0.1, 0.2, 0.2])
    for i in range(4):
        p.reg_scale[i] = float(scale[i])
    if family in ("retinanet", "retinanet_exp"):
        levels = (3, 4, 5, 6, 7)
        p.num_levels = len(levels)
        p.anchors_per_cell = 9
        for i, lvl in enumerate(levels):
            p.level_h[i] = (img_h - 1) // 2 ** lvl + 1
            p.level_w[i] = (img_w - 1) // 2 ** lvl + 1
            p.level_stride[i] = float(2 ** lvl)
            base = retinanet_base_anchors(2 ** (lvl + 2))
            for a in range(9):
                for c in range(4):
                    p.anchor[i][a][c] = float(base[a, c])
        return p
    if level_shapes is None:
        raise ValueError("level_shapes is required for grid families")
    p.num_levels = len(level_shapes)
    for i, (h, w) in enumerate(level_shapes):
        p.level_h[i], p.level_w[i] = int(h), int(w)
        if family == "fcos":
            # level_i_stride = self.inp_h / fm_h with inp_h from hyp['input_img_size'], not from the actual input
            # (trainer/eval_fcos.py:17,137): mirrored, so an input of another size decodes like the reference does
            p.level_stride[i] = float(int(hyp.get("input_img_size", (img_h, img_w))[0]) / h)
        elif family in _FIXED_STRIDES:
            p.level_stride[i] = float(_FIXED_STRIDES[family][i])
        else:
            p.level_stride[i] = float(img_h / h)  # eval_yolox.py:144, eval_yolov8.py:91
    if family in ("yolov5", "yolov7"):
        if anchors is None:
            raise ValueError("anchors (L, A, 2) are required for yolov5 / yolov7")
        anc = torch.as_tensor(anchors).detach().cpu()
        p.anchors_per_cell = int(anc.shape[1])
        for i in range(anc.shape[0]):
            # (self.anchors[i] / self.ds_scales[i]).type_as(inputs): true divide, then float32 (eval_yolov5.py:192)
            stage = (anc[i] / _FIXED_STRIDES[family][i]).to(torch.float32)
            for a in range(anc.shape[1]):
                p.anchor[i][a][0] = float(stage[a, 0])
                p.anchor[i][a][1] = float(stage[a, 1])
    elif family == "yolox":
        p.anchors_per_cell = int(num_anchors or hyp.get("num_anchors", 1))
    else:
        p.anchors_per_cell = 1
    return p


def letterbox_table(info):
    """list of the reference's per-image info dicts -> (b, 5) rows {scale, pad_top, pad_left, org_h, org_w}."""
    return [[float(d["scale"]), float(d["pad_top"]), float(d["pad_left"]), float(d["org_shape"][0]),
             float(d["org_shape"][1])] for d in info]


class DetectionList(list):
    """The evaluator's output list (CPU rows, the reference's contract) that also remembers the device rows it was read
    from, so that preds_postprocess can map them without staging them back to the GPU."""
    device_rows = None    # (b, max_det, 6) float32 CUDA tensor (a private copy, not the reused output buffer)
    device_cnt = None     # (b,) int32 CUDA tensor


def preds_postprocess(outputs, info):
    """val_yolov5.py:164-177 for an evaluator's output list: list[Tensor(K,6) | None] -> list[ndarray(K,6) | None].

    (The image half of the reference's method -- un-padding and resizing the input picture for drawing -- is
    visualisation and stays with the caller.)  When ``outputs`` is the DetectionList an evaluator mirror returned, the
    rows are still on the device: ysb_undo_letterbox runs on them and only the mapped rows come back.  Any other list is
    staged to the device first.  (Callers that pass ``info`` to the evaluator get the undo fused into the NMS kernel's
    row write and need neither.)
    """
    if len(outputs) != len(info):
        raise ValueError(f"need one info dict per image: {len(info)} != {len(outputs)}")
    if not outputs:
        return []
    if not torch.cuda.is_available():
        raise RuntimeError("yoloseries_b200 needs a CUDA device: the engine has no CPU fallback")
    lib = _lib.load()
    if isinstance(outputs, DetectionList) and outputs.device_rows is not None:
        dets, dcnt = outputs.device_rows.clone(), outputs.device_cnt
        dev = dets.device
        cnt = [(-1 if o is None else int(o.shape[0])) for o in outputs]
        kmax = dets.shape[1]
    else:
        dev = torch.device("cuda", torch.cuda.current_device())
        kmax = max([1] + [int(o.shape[0]) for o in outputs if o is not None])
        host = torch.zeros((len(outputs), kmax, 6), dtype=torch.float32)
        cnt = [(-1 if o is None else int(o.shape[0])) for o in outputs]
        for i, o in enumerate(outputs):
            if o is not None and o.shape[0]:
                host[i, : o.shape[0]] = torch.as_tensor(o, dtype=torch.float32).reshape(-1, 6)
        dets, dcnt = host.to(dev), torch.tensor(cnt, dtype=torch.int32).to(dev)
    table = torch.tensor(letterbox_table(info), dtype=torch.float32).to(dev)
    with torch.cuda.device(dev):
        _lib.check(lib.ysb_undo_letterbox(dets.data_ptr(), dcnt.data_ptr(), len(outputs), kmax, table.data_ptr(),
                                          ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream)), "ysb_undo_letterbox")
    back = dets[:, : max(max(cnt), 1)].cpu().numpy()
    return [None if c < 0 else back[i, :c].copy() for i, c in enumerate(cnt)]


def validate_heads(family, flat, num_class, anchors=None, dfl_bins=16, decoded_row_w=None):
    """Full shape check of the head tensors against the family's layout (include/ysb_postproc.h, enum ysb_family).  The C
    ABI receives bare pointers, so this is the only place a wrong channel count / batch / level count can be caught; the
    reference raises a reshape error in the same situations (e.g. trainer/eval_yolov5.py:194)."""
    C = int(num_class)

    def bad(msg):
        raise ValueError(f"{family} heads: {msg}")
    if not flat:
        bad("no tensors")
    b = flat[0].shape[0]
    for i, t in enumerate(flat):
        if t.dtype != torch.float32 or not t.is_contiguous():
            raise ValueError("head tensors must be contiguous float32")
        if t.dim() < 2 or t.shape[0] != b:
            bad(f"tensor {i} has batch {tuple(t.shape)[:1]}, expected {b}")
        if t.device != flat[0].device:
            bad("tensors live on different devices")
    if decoded_row_w is not None:
        if len(flat) != 1 or flat[0].dim() != 3 or flat[0].shape[2] != decoded_row_w:
            raise ValueError(f"decoded tensor must be (b, N, {decoded_row_w}), got {tuple(flat[0].shape)}")
        return
    if family in ("yolov5", "yolov7"):
        if anchors is None:
            bad("anchors (L, A, 2) are required")
        L, A = int(anchors.shape[0]), int(anchors.shape[1])
        if len(flat) != L:
            bad(f"{len(flat)} levels but {L} anchor groups")
        for i, t in enumerate(flat):
            want = (b, A * (5 + C)) if family == "yolov5" else (b, A)
            if family == "yolov5" and (t.dim() != 4 or tuple(t.shape[:2]) != want):
                bad(f"level {i} must be (b, {A * (5 + C)}, H, W), got {tuple(t.shape)}")
            if family == "yolov7" and (t.dim() != 5 or tuple(t.shape[:2]) != want or t.shape[4] != 5 + C):
                bad(f"level {i} must be (b, {A}, H, W, {5 + C}), got {tuple(t.shape)}")
    elif family == "yolox":
        na = flat[0].shape[1] if flat[0].dim() == 5 else -1
        for i, t in enumerate(flat):
            if t.dim() != 5 or t.shape[1] != na or t.shape[2] != 5 + C:
                bad(f"level {i} must be (b, {na}, {5 + C}, H, W), got {tuple(t.shape)}")
    elif family == "yolov8":
        for i, t in enumerate(flat):
            if t.dim() != 4 or t.shape[1] != 4 * int(dfl_bins) + C:
                bad(f"level {i} must be (b, {4 * int(dfl_bins) + C}, H, W), got {tuple(t.shape)}")
    elif family in ("retinanet", "retinanet_exp"):
        w = 5 if family == "retinanet_exp" else 4
        if len(flat) != 2:
            bad("expected (reg, cls)")
        reg, cls = flat
        if reg.dim() != 3 or reg.shape[2] != w or cls.dim() != 3 or cls.shape[2] != C or reg.shape[1] != cls.shape[1]:
            bad(f"expected reg (b, N, {w}) and cls (b, N, {C}), got {tuple(reg.shape)} and {tuple(cls.shape)}")
    elif family == "fcos":
        if len(flat) % 3:
            bad("expected three lists (cls, reg, ctr) of equal length")
        n = len(flat) // 3
        for i in range(n):
            c, r, q = flat[i], flat[n + i], flat[2 * n + i]
            hw = tuple(c.shape[2:])
            if c.dim() != 4 or c.shape[1] != C or r.dim() != 4 or r.shape[1] != 4 or tuple(r.shape[2:]) != hw or \
                    q.dim() != 4 or q.shape[1] != 1 or tuple(q.shape[2:]) != hw:
                bad(f"level {i} must be cls (b, {C}, H, W), reg (b, 4, H, W), ctr (b, 1, H, W), got "
                    f"{tuple(c.shape)}, {tuple(r.shape)}, {tuple(q.shape)}")
    else:
        raise ValueError(f"unknown family {family!r}")


class DetectionBuffers:
    """Per-call device outputs: rows (b, max_det, 6), candidate indices (b, max_det), counts (b)."""

    def __init__(self, batch, max_det, device):
        self.dets = torch.empty((batch, max_det, 6), dtype=torch.float32, device=device)
        self.det_idx = torch.empty((batch, max_det), dtype=torch.int32, device=device)
        self.det_cnt = torch.empty((batch,), dtype=torch.int32, device=device)


class PostProcessor:
    """decode -> filter -> top-k -> class-aware NMS -> post-filter for one family / one head geometry.

    Buffers are allocated once per (batch, geometry) and reused; every call enqueues on torch's current stream.
    """

    def __init__(self, family, hyp, anchors=None, compute_metric=False):
        self.family = family
        self.hyp = dict(hyp)
        self.anchors = anchors
        self.compute_metric = compute_metric
        self._lib = _lib.load()
        self._ops = _ops.load()
        self._cache = {}

    # ---- plumbing ------------------------------------------------------------------------------------------
    def _prepare(self, flat, batch, img_h, img_w, input_kind, tta=None):
        dev = flat[0].device
        if dev.type != "cuda":
            raise RuntimeError("yoloseries_b200 runs on CUDA devices only (no CPU fallback); got " + str(dev))
        if input_kind == _lib.INPUT_RAW_HEADS:
            validate_heads(self.family, flat, self.hyp["num_class"], self.anchors, self.hyp.get("reg", 16))
        else:
            for t in flat:
                if t.dtype != torch.float32 or not t.is_contiguous():
                    raise ValueError("head tensors must be contiguous float32")
        shapes = None
        if input_kind == _lib.INPUT_RAW_HEADS and self.family not in ("retinanet", "retinanet_exp"):
            shapes = _level_shapes(self.family, flat, self.hyp["num_class"])
        elif self.family not in ("retinanet", "retinanet_exp"):
            strides = _FIXED_STRIDES.get(self.family) or ((4, 8, 16, 32) if self.family == "yolov8" else (8, 16, 32))
            shapes = [(img_h // s, img_w // s) for s in strides]
        na = flat[0].shape[1] if (self.family == "yolox" and input_kind == _lib.INPUT_RAW_HEADS) else None
        rows = int(flat[0].shape[1]) if input_kind == _lib.INPUT_DECODED_ROWS else 0
        key = (batch, img_h, img_w, input_kind, tuple(shapes or ()), dev.index, na, rows, tta)
        ent = self._cache.get(key)
        if ent is None:
            params = make_params(self.family, self.hyp, batch, img_h, img_w, shapes, self.anchors, input_kind,
                                 self.compute_metric, na, rows, tta)
            n = ctypes.c_int64()
            rw = ctypes.c_int32()
            _lib.check(self._lib.ysb_num_candidates(ctypes.byref(params), ctypes.byref(n), ctypes.byref(rw)),
                       "ysb_num_candidates")
            ws = ctypes.c_size_t()
            _lib.check(self._lib.ysb_postprocess_workspace_bytes(ctypes.byref(params), ctypes.byref(ws)),
                       "ysb_postprocess_workspace_bytes")
            ent = dict(params=params, params_t=_ops.params_tensor(params), N=n.value, row_w=rw.value,
                       workspace=torch.empty(max(ws.value, 1), dtype=torch.uint8, device=dev),
                       out=DetectionBuffers(batch, params.max_det, dev))
            self._cache[key] = ent
        if input_kind == _lib.INPUT_RAW_HEADS and self.family in ("retinanet", "retinanet_exp") and flat[0].shape[1] != ent["N"]:
            raise ValueError(f"{self.family} heads: {flat[0].shape[1]} anchors per image, but a {img_h}x{img_w} input has {ent['N']}")
        if input_kind == _lib.INPUT_DECODED_ROWS:
            validate_heads(self.family, flat, self.hyp["num_class"], decoded_row_w=ent["row_w"])
        return ent

    @staticmethod
    def _stream(dev=None):
        """torch's current stream ON THE HEADS' DEVICE (not on whatever device happens to be current)."""
        return ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream)

    # ---- public ops ----------------------------------------------------------------------------------------
    def decode(self, heads, img_h, img_w):
        """do_inference: raw heads -> (b, N, C') float32 on the heads' device."""
        flat = flatten_heads(self.family, heads)
        batch = flat[0].shape[0]
        ent = self._prepare(flat, batch, img_h, img_w, _lib.INPUT_RAW_HEADS)
        return self._ops.decode(flat, ent["params_t"])

    def _letterbox(self, holder, params_list, info, batch, dev):
        """Stage the reference's per-image info dicts as the (b, 5) table the NMS kernel's row write reads (fused
        preds_postprocess, val_yolov5.py:166-172); ``info`` None switches the undo off."""
        ptr = None
        if info is not None:
            if len(info) != batch:
                raise ValueError(f"need one info dict per image: {len(info)} != {batch}")
            if "lb" not in holder:
                holder["lb"] = torch.empty((batch, 5), dtype=torch.float32, device=dev)
            holder["lb"].copy_(torch.tensor(letterbox_table(info), dtype=torch.float32), non_blocking=False)
            ptr = holder["lb"].data_ptr()
        for p in params_list:
            p.d_letterbox = ptr

    def run(self, heads, img_h, img_w, decoded=False, info=None):
        """Whole path on the device.  Returns DetectionBuffers (views into reused storage).  ``info``: optional list of
        the reference's letterbox dicts (scale, pad_top, pad_left, org_shape): rows come out in original-picture pixels."""
        flat = [heads] if decoded else flatten_heads(self.family, heads)
        batch = flat[0].shape[0]
        if batch == 0:  # the reference loops over zero images and returns []
            return DetectionBuffers(0, int(self.hyp["max_predictions_per_img"]), flat[0].device)
        kind = _lib.INPUT_DECODED_ROWS if decoded else _lib.INPUT_RAW_HEADS
        ent = self._prepare(flat, batch, img_h, img_w, kind)
        out = ent["out"]
        ws = ent["workspace"]
        dev = flat[0].device
        with torch.cuda.device(dev):
            self._letterbox(ent, [ent["params"]], info, batch, dev)
        # through the torch extension (csrc/torch_adapter.cpp): tensor checks, device guard, torch's current stream
        self._ops.postprocess(flat, ent["params_t"], ws, out.dets, out.det_idx, out.det_cnt)
        return out

    # ---- test-time augmentation (trainer/eval_yolov5.py:152-179 and the same method of every evaluator) ------------
    def _tta_entries(self, passes, org_hw):
        """passes: [(heads, img_h, img_w, scale, flip_axis), ...] -> [(flat heads, cache entry), ...]."""
        if not 1 <= len(passes) <= _lib.YSB_MAX_PASSES:
            raise ValueError(f"1..{_lib.YSB_MAX_PASSES} passes, got {len(passes)}")
        ents = []
        for heads, img_h, img_w, scale, flip in passes:
            flat = flatten_heads(self.family, heads)
            tta = (float(scale), int(flip or 0), int(org_hw[0]), int(org_hw[1]))
            ents.append((flat, self._prepare(flat, flat[0].shape[0], int(img_h), int(img_w), _lib.INPUT_RAW_HEADS, tta)))
        if len({e["params"].batch for _, e in ents}) != 1:
            raise ValueError("every pass must hold the same images")
        return ents

    def decode_tta(self, passes, org_hw):
        """test_time_augmentation: every pass decoded, un-scaled and un-flipped straight into its slot of the merged
        (b, sum N_i, C') tensor.  Returns (merged, [per-pass views])."""
        ents = self._tta_entries(passes, org_hw)
        batch, total = ents[0][1]["params"].batch, sum(e["N"] for _, e in ents)
        out = torch.empty((batch, total, ents[0][1]["row_w"]), dtype=torch.float32, device=ents[0][0][0].device)
        off, views = 0, []
        dev = out.device
        for flat, ent in ents:
            self._ops.decode_into(flat, ent["params_t"], out, off)
            views.append(out[:, off:off + ent["N"]])
            off += ent["N"]
        return out, views

    def run_tta(self, passes, org_hw, info=None):
        """The whole path over several augmented passes without the merged tensor (ysb_postprocess_tta).  Candidate
        indices (det_idx) count through the passes in order, like rows of the reference's concatenated tensor."""
        ents = self._tta_entries(passes, org_hw)
        dev = ents[0][0][0].device
        batch = ents[0][1]["params"].batch
        if batch == 0:
            return DetectionBuffers(0, int(self.hyp["max_predictions_per_img"]), dev)
        n = len(ents)
        key = ("tta", dev.index) + tuple(id(e) for _, e in ents)
        bundle = self._cache.get(key)
        if bundle is None:
            arr = (YsbParams * n)(*[e["params"] for _, e in ents])
            ws = ctypes.c_size_t()
            _lib.check(self._lib.ysb_postprocess_tta_workspace_bytes(arr, n, ctypes.byref(ws)),
                       "ysb_postprocess_tta_workspace_bytes")
            bundle = dict(params=arr, workspace=torch.empty(max(ws.value, 1), dtype=torch.uint8, device=dev),
                          out=DetectionBuffers(batch, arr[0].max_det, dev), keep=[e for _, e in ents])
            self._cache[key] = bundle
        flat_all = [t for flat, _ in ents for t in flat]
        per_pass = (ctypes.c_int32 * n)(*[len(flat) for flat, _ in ents])
        out, ws = bundle["out"], bundle["workspace"]
        with torch.cuda.device(dev):
            self._letterbox(bundle, [bundle["params"][i] for i in range(n)], info, batch, dev)
            _lib.check(self._lib.ysb_postprocess_tta(bundle["params"], n, _lib.head_pointer_array(flat_all), per_pass,
                                                     ws.data_ptr(), ws.numel(), out.dets.data_ptr(), out.det_idx.data_ptr(),
                                                     out.det_cnt.data_ptr(), self._stream(dev)), "ysb_postprocess_tta")
        return out

    def capture(self, heads, img_h, img_w, decoded=False):
        """Capture {zero counters, filter kernel, NMS kernel} for THESE head tensors into a CUDA graph.

        Returns a callable: every call replays the graph on the current stream (one driver call instead of a Python
        -> ctypes -> three launches round trip; +47 % images/s at 8 images per batch) and returns the DetectionBuffers.
        The head tensors are baked in by address: refill them in place (as a CUDA-graphed model does).
        """
        self.run(heads, img_h, img_w, decoded=decoded)  # warm-up outside capture: function attributes, buffers
        torch.cuda.synchronize()
        graph = torch.cuda.CUDAGraph()
        side = torch.cuda.Stream(device=(heads if decoded else flatten_heads(self.family, heads)[0]).device)
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.graph(graph, stream=side, capture_error_mode="thread_local"):
            out = self.run(heads, img_h, img_w, decoded=decoded)
        keep_alive = heads

        def replay():
            graph.replay()
            return out

        replay.graph, replay.heads = graph, keep_alive
        return replay

    def filter_only(self, heads, img_h, img_w, decoded=False):
        """K1 alone: returns (keys (b, N) uint64 as int64 tensor, counts (b, 4) int32)."""
        flat = [heads] if decoded else flatten_heads(self.family, heads)
        batch = flat[0].shape[0]
        kind = _lib.INPUT_DECODED_ROWS if decoded else _lib.INPUT_RAW_HEADS
        ent = self._prepare(flat, batch, img_h, img_w, kind)
        dev = flat[0].device
        slots = ent["N"] * (int(self.hyp["num_class"]) if self.hyp["mutil_label"] else 1)
        keys = torch.empty((batch, slots), dtype=torch.int64, device=dev)
        counts = torch.empty((batch, 4), dtype=torch.int32, device=dev)
        self._ops.filter_candidates(flat, ent["params_t"], keys, counts)
        return keys, counts

    def undo_letterbox(self, out, info):
        """The box part of val_yolov5.py:140-179 (preds_postprocess) on the kept rows, in place, before the D2H copy.

        ``info``: per image the reference's dict (``scale``, ``pad_top``, ``pad_left``, ``org_shape`` = (h, w)).
        x = clamp((x - pad_left) / scale, 1, org_w - 1), y likewise with pad_top / org_h; float32, op by op.
        """
        if out.det_cnt.numel() == 0:
            return out
        if len(info) != out.det_cnt.numel():
            raise ValueError(f"need one info dict per image: {len(info)} != {out.det_cnt.numel()}")
        dev = out.dets.device
        table = torch.tensor(letterbox_table(info), dtype=torch.float32).to(dev, non_blocking=True)
        with torch.cuda.device(dev):
            _lib.check(self._lib.ysb_undo_letterbox(out.dets.data_ptr(), out.det_cnt.data_ptr(), out.dets.shape[0],
                                                    out.dets.shape[1], table.data_ptr(), self._stream(dev)),
                       "ysb_undo_letterbox")
        return out

    @staticmethod
    def to_list(out, as_numpy=False, with_index=False):
        """DetectionBuffers -> the reference's output contract: list of CPU float32 (K, 6) tensors or None."""
        cnt = out.det_cnt.cpu()
        kmax = int(cnt.max().item()) if cnt.numel() else 0
        rows = out.dets[:, : max(kmax, 1)].cpu()
        idx = out.det_idx[:, : max(kmax, 1)].cpu() if with_index else None
        res, ids = [], []
        for i, c in enumerate(cnt.tolist()):
            if c < 0:
                res.append(None)
                ids.append(None)
                continue
            r = rows[i, :c].clone()
            res.append(r.numpy() if as_numpy else r)
            if with_index:
                ids.append(idx[i, :c].clone().numpy())
        if not with_index and not as_numpy:
            res = DetectionList(res)
            res.device_rows, res.device_cnt = out.dets[:, : max(kmax, 1)].clone(), out.det_cnt.clone()
        return (res, ids) if with_index else res


class PipelinedPostProcessor:
    """Several batches in flight: batch i runs filter -> NMS on CUDA stream ``i % lanes`` with its own buffers, so the
    NMS kernel of one batch overlaps the HBM-bound filter kernels of the next ones (the single-GPU form of
    dist.ShardedPostProcessor, which bench.py runs; 4 lanes hide the NMS kernel completely at 64 images per batch, batches
    below 64 images want 8).

        t = ppp.submit(heads, h, w)      # returns immediately
        dets = ppp.result(t)             # list[Tensor(K,6) | None], blocks on that batch only

    A ticket's buffers are reused ``lanes`` submissions later: fetch results before submitting that many new batches.
    """

    def __init__(self, family, hyp, anchors=None, compute_metric=False, lanes=4):
        self._pps = [PostProcessor(family, hyp, anchors, compute_metric) for _ in range(lanes)]
        self._streams = None
        self._next = 0

    def submit(self, heads, img_h, img_w, decoded=False):
        flat = [heads] if decoded else flatten_heads(self._pps[0].family, heads)
        dev = flat[0].device
        if self._streams is None:
            self._streams = [torch.cuda.Stream(device=dev) for _ in self._pps]
        i = self._next
        self._next = (i + 1) % len(self._pps)
        st = self._streams[i]
        st.wait_stream(torch.cuda.current_stream(dev))  # the heads were produced on the caller's stream
        with torch.cuda.stream(st):
            out = self._pps[i].run(heads, img_h, img_w, decoded=decoded)
            done = torch.cuda.Event()
            done.record(st)
        for t in flat:
            t.record_stream(st)
        return (i, done, out)

    def result(self, ticket, as_numpy=False, with_index=False):
        i, done, out = ticket
        done.synchronize()
        with torch.cuda.stream(self._streams[i]):
            return PostProcessor.to_list(out, as_numpy=as_numpy, with_index=with_index)
