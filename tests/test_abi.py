"""CPU-side checks of the C-ABI boundary: the library loads, exports every symbol include/ysb_postproc.h declares,
validates arguments and reports sizes -- no compute call is made (there is no GPU in the build container)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "ysb_postproc.h")


@pytest.fixture(scope="module")
def lib():
    from yoloseries_b200 import _lib, build
    build.build()
    return _lib.load()


def declared_symbols():
    text = open(HEADER).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(ysb_[a-z_0-9]+)\s*\(", text)))


def test_header_symbols_are_exported(lib):
    from yoloseries_b200 import _lib
    names = declared_symbols()
    assert len(names) >= 12
    for n in names:
        assert hasattr(lib, n), f"{n} declared in the header but not exported"
    assert sorted(_lib.EXPORTED_SYMBOLS) == names, "ctypes signature table out of sync with the header"


def test_struct_layout_matches_header(lib):
    """sizeof(ysb_params) as ctypes sees it must equal what the C compiler sees."""
    import subprocess
    import tempfile
    from yoloseries_b200._lib import YsbParams
    src = '#include <stdio.h>\n#include "ysb_postproc.h"\nint main(){printf("%zu %zu %zu", sizeof(ysb_params), ' \
          '__builtin_offsetof(ysb_params, iou_thr), __builtin_offsetof(ysb_params, decoded_rows));return 0;}\n'
    with tempfile.TemporaryDirectory() as d:
        c = os.path.join(d, "s.c")
        open(c, "w").write(src)
        exe = os.path.join(d, "s")
        subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), c, "-o", exe])
        size, off_iou, off_rows = map(int, subprocess.check_output([exe]).split())
    assert ctypes.sizeof(YsbParams) == size
    assert YsbParams.iou_thr.offset == off_iou
    assert YsbParams.decoded_rows.offset == off_rows


def _v5_params(**over):
    import torch
    from yoloseries_b200 import engine
    import oracle
    hyp = oracle.default_hyp()
    hyp.update(over)
    return engine.make_params("yolov5", hyp, 4, 640, 640, [(80, 80), (40, 40), (20, 20)],
                              anchors=torch.tensor([[[10, 13], [16, 30], [33, 23]], [[30, 61], [62, 45], [59, 119]],
                                                    [[116, 90], [156, 198], [373, 326]]]))


def test_num_candidates_and_workspace(lib):
    p = _v5_params()
    n, rw, ws = ctypes.c_int64(), ctypes.c_int32(), ctypes.c_size_t()
    assert lib.ysb_num_candidates(ctypes.byref(p), ctypes.byref(n), ctypes.byref(rw)) == 0
    assert (n.value, rw.value) == (25200, 85)
    assert lib.ysb_postprocess_workspace_bytes(ctypes.byref(p), ctypes.byref(ws)) == 0
    assert ws.value >= 4 * 25200 * 8 + 4 * 16
    assert abs(p.anchor[0][0][0] - 10 / 8) < 1e-7 and abs(p.anchor[2][2][1] - 326 / 32) < 1e-7


@pytest.mark.parametrize("family,img,expect", [("yolox", 640, 8400), ("yolov8", 640, 34000), ("fcos", 640, 8525),
                                               ("retinanet", 640, 76725), ("yolov5", 1280, 100800)])
def test_candidate_counts_of_baseline_configs(lib, family, img, expect):
    import torch
    import oracle
    from yoloseries_b200 import engine, synth
    hyp = oracle.default_hyp()
    shapes = None if family.startswith("retinanet") else synth.level_shapes(family, img, img)
    anchors = torch.tensor(synth.V5_ANCHORS_PX) if family == "yolov5" else None
    p = engine.make_params(family, hyp, 1, img, img, shapes, anchors=anchors)
    n = ctypes.c_int64()
    assert lib.ysb_num_candidates(ctypes.byref(p), ctypes.byref(n), None) == 0
    assert n.value == expect == synth.num_candidates(family, img, img)


def test_argument_validation(lib):
    from yoloseries_b200 import _lib
    n = ctypes.c_int64()
    assert lib.ysb_num_candidates(None, ctypes.byref(n), None) == _lib.YSB_ERR_BAD_ARG
    p = _v5_params()
    p.family = 99
    assert lib.ysb_num_candidates(ctypes.byref(p), ctypes.byref(n), None) == _lib.YSB_ERR_BAD_ARG
    p = _v5_params(max_predictions_per_img=5000)
    assert lib.ysb_num_candidates(ctypes.byref(p), ctypes.byref(n), None) == _lib.YSB_ERR_LIMIT
    import oracle
    from yoloseries_b200 import engine
    p = engine.make_params("retinanet", oracle.default_hyp(mutil_label=True), 1, 64, 64)  # broken in the reference too
    assert lib.ysb_num_candidates(ctypes.byref(p), ctypes.byref(n), None) == _lib.YSB_ERR_UNSUPPORTED
    p = _v5_params(mutil_label=True)
    ws1, ws80 = ctypes.c_size_t(), ctypes.c_size_t()
    assert lib.ysb_postprocess_workspace_bytes(ctypes.byref(_v5_params()), ctypes.byref(ws1)) == 0
    assert lib.ysb_postprocess_workspace_bytes(ctypes.byref(p), ctypes.byref(ws80)) == 0
    assert ws80.value > 70 * ws1.value  # one key slot per (candidate, class)
    p = _v5_params()
    assert lib.ysb_postprocess(ctypes.byref(p), None, 3, None, 0, None, None, None, None) == _lib.YSB_ERR_BAD_ARG
    ws = ctypes.c_size_t()
    assert lib.ysb_nms_workspace_bytes(-1, ctypes.byref(ws)) == _lib.YSB_ERR_BAD_ARG
    assert lib.ysb_nms_workspace_bytes(10 ** 9, ctypes.byref(ws)) == _lib.YSB_ERR_LIMIT
    assert lib.ysb_status_string(_lib.YSB_ERR_WORKSPACE) == b"workspace too small"
    # soft-NMS / letterbox entry points validate before touching the device
    assert lib.ysb_soft_nms(None, None, 0, 0.3, _lib.GIOU, 0, 0.0, None, 0, None, None) == _lib.YSB_OK
    assert lib.ysb_soft_nms(None, None, 4, 0.3, _lib.GIOU, 0, 0.0, None, 0, None, None) == _lib.YSB_ERR_BAD_ARG
    assert lib.ysb_soft_nms(1, 1, 4, 0.3, _lib.IOU_NUMBA_F64MIX, 0, 0.0, 1, 16, 1, None) == _lib.YSB_ERR_BAD_ARG
    assert lib.ysb_soft_nms(1, 1, 4, 0.3, _lib.DIOU, 1, 0.0, 1, 16, 1, None) == _lib.YSB_ERR_BAD_ARG  # sigma must be > 0
    assert lib.ysb_soft_nms(16, 1, 4, 0.3, _lib.DIOU, 0, 0.0, 1, 15, 1, None) == _lib.YSB_ERR_WORKSPACE
    assert lib.ysb_soft_nms(20, 1, 4, 0.3, _lib.DIOU, 0, 0.0, 1, 16, 1, None) == _lib.YSB_ERR_BAD_ARG  # boxes not 16-byte aligned
    # detection gather: geometry and struct validation happen before any device call
    nb = ctypes.c_size_t()
    assert lib.ysb_gather_buffer_bytes(8, 4, 64, 300, ctypes.byref(nb)) == _lib.YSB_OK
    assert nb.value >= 4 * 8 * 64 * 300 * 24 + 4 * 8 * 64 * 4
    assert lib.ysb_gather_buffer_bytes(_lib.YSB_MAX_PEERS + 1, 4, 64, 300, ctypes.byref(nb)) == _lib.YSB_ERR_BAD_ARG
    assert lib.ysb_gather_buffer_bytes(2, 0, 64, 300, ctypes.byref(nb)) == _lib.YSB_ERR_BAD_ARG
    assert lib.ysb_gather_buffer_bytes(2, 2, 64, 5000, ctypes.byref(nb)) == _lib.YSB_ERR_LIMIT
    g = _lib.YsbGather()
    g.world, g.rank, g.slots, g.batch, g.max_det = 2, 0, 2, 4, 300
    assert lib.ysb_gather_wait(ctypes.byref(g), 0, None) == _lib.YSB_ERR_BAD_ARG      # peer buffers not mapped
    assert lib.ysb_gather_begin(ctypes.byref(g), 0, None, 0, None) == _lib.YSB_ERR_BAD_ARG
    g.d_buf[0], g.d_buf[1] = 256, 512
    assert lib.ysb_gather_wait(ctypes.byref(g), 2, None) == _lib.YSB_ERR_BAD_ARG      # slot out of range
    assert lib.ysb_gather_open(None, None) == _lib.YSB_ERR_BAD_ARG
    assert lib.ysb_undo_letterbox(None, None, 0, 300, None, None) == _lib.YSB_OK
    assert lib.ysb_undo_letterbox(None, None, 2, 300, None, None) == _lib.YSB_ERR_BAD_ARG
    assert lib.ysb_undo_letterbox(1, 1, 2, 0, 1, None) == _lib.YSB_ERR_BAD_ARG
    with pytest.raises(ValueError):
        _lib.check(_lib.YSB_ERR_BAD_ARG, "x")
    with pytest.raises(NotImplementedError):
        _lib.check(_lib.YSB_ERR_UNSUPPORTED, "x")


def test_no_cpu_fallback():
    """Without a CUDA device the product path must fail loudly, not compute on the host."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    import numpy as np
    from yoloseries_b200.utils import numba_nms
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        numba_nms(np.zeros((3, 4), np.float32), np.ones(3, np.float32), 0.5)
    from yoloseries_b200.engine import PostProcessor
    import oracle
    pp = PostProcessor("yolov5", oracle.default_hyp(), anchors=torch.tensor([[[1, 1]] * 3] * 3))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        pp.run([torch.zeros(1, 255, 8, 8), torch.zeros(1, 255, 4, 4), torch.zeros(1, 255, 2, 2)], 64, 64)


def test_product_never_imports_oracle():
    """The oracle is test infrastructure: nothing under yoloseries_b200/ may reference it."""
    pkg = os.path.join(ROOT, "yoloseries_b200")
    for d, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(d, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", text, flags=re.M), f"{f} imports the oracle"


def test_tta_entry_points_validate_on_the_host(lib):
    """ysb_postprocess_tta / ysb_decode_into: pass lists are checked before anything touches the device."""
    from yoloseries_b200 import _lib
    from yoloseries_b200._lib import YsbParams
    base = _v5_params()
    ws1, ws3 = ctypes.c_size_t(), ctypes.c_size_t()
    assert lib.ysb_postprocess_workspace_bytes(ctypes.byref(base), ctypes.byref(ws1)) == 0

    def passes(*ps):
        return (YsbParams * len(ps))(*ps)

    p1, p2 = _v5_params(), _v5_params()
    p1.tta_scale, p1.tta_flip, p1.tta_img_h, p1.tta_img_w = 0.83, 2, 640, 640
    p2.tta_scale, p2.tta_flip, p2.tta_img_h, p2.tta_img_w = 0.67, 3, 640, 640
    arr = passes(base, p1, p2)
    assert lib.ysb_postprocess_tta_workspace_bytes(arr, 3, ctypes.byref(ws3)) == 0
    assert ws3.value >= 3 * 4 * 25200 * 8                      # one key slot per candidate of every pass
    assert lib.ysb_postprocess_tta_workspace_bytes(arr, 0, ctypes.byref(ws3)) == _lib.YSB_ERR_BAD_ARG
    assert lib.ysb_postprocess_tta_workspace_bytes(arr, 5, ctypes.byref(ws3)) == _lib.YSB_ERR_BAD_ARG
    assert lib.ysb_postprocess_tta_workspace_bytes(None, 3, ctypes.byref(ws3)) == _lib.YSB_ERR_BAD_ARG
    bad = _v5_params()
    bad.tta_flip = 1                                           # the reference flips along dims 2 or 3 only
    assert lib.ysb_postprocess_tta_workspace_bytes(passes(base, bad), 2, ctypes.byref(ws3)) == _lib.YSB_ERR_BAD_ARG
    other = _v5_params()
    other.iou_thr = 0.5                                        # one evaluator, one hyp: thresholds must agree
    assert lib.ysb_postprocess_tta_workspace_bytes(passes(base, other), 2, ctypes.byref(ws3)) == _lib.YSB_ERR_BAD_ARG
    dec = _v5_params()
    dec.input_kind = _lib.INPUT_DECODED_ROWS                   # decoded rows already carry their undo
    assert lib.ysb_postprocess_tta_workspace_bytes(passes(base, dec), 2, ctypes.byref(ws3)) == _lib.YSB_ERR_BAD_ARG
    dec.tta_flip = 2
    n = ctypes.c_int64()
    assert lib.ysb_num_candidates(ctypes.byref(dec), ctypes.byref(n), None) == _lib.YSB_ERR_BAD_ARG
    assert lib.ysb_postprocess_tta(arr, 3, None, None, None, 0, None, None, None, None) == _lib.YSB_ERR_BAD_ARG
    assert lib.ysb_decode_into(ctypes.byref(base), None, 3, None, 25200, 0, None) == _lib.YSB_ERR_BAD_ARG


def test_torch_extension_loads_and_registers_the_operators():
    """csrc/torch_adapter.cpp -> _lib/libysb_torch.so: TORCH_LIBRARY(ysb) over the C ABI.  Loadable without a GPU; every
    operator refuses CPU tensors (no fallback) with the reference's exception type, before any CUDA call."""
    import torch

    from yoloseries_b200 import _lib, _ops
    ops = _ops.load()
    assert ops.abi_version() == _lib.ABI_VERSION
    assert ops.params_bytes() == ctypes.sizeof(_lib.YsbParams)
    for name in ("postprocess", "filter_candidates", "select_nms", "decode", "decode_into", "nms", "pairwise_iou",
                 "pairwise_iou_backward", "elementwise_iou", "elementwise_iou_backward"):
        assert hasattr(ops, name), name
    p = _lib.YsbParams()
    blob = _ops.params_tensor(p)
    p.batch = 5                                           # the tensor shares the struct's memory
    assert blob.numel() == ctypes.sizeof(_lib.YsbParams) and int(blob[8:12].view(torch.int32)[0]) == 5
    with pytest.raises(ValueError):
        ops.pairwise_iou(torch.zeros(2, 4), torch.zeros(3, 4), _lib.IOU_F32)
    with pytest.raises(ValueError):
        ops.nms(torch.zeros(2, 4), torch.zeros(2), 0.5, _lib.CMP_GE, _lib.IOU_NUMBA_F64MIX, 0)
    with pytest.raises(ValueError):
        ops.decode([torch.zeros(1, 255, 8, 8)], blob)
    with pytest.raises(ValueError):
        ops.decode([], blob)
    with pytest.raises(ValueError):                       # a params blob of the wrong size
        ops.filter_candidates([torch.zeros(1)], torch.zeros(10, dtype=torch.uint8), torch.zeros(1, 1, dtype=torch.int64),
                              torch.zeros(1, 4, dtype=torch.int32))
