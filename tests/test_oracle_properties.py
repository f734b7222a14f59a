"""Property tests of the oracle's C loops against a literal pure-Python restatement of utils/nms.py:10-27 and
utils/bbox_tools.py:12-35 on small random cases (ties, zero scores, degenerate and identical boxes), plus a static check
that nothing that runs on the GPU box reads the reference checkout."""
import glob
import os

import numpy as np
from hypothesis import given, settings, strategies as st

import oracle

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
F = np.float32


def py_iou_row(b1, boxes):
    """utils/bbox_tools.py:12-35 for one box against many: float32 areas and min/max, float64 from max(0., .) on."""
    out = np.zeros(len(boxes), dtype=np.float64)
    a1 = F(F(b1[2] - b1[0]) * F(b1[3] - b1[1]))
    for j, b2 in enumerate(boxes):
        a2 = F(F(b2[2] - b2[0]) * F(b2[3] - b2[1]))
        w = max(0.0, float(F(min(b1[2], b2[2]) - max(b1[0], b2[0]))))
        h = max(0.0, float(F(min(b1[3], b2[3]) - max(b1[1], b2[1]))))
        inter = w * h
        den = float(F(a1 + a2)) - inter
        with np.errstate(divide="ignore", invalid="ignore"):
            out[j] = np.float64(inter) / np.float64(den)
    return out


def py_nms(boxes, scores, thr):
    """utils/nms.py:10-27: while sum > 0: argmax (first maximum), keep, zero it, zero everything with iou >= thr."""
    scores = scores.copy()
    keep = []
    while scores.sum() > 0:
        i = int(np.argmax(scores))
        keep.append(i)
        scores[i] = 0
        iou = py_iou_row(boxes[i], boxes)
        scores[iou >= thr] = 0          # NaN >= thr is False, as in the reference
    return keep


coords = st.integers(min_value=0, max_value=12)


@st.composite
def box_sets(draw):
    n = draw(st.integers(min_value=1, max_value=14))
    boxes = []
    for _ in range(n):
        x1, y1 = draw(coords), draw(coords)
        w, h = draw(st.integers(min_value=0, max_value=8)), draw(st.integers(min_value=0, max_value=8))  # 0: degenerate
        cls = draw(st.integers(min_value=0, max_value=2))
        boxes.append([x1 + 4096 * cls, y1, x1 + w + 4096 * cls, y1 + h])
    scores = [draw(st.sampled_from([0.0, 0.1, 0.25, 0.5, 0.5, 0.75, 0.9])) for _ in range(n)]   # ties and zeros
    thr = draw(st.sampled_from([0.2, 0.5, 0.65, 1.0]))
    return np.array(boxes, dtype=F), np.array(scores, dtype=F), thr


@settings(max_examples=150, deadline=None)
@given(box_sets())
def test_c_nms_equals_literal_python(case):
    boxes, scores, thr = case
    assert oracle.numba_nms(boxes, scores, thr) == py_nms(boxes, scores, thr)
    # the kept list is prefix-stable: stopping early returns the head of the full list (what the CUDA kernel relies on)
    full = oracle.numba_nms(boxes, scores, thr)
    for k in (1, 2, 5):
        assert oracle.numba_nms(boxes, scores, thr, max_keep=k) == full[:k]


@settings(max_examples=100, deadline=None)
@given(box_sets())
def test_c_iou_equals_literal_python(case):
    boxes, _, _ = case
    got = oracle.numba_iou(boxes, boxes)
    want = np.stack([py_iou_row(b, boxes) for b in boxes])
    np.testing.assert_array_equal(got, want)          # NaN == NaN here (0/0 of two degenerate boxes)
    assert got.dtype == np.float64


def test_gpu_side_code_never_reads_the_reference_checkout():
    """/root/reference does not exist on the GPU box.  The product package and smoke() never touch the reference at all;
    the -m gpu tests and bench.py reach it only through oracle/refharness.py / baseline/ref_worker.py, which fall back to
    the byte-for-byte staged copy baseline/_ref/ (baseline/stage_reference.py) -- never through a hard-coded path."""
    strict = glob.glob(os.path.join(ROOT, "yoloseries_b200", "**", "*.py"), recursive=True) + [os.path.join(ROOT, "__graft_entry__.py")]
    for f in strict:
        src = open(f).read()
        assert "refharness" not in src and "gen_golden" not in src and "ref_worker" not in src and "baseline/_ref" not in src, f
        assert "import oracle" not in src or f.endswith("__graft_entry__.py"), f
    loose = glob.glob(os.path.join(ROOT, "tests", "test_gpu_*.py")) + [os.path.join(ROOT, "bench.py")]
    for f in loose:
        src = open(f).read()
        assert '"/root/reference' not in src and "'/root/reference" not in src and "gen_golden" not in src, f
    from oracle import refharness
    assert refharness.STAGED_ROOT.endswith(os.path.join("baseline", "_ref"))
