"""dropin.install() against the REAL reference modules (the checkout in the build container, the staged byte-for-byte
copy baseline/_ref on the GPU box) instead of stubs, in a child process so that the patched modules never leak into the
other tests; and, on a GPU, the reference's own call shape through the patched names (val_yolov5.py:106-107, 388)."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _child(mode):
    sys.path.insert(0, ROOT)
    from oracle import refharness
    if not refharness.available():
        pytest.skip("neither /root/reference nor the staged copy baseline/_ref is present")
    p = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "helpers", "dropin_real.py"), mode], cwd=ROOT,
                       capture_output=True, text=True, timeout=900)
    assert p.returncode == 0, p.stdout[-2000:] + p.stderr[-3000:]
    return json.loads(p.stdout.strip().splitlines()[-1])


def test_install_rebinds_the_real_reference_modules():
    rep = _child("structure")
    assert rep["utils_rebound"] and rep["submodules_rebound"] and rep["eval_modules_rebound"], rep
    assert rep["evaluators_rebound"] and rep["compute_tp_rebound"] and not rep["old_still_bound"], rep
    assert "utils.mAP_v2.compute_tp" in rep["done"] and "trainer.YOLOV5Evaluator" in rep["done"]


@pytest.mark.gpu
def test_reference_call_shape_through_the_patched_names():
    rep = _child("run")
    rep.pop("done")
    assert rep["evaluators_rebound"] and rep["iou_family_rebound"], rep
    assert rep["rows_equal_reference"] and sum(max(k, 0) for k in rep["kept"]) > 0, json.dumps(rep)
    assert rep["compute_tp_equal_reference"] and rep["numba_nms_equal_reference"], rep
