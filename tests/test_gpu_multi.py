"""Multi-GPU path on real devices (skipped with fewer than 2 GPUs): the P2P detection gather and its NCCL fallback."""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("mode", ["p2p", "nccl"])
def test_sharded_postprocessor_gathers_every_ranks_rows(mode):
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs at least 2 GPUs")
    world = 4 if n >= 4 else 2
    port = 29600 + os.getpid() % 300 + (7 if mode == "nccl" else 0)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr",
           "127.0.0.1", "--master-port", str(port), os.path.join(ROOT, "tests", "helpers", "mgpu_gather_worker.py"), mode]
    r = subprocess.run(cmd, cwd=ROOT, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "MGPU_GATHER_OK" in r.stdout, (r.stdout[-3000:], r.stderr[-3000:])


def test_heads_on_another_device_than_the_current_one():
    """The engine switches to the heads' device for every C-ABI call (kernels, function attributes and the stream belong to
    that device): heads on cuda:1 while cuda:0 is current give the rows cuda:0 gives for the same heads."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs at least 2 GPUs")
    import numpy as np

    import oracle
    from yoloseries_b200 import synth
    from yoloseries_b200.engine import PostProcessor
    hyp = oracle.default_hyp()
    anchors = torch.tensor(synth.V5_ANCHORS_PX)
    torch.cuda.set_device(0)
    for family, img in (("yolov5", 320), ("yolox", 320), ("retinanet", 256)):
        a = anchors if family == "yolov5" else None
        heads0 = synth.make_heads(family, 3, img, img, 80, "crowd" if family == "yolov5" else "dense", seed=5, device="cuda:0")
        heads1 = [h.to("cuda:1") for h in heads0] if isinstance(heads0, list) else tuple(h.to("cuda:1") for h in heads0)
        pp0, pp1 = PostProcessor(family, hyp, anchors=a), PostProcessor(family, hyp, anchors=a)
        want = pp0.to_list(pp0.run(heads0, img, img), as_numpy=True)
        assert torch.cuda.current_device() == 0
        out1 = pp1.run(heads1, img, img)
        assert out1.dets.device == torch.device("cuda", 1)
        got = pp1.to_list(out1, as_numpy=True)
        dec1 = pp1.decode(heads1, img, img)
        assert dec1.device == torch.device("cuda", 1)
        torch.testing.assert_close(dec1.cpu(), pp0.decode(heads0, img, img).cpu(), rtol=0, atol=0)
        assert torch.cuda.current_device() == 0
        for g, w in zip(got, want):
            assert (g is None) == (w is None)
            if g is not None:
                np.testing.assert_array_equal(g, w)
