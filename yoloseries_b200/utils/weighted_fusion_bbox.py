"""Drop-in mirror of utils/weighted_fusion_bbox.py (hyp['wfb']); compute runs in
libysb_postproc.so (ysb_wbf), no CPU path.

    weighted_fusion_bbox(bbox_list, iou_thr=0.5) -> (Cluster, Fusion)      utils/weighted_fusion_bbox.py:63-96

``bbox_list``: (N, 7) [xmin, ymin, xmax, ymax, score, class, weight].  ``Fusion``: per label (ascending) the list of fused
boxes, float64 arrays [xmin, ymin, xmax, ymax, score, class]; ``Cluster``: per label, per fused box, its member rows as
lists of 7 floats in the order they joined.  The fused box is the reference's own arithmetic (the score-weighted mean
divided by the member count once more, update_fusion_bbox :41-60).

The reference as shipped cannot run this function: ``cpu_iou`` calls ``np.clip(x, a_min=1e-6)`` without ``a_max``
(utils/bbox_tools.py:82), a TypeError under every numpy since 1.17.  The mirror implements the evident intent
(``a_max=None``); the golden vectors the tests compare against were produced by the reference with exactly that one
argument supplied (see the fixture generator's ``_ClipShim``).
"""
import ctypes

import numpy as np
import torch

from .. import _lib
from ._common import cuda_device, stream_ptr

__all__ = ["weighted_fusion_bbox", "fuse_batch"]


def fuse_batch(rows, counts, iou_thr, row_width, want_clusters=False):
    """ysb_wbf on device rows (batch, stride, row_width) / counts (batch,) int32.

    Returns per image ``None`` (no rows) or ``(fusion (K, 6) float64 ndarray in (label asc, creation) order, clusters)``
    where ``clusters`` is a list over the K fused boxes of row-index lists (indices into the image's rows, joining
    order), or None when not requested.  Raises IndexError where the reference would (a label whose best box does not
    overlap itself: degenerate box)."""
    lib = _lib.load()
    batch, stride = int(rows.shape[0]), int(rows.shape[1])
    dev = rows.device
    if batch == 0:
        return []
    if stride == 0:
        return [None] * batch
    ws_bytes = ctypes.c_size_t()
    _lib.check(lib.ysb_wbf_workspace_bytes(batch, stride, ctypes.byref(ws_bytes)), "ysb_wbf_workspace_bytes")
    ws = torch.empty(ws_bytes.value, dtype=torch.uint8, device=dev)
    order = torch.empty((batch, stride), dtype=torch.int32, device=dev)
    fusion = torch.empty((batch, stride, 6), dtype=torch.float64, device=dev)
    members = torch.empty((batch, stride), dtype=torch.int32, device=dev)
    status = torch.empty((batch,), dtype=torch.int32, device=dev)
    pair_cap = 4 * batch * stride if want_clusters else 0
    while True:
        pairs = torch.empty((max(pair_cap, 1), 2), dtype=torch.int32, device=dev) if want_clusters else None
        pcount = torch.zeros(1, dtype=torch.int64, device=dev) if want_clusters else None
        with torch.cuda.device(dev):
            _lib.check(lib.ysb_wbf(rows.data_ptr(), row_width, counts.data_ptr(), batch, stride, float(iou_thr), ws.data_ptr(),
                                   ws.numel(), order.data_ptr(), fusion.data_ptr(), members.data_ptr(),
                                   pairs.data_ptr() if want_clusters else None, pair_cap,
                                   pcount.data_ptr() if want_clusters else None, status.data_ptr(), stream_ptr()), "ysb_wbf")
        if not want_clusters or int(pcount.item()) <= pair_cap:
            break
        pair_cap = int(pcount.item())      # a box may join several clusters: rerun with the exact size
    if int(status.max().item()) != 0:
        raise IndexError("too many indices for array: a label's best box does not overlap itself (degenerate box); the "
                         "reference fails the same way in update_fusion_bbox (utils/weighted_fusion_bbox.py:47)")
    cnt_h = counts.cpu().numpy()
    mem_h, fus_h, ord_h = members.cpu().numpy(), fusion.cpu().numpy(), order.cpu().numpy()
    pairs_h = pairs[: int(pcount.item())].cpu().numpy() if want_clusters else None
    out = []
    for i in range(batch):
        n = int(cnt_h[i])
        if n == 0:
            out.append(None)
            continue
        slots = np.nonzero(mem_h[i, :n] > 0)[0]
        clusters = None
        if want_clusters:
            clusters = []
            for s in slots:   # sorted positions ascend in visiting order = the order the rows joined
                pos = np.sort(pairs_h[pairs_h[:, 0] == i * stride + s][:, 1]) - i * stride
                clusters.append([int(ord_h[i, p]) for p in pos])
        out.append((fus_h[i, slots], clusters))
    return out


def weighted_fusion_bbox(bbox_list, iou_thr=0.5):
    """utils/weighted_fusion_bbox.py:63-96."""
    arr = np.asarray(bbox_list)
    assert arr.ndim == 2 and arr.shape[1] >= 7
    rows32 = np.ascontiguousarray(arr[:, :7], dtype=np.float32)
    dev = cuda_device()
    n = rows32.shape[0]
    if n == 0:
        return [], []
    rows = torch.from_numpy(rows32).to(dev).reshape(1, n, 7)
    counts = torch.tensor([n], dtype=torch.int32, device=dev)
    fus, clusters = fuse_batch(rows, counts, iou_thr, 7, want_clusters=True)[0]
    Cluster, Fusion = [], []
    labels = fus[:, 5]
    for lab in np.unique(labels):
        sel = np.nonzero(labels == lab)[0]
        Fusion.append([fus[k].copy() for k in sel])
        Cluster.append([[arr[r][:7].tolist() for r in clusters[k]] for k in sel])
    return Cluster, Fusion
