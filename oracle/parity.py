"""Stage-attributed parity report: fused CUDA path vs the reference's own outputs on the same raw heads.

TEST INFRASTRUCTURE ONLY (imported by tests/ and by bench.py's parity leg, never by the product package).

The north star asks for "kept-index sets bit-exact vs the reference, disagreements from IoU values within 1e-6 of the
threshold counted and reported; decoded boxes and scores within 1e-5 relative".  Given, per image,
  * the reference's decoded tensor and kept rows (XEvaluator.do_inference / numba_nms run unmodified, baseline/ref_worker.py),
  * the engine's sort keys (ysb_filter_candidates), kept rows, candidate indices and counts (ysb_postprocess),
this module reports where the two agree and, where they do not, which stage the first disagreement belongs to:
  filter   the survivor sets differ, or a survivor's score bits differ (decode arithmetic: sigmoid/exp ulps)
  order    same survivors, different visiting order (only possible through differing score bits)
  nms      same candidates visited in the same order, different keep decision (|IoU - thr| is reported)
  rows     same kept candidates, different row values (box decode ulps)
"""
import numpy as np

from . import cnms
from .pipeline import evaluator_nms

CAND_MASK = (1 << 22) - 1


def unpack_keys(keys_u64):
    """uint64 sort keys -> (score float32, cand int64, cls int64)."""
    k = np.asarray(keys_u64).astype(np.uint64)
    score = (k >> np.uint64(32)).astype(np.uint32).view(np.float32)
    cand = CAND_MASK - ((k >> np.uint64(10)) & np.uint64(CAND_MASK)).astype(np.int64)
    cls = 1023 - (k & np.uint64(1023)).astype(np.int64)
    return score, cand, cls


def _ulp_diff(a, b):
    a = np.asarray(a, dtype=np.float32).view(np.int32).astype(np.int64)
    b = np.asarray(b, dtype=np.float32).view(np.int32).astype(np.int64)
    return np.abs(a - b)


def image_report(family, hyp, ref_decoded, ref_rows, ref_count, gpu_keys, gpu_m, gpu_rows, gpu_idx, gpu_count,
                 near_eps=1e-6):
    """One image.  ref_decoded (N, C'); ref_rows (K, 6) valid rows only; ref_count (-1 = None); gpu_keys (cap,) uint64 with
    gpu_m valid entries; gpu_rows (K', 6), gpu_idx (K',), gpu_count."""
    rep = {"stage": "ok", "ref_rows": int(max(ref_count, 0)), "gpu_rows": int(max(gpu_count, 0))}
    want = evaluator_nms(family, ref_decoded[None], hyp)[0]
    # the pinned oracle must reproduce the reference on the reference's own decoded tensor; if it does not, the
    # attribution below (which needs candidate indices for the reference's rows) is void -- say so
    if ref_count < 0:
        rep["oracle_reproduces_reference"] = want.rows is None
    else:
        if family.startswith("retinanet"):
            same = want.rows is not None and want.rows.shape == ref_rows.shape and \
                np.array_equal(want.rows[:, 4:], ref_rows[:, 4:]) and np.allclose(want.rows[:, :4], ref_rows[:, :4], rtol=1e-5, atol=1e-5)
        else:
            same = want.rows is not None and want.rows.shape == ref_rows.shape and np.array_equal(want.rows, ref_rows)
        rep["oracle_reproduces_reference"] = bool(same)
    # ---- filter stage: survivor sets and score bits ------------------------------------------------------------
    g_score, g_cand, g_cls = unpack_keys(gpu_keys[:gpu_m])
    multi = bool(hyp.get("mutil_label"))
    r_cand = np.asarray(want.survivors, dtype=np.int64)
    r_score = np.asarray(want.nms_scores, dtype=np.float32)
    if family == "fcos":   # the NMS array holds sqrt(score) of the top-k; compare on the squared-back key only by index
        r_key = {int(c): None for c in r_cand}
    else:
        r_key = {int(c): s for c, s in zip(r_cand, r_score)} if not multi else None
    if r_key is not None:
        g_set = set(int(c) for c in g_cand)
        r_set = set(r_key)
        if family == "fcos":
            only_ref = len(r_set - g_set)   # the GPU list is pre-top-k: it must contain the reference's top-k
            only_gpu = 0
        else:
            only_ref, only_gpu = len(r_set - g_set), len(g_set - r_set)
        rep["survivors_only_in_reference"] = only_ref
        rep["survivors_only_in_engine"] = only_gpu
        if family != "fcos":
            common = [i for i, c in enumerate(g_cand) if int(c) in r_key]
            if common:
                gs = g_score[common]
                rs = np.array([r_key[int(g_cand[i])] for i in common], dtype=np.float32)
                ulp = _ulp_diff(gs, rs)
                rep["score_bits_differ"] = int((ulp != 0).sum())
                rep["score_max_ulp"] = int(ulp.max())
                rel = np.abs(gs.astype(np.float64) - rs) / np.maximum(np.abs(rs.astype(np.float64)), 1e-30)
                rep["score_max_rel_err"] = float(rel.max())
            else:
                rep["score_bits_differ"] = 0
        if only_ref or only_gpu:
            rep["stage"] = "filter"
    # ---- kept rows ------------------------------------------------------------------------------------------------
    if (ref_count < 0) != (gpu_count < 0):
        rep["stage"] = rep["stage"] if rep["stage"] != "ok" else "nms"
        rep["rows_bit_exact"] = False
        return rep
    if ref_count < 0:
        rep["rows_bit_exact"] = True
        return rep
    r_idx = np.asarray(want.cand_index, dtype=np.int64) if rep["oracle_reproduces_reference"] else None
    same_shape = gpu_rows.shape == ref_rows.shape
    rep["rows_bit_exact"] = bool(same_shape and np.array_equal(gpu_rows, ref_rows))
    if r_idx is not None:
        rep["kept_indices_equal"] = bool(same_shape and np.array_equal(np.asarray(gpu_idx, dtype=np.int64), r_idx))
    if same_shape and ref_rows.size:
        # north-star tolerance |a-b| <= tol * max(|b|, 1), with the box corners measured against the UN-CANCELLED operands
        # of xywh->xyxy: x1 = cx - w/2 inherits 1e-5 * (|cx| + w/2) = 1e-5 * max(|x1|, |x2|) from decoded cx, w that are
        # each within 1e-5 relative (a 600-px-wide box whose x1 is 3 px cannot be held to 1e-5 * 3 px)
        ref64 = ref_rows.astype(np.float64)
        den = np.maximum(np.abs(ref64), 1.0)
        den[:, [0, 2]] = np.maximum(den[:, [0]], den[:, [2]])
        den[:, [1, 3]] = np.maximum(den[:, [1]], den[:, [3]])
        rep["rows_max_rel_err"] = float((np.abs(gpu_rows.astype(np.float64) - ref64) / den).max())
    if rep["rows_bit_exact"]:
        pass
    elif r_idx is not None and rep.get("kept_indices_equal"):
        rep["stage"] = "rows" if rep["stage"] == "ok" else rep["stage"]
    elif rep["stage"] == "ok":
        rep["stage"] = "order" if rep.get("score_bits_differ") else "nms"
    # ---- IoU values within near_eps of the threshold among the pairs the greedy loop looked at ----------------------
    if len(want.keep) and len(want.nms_scores):
        order = np.lexsort((np.arange(len(want.nms_scores)), -want.nms_scores.astype(np.float64)))[:4096]
        iou = cnms.numba_iou(want.nms_boxes[np.asarray(want.keep, dtype=np.int64)], want.nms_boxes[order])
        rep["iou_within_eps_of_threshold"] = int(np.sum(np.abs(iou - float(hyp["iou_threshold"])) < near_eps))
    return rep


def summarize(reports):
    """Per-image reports -> one dict of counters (what bench.py and the tests print)."""
    keys = ("survivors_only_in_reference", "survivors_only_in_engine", "score_bits_differ", "iou_within_eps_of_threshold")
    out = {"images": len(reports),
           "images_rows_bit_exact": sum(1 for r in reports if r.get("rows_bit_exact")),
           "images_kept_indices_equal": sum(1 for r in reports if r.get("kept_indices_equal", r.get("rows_bit_exact"))),
           "oracle_reproduces_reference": all(r.get("oracle_reproduces_reference", True) for r in reports),
           "stages": {}}
    for r in reports:
        out["stages"][r["stage"]] = out["stages"].get(r["stage"], 0) + 1
    for k in keys:
        out[k] = int(sum(r.get(k, 0) for r in reports))
    out["score_max_ulp"] = int(max([r.get("score_max_ulp", 0) for r in reports] + [0]))
    out["rows_max_rel_err"] = float(max([r.get("rows_max_rel_err", 0.0) for r in reports] + [0.0]))
    return out
