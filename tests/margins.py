"""Margin-aware end-to-end decision parity (SURVEY.md hard parts 1 and 3).

The GPU path computes the conv stack in fp16 (or fp32 with another summation order), the oracle in fp32,
so the candidate rows differ by a small measured amount.  Discrete decisions (threshold pass, background
filter, visiting order, IoU suppression) may legitimately flip where the ORACLE's own margin is smaller
than that row error -- and nowhere else.  For one image this module

  1. pushes both row sets through the oracle's score / box restatement (oracle/tails.py) and MEASURES the
     bounds: ds = max |score_gpu - score_oracle|, di = max |IoU_gpu - IoU_oracle| over the pairs that matter;
  2. classifies every class as CLEAN (no threshold / background / ordering / IoU margin of the oracle within
     the measured bound among the candidates greedy NMS can visit) or AMBIGUOUS (reasons listed);
  3. requires the GPU's keep list (candidate row indices, in order) to be IDENTICAL to the oracle's for every
     clean class, and ADMISSIBLE for every ambiguous class: each kept box is a candidate within ds of the
     threshold rule, the order is descending within 2 ds, no kept pair overlaps by more than iou_thr + di, and
     every confident oracle candidate that is missing is either beyond the max_boxes cut or suppressed (within
     di) by a kept box that precedes it (within 2 ds);
  4. reports the counts, the measured bounds and the worst box error of the identical keeps.
"""
import numpy as np

from oracle import tails as OT

F32 = np.float32


def iou_matrix(a, b):
    """TF NonMaxSuppressionV3 IoU of boxes a [n,4] x b [m,4], float32."""
    a, b = np.asarray(a, F32), np.asarray(b, F32)
    ay1, ax1 = np.minimum(a[:, 0], a[:, 2]), np.minimum(a[:, 1], a[:, 3])
    ay2, ax2 = np.maximum(a[:, 0], a[:, 2]), np.maximum(a[:, 1], a[:, 3])
    by1, bx1 = np.minimum(b[:, 0], b[:, 2]), np.minimum(b[:, 1], b[:, 3])
    by2, bx2 = np.maximum(b[:, 0], b[:, 2]), np.maximum(b[:, 1], b[:, 3])
    area_a = ((ay2 - ay1) * (ax2 - ax1)).astype(F32)
    area_b = ((by2 - by1) * (bx2 - bx1)).astype(F32)
    h = np.maximum(np.minimum(ay2[:, None], by2[None]) - np.maximum(ay1[:, None], by1[None]), F32(0))
    w = np.maximum(np.minimum(ax2[:, None], bx2[None]) - np.maximum(ax1[:, None], bx1[None]), F32(0))
    inter = (h * w).astype(F32)
    with np.errstate(divide="ignore", invalid="ignore"):
        iou = inter / ((area_a[:, None] + area_b[None]).astype(F32) - inter)
    iou[(area_a <= 0)[:, None] | (area_b <= 0)[None]] = 0
    return np.nan_to_num(iou.astype(F32), nan=0.0)


def per_row(kind, rows, geom):
    """rows [N,25] of ONE image -> (scores [N,C], boxes [N,4], valid [N] bool, bg_margin [N] or None).
    kind: 'ssd' (geom: size, shapes) | 'retina' (data_shape, shapes) | 'yolo' (levels, priors) | 'fcos' (levels)."""
    rows = np.asarray(rows, F32)
    if kind in ("ssd", "retina"):
        if kind == "ssd":
            _, _, ayx, ahw = OT.ssd_anchors(geom["size"], geom["shapes"])
        else:
            _, _, ayx, ahw = OT.retina_anchors(geom["data_shape"], geom["shapes"])
        prob, boxes, valid = OT.softmax_scores_boxes(rows[:, :21], rows[:, 21:23], rows[:, 23:], ayx, ahw, 20)
        margin = prob[:, :20].max(axis=1) - prob[:, 20]   # >= 0 <=> a foreground class is the (first) arg-max
        return prob[:, :20], boxes, valid, margin
    if kind == "yolo":
        preds = OT.rows_to_levels(rows[None], geom["levels"], "yolo")
        conf, boxes = OT.yolo_scores_boxes(preds, geom["priors"])
    else:
        heads = OT.rows_to_levels(rows[None], geom["levels"], "fcos")
        conf, boxes = OT.fcos_scores_boxes(heads)
    return conf, boxes, np.ones(conf.shape[0], bool), None


def oracle_keep(scores, boxes, valid, n_classes, thr, max_boxes, iou_thr):
    """class -> ordered keep indices (candidate rows), via the oracle's per-class NMS loop."""
    idx = np.nonzero(valid)[0].astype(np.int32)
    _, _, cid, keep = OT._per_class_nms(scores[valid], boxes[valid], idx, n_classes, thr, max_boxes, iou_thr)
    return {c: keep[cid == c].tolist() for c in range(n_classes)}


def compare_image(kind, rows_gpu, rows_ref, geom, gpu_ids, gpu_keep, gpu_boxes, thr, max_boxes, iou_thr,
                  n_classes=20, slack=1.25):
    """Asserts margin-aware identity of the GPU's decisions for one image; returns the report dict."""
    thr, iou_thr = F32(thr), F32(iou_thr)
    S_o, B_o, V_o, M_o = per_row(kind, rows_ref, geom)
    S_g, B_g, V_g, _ = per_row(kind, rows_gpu, geom)
    finite = np.isfinite(S_g).all() and np.isfinite(S_o).all()
    assert finite, "non-finite scores"
    ds = float(np.abs(S_g - S_o).max()) * slack
    if M_o is not None:
        Mg = per_row(kind, rows_gpu, geom)[3]
        ds = max(ds, float(np.abs(Mg - M_o).max()) * slack / 2)
    keep_o = oracle_keep(S_o, B_o, V_o, n_classes, thr, max_boxes, iou_thr)
    gpu_ids, gpu_keep = np.asarray(gpu_ids), np.asarray(gpu_keep)
    keep_g = {c: gpu_keep[gpu_ids == c].tolist() for c in range(n_classes)}
    assert set(np.unique(gpu_ids).tolist()) <= set(range(n_classes))
    rep = {"ds": ds, "di": 0.0, "clean": 0, "ambiguous": 0, "identical": 0, "reasons": {}, "box_err": 0.0,
           "kept_oracle": sum(len(v) for v in keep_o.values()), "kept_gpu": int(len(gpu_keep)), "notes": []}
    for c in range(n_classes):
        ko, kg = keep_o[c], keep_g[c]
        s = S_o[:, c]
        full = len(ko) == max_boxes
        s_cut = (s[ko[-1]] - 2 * ds) if full else (thr - ds)
        maybe_fg = V_o if M_o is None else (M_o >= -2 * ds)
        R = np.nonzero((s >= s_cut) & maybe_fg)[0]
        reasons = []
        if R.size and np.any(np.abs(s[R] - thr) <= ds):
            reasons.append("threshold")
        if M_o is not None and R.size and np.any(np.abs(M_o[R]) <= 2 * ds):
            reasons.append("background")
        if R.size > 1:
            srt = np.sort(s[R])
            if np.any(np.diff(srt) <= 2 * ds):
                reasons.append("order")
        K = sorted(set(ko) | set(kg))
        di = 0.0
        if K and R.size:
            io = iou_matrix(B_o[K], B_o[R])
            ig = iou_matrix(B_g[K], B_g[R])
            di = float(np.abs(io - ig).max()) * slack
            if np.any(np.abs(io[[K.index(k) for k in ko]] - iou_thr) <= di):
                reasons.append("iou")
        rep["di"] = max(rep["di"], di)
        if kg == ko:
            rep["identical"] += 1
        if not reasons:
            rep["clean"] += 1
            assert kg == ko, ("class %d is clean (margins > ds=%.3g, di=%.3g) but keep lists differ: gpu %s oracle %s"
                              % (c, ds, di, kg[:12], ko[:12]))
            continue
        rep["ambiguous"] += 1
        for r in reasons:
            rep["reasons"][r] = rep["reasons"].get(r, 0) + 1
        # admissibility of the GPU's list under the measured bounds
        kg_a = np.asarray(kg, np.int64)
        if kg:
            assert np.all(s[kg_a] >= thr - ds), ("class %d: kept box below threshold - ds" % c, s[kg_a], thr, ds)
            assert np.all(maybe_fg[kg_a]), "class %d: kept a confidently-background row" % c
            assert np.all(s[kg_a][:-1] >= s[kg_a][1:] - 2 * ds), "class %d: keep order beyond 2 ds" % c
            assert len(set(kg)) == len(kg) and len(kg) <= max_boxes
            if len(kg) > 1:
                iog = iou_matrix(B_o[kg_a], B_o[kg_a])
                iu = np.triu(iog, 1)
                assert np.all(iu <= iou_thr + di + 1e-7), "class %d: two kept boxes overlap beyond iou_thr + di" % c
        conf = R[(s[R] >= thr + ds) & (V_o[R] if M_o is None else (M_o[R] > 2 * ds))]
        miss = [m for m in conf.tolist() if m not in set(kg)]
        if miss:
            if kg:
                iom = iou_matrix(B_o[kg_a], B_o[np.asarray(miss)])
            for q, m in enumerate(miss):
                if len(kg) == max_boxes and s[m] <= s[kg[-1]] + 2 * ds:
                    continue  # beyond the max_boxes cut
                ok = bool(kg) and bool(np.any((s[kg_a] >= s[m] - 2 * ds) & (iom[:, q] >= iou_thr - di - 1e-7)))
                assert ok, ("class %d: confident oracle candidate %d (score %.6f) is neither kept nor suppressed "
                            "within the measured bounds" % (c, m, s[m]))
    # boxes of the keeps both sides agree on (position by position)
    same = [n for c in range(n_classes) for n in keep_g[c] if n in set(keep_o[c])]
    if len(gpu_keep):
        pos = {(int(c), int(n)): i for i, (c, n) in enumerate(zip(gpu_ids, gpu_keep))}
        errs, rel = [], []
        for c in range(n_classes):
            for n in keep_g[c]:
                if n in set(keep_o[c]):
                    e = np.abs(np.asarray(gpu_boxes)[pos[(c, n)]] - B_o[n])
                    errs.append(float(e.max()))
                    rel.append(float(e.max() / max(1.0, float(np.abs(B_o[n]).max()))))
        if errs:
            rep["box_err"], rep["box_err_rel"] = max(errs), max(rel)
    rep["same_keeps"] = len(same)
    return rep


def summarize(reps):
    tot = {k: sum(r[k] for r in reps) for k in ("clean", "ambiguous", "identical", "kept_oracle", "kept_gpu", "same_keeps")}
    tot["ds"] = max(r["ds"] for r in reps)
    tot["di"] = max(r["di"] for r in reps)
    tot["box_err"] = max(r["box_err"] for r in reps)
    tot["box_err_rel"] = max(r.get("box_err_rel", 0.0) for r in reps)
    reasons = {}
    for r in reps:
        for k, v in r["reasons"].items():
            reasons[k] = reasons.get(k, 0) + v
    tot["reasons"] = reasons
    return tot
