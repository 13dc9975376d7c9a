"""Inference tails of the five detectors restated on CPU (test oracle):
anchor / prior generation, score activation, box decode, thresholding and the
per-class NMS loop.  numpy float32 in the reference's op order.  Not product
code -- see oracle/__init__.py.
"""
import ctypes
import os
import subprocess

import numpy as np

from . import tfops as T

F32 = np.float32
_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def build_nms_lib(force=False):
    """Compile oracle/nms_ref.c -> oracle/libodt_oracle.so (gcc, no FMA contraction)."""
    so = os.path.join(_HERE, "libodt_oracle.so")
    src = os.path.join(_HERE, "nms_ref.c")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["gcc", "-O2", "-ffp-contract=off", "-shared", "-fPIC", src, "-o", so])
    return so


def _lib():
    global _LIB
    if _LIB is None:
        _LIB = ctypes.CDLL(build_nms_lib())
        _LIB.odt_oracle_nms.restype = ctypes.c_int
        _LIB.odt_oracle_nms.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_int,
                                        ctypes.c_float, ctypes.c_void_p]
    return _LIB


def nms_c(boxes, scores, max_out, iou_thr):
    boxes = np.ascontiguousarray(boxes, dtype=F32)
    scores = np.ascontiguousarray(scores, dtype=F32)
    n = int(scores.shape[0])
    sel = np.zeros(max(max_out, 1), dtype=np.int32)
    k = _lib().odt_oracle_nms(boxes.ctypes.data, scores.ctypes.data, n, int(max_out),
                              float(iou_thr), sel.ctypes.data)
    return sel[:k].copy()


def _iou_py(a, b):
    f = F32
    ymin_i, xmin_i = min(a[0], a[2]), min(a[1], a[3])
    ymax_i, xmax_i = max(a[0], a[2]), max(a[1], a[3])
    ymin_j, xmin_j = min(b[0], b[2]), min(b[1], b[3])
    ymax_j, xmax_j = max(b[0], b[2]), max(b[1], b[3])
    area_i = f(f(ymax_i - ymin_i) * f(xmax_i - xmin_i))
    area_j = f(f(ymax_j - ymin_j) * f(xmax_j - xmin_j))
    if area_i <= 0 or area_j <= 0:
        return f(0)
    h = max(f(min(ymax_i, ymax_j) - max(ymin_i, ymin_j)), f(0))
    w = max(f(min(xmax_i, xmax_j) - max(xmin_i, xmin_j)), f(0))
    inter = f(h * w)
    return f(inter / f(f(area_i + area_j) - inter))


def nms_py(boxes, scores, max_out, iou_thr):
    """Pure-Python statement of App. A.8 for small cases (validates nms_ref.c)."""
    boxes = np.asarray(boxes, dtype=F32)
    scores = np.asarray(scores, dtype=F32)
    order = sorted(range(len(scores)), key=lambda i: (-float(scores[i]), i))
    sel = []
    for i in order:
        if len(sel) >= max_out:
            break
        keep = True
        for j in reversed(sel):
            if _iou_py(boxes[i], boxes[j]) > F32(iou_thr):
                keep = False
                break
        if keep:
            sel.append(i)
    return np.array(sel, dtype=np.int32)


# ------------------------------------------------------------ anchors -------
def ssd_scales(size):
    """SSD300.py:112-113 / SSD512.py:116-118 (Python doubles, as the reference)."""
    if size == 300:
        s = [(0.2 + (0.9 - 0.2) / 5 * (i - 1)) * size for i in range(1, 8)]
        return [[s[i], (s[i] * s[i + 1]) ** 0.5] for i in range(0, 6)]
    s = [0.07 * size]
    s = s + [(0.15 + (0.9 - 0.15) / 5 * (i - 1)) * size for i in range(1, 8)]
    return [[s[i], (s[i] * s[i + 1]) ** 0.5] for i in range(0, 7)]


def ssd_aspect_ratios(size):
    """SSD300.py:114-119 / SSD512.py:119-125."""
    if size == 300:
        return [[2, 1 / 2], [2, 1 / 2, 3, 1 / 3], [2, 1 / 2, 3, 1 / 3], [2, 1 / 2, 3, 1 / 3],
                [2, 1 / 2], [2, 1 / 2]]
    return [[2, 1 / 2], [2, 1 / 2, 3, 1 / 3], [2, 1 / 2, 3, 1 / 3], [2, 1 / 2, 3, 1 / 3],
            [2, 1 / 2, 3, 1 / 3], [2, 1 / 2], [2, 1 / 2]]


def ssd_priors(size_pair, ars):
    """SSD300.py:333-336: [[s0,s0],[s1,s1]] + [[s0*sqrt(ar), s0/sqrt(ar)]] as (h,w)."""
    pri = [[size_pair[0], size_pair[0]], [size_pair[1], size_pair[1]]]
    for ar in ars:
        pri.append([size_pair[0] * (ar ** 0.5), size_pair[0] / (ar ** 0.5)])
    return np.array(pri, dtype=F32)


def _grid_anchors(h, w, cy, cx, priors):
    """SSD300.py:330-343 / RetinaNet.py:341-355: corners first, yx and hw re-derived."""
    a = priors.shape[0]
    yx = np.zeros((h, w, a, 2), dtype=F32)
    yx[..., 0] = cy.reshape(h, 1, 1)
    yx[..., 1] = cx.reshape(1, w, 1)
    pr = priors.reshape(1, 1, a, 2)
    y1x1 = (yx - pr / F32(2.0)).astype(F32).reshape(-1, 2)
    y2x2 = (yx + pr / F32(2.0)).astype(F32).reshape(-1, 2)
    ayx = (y1x1 / F32(2.0) + y2x2 / F32(2.0)).astype(F32)
    ahw = (y2x2 - y1x1).astype(F32)
    return y1x1, y2x2, ayx, ahw


def ssd_anchors(size, shapes):
    """All levels concatenated.  shapes: [(H,W)] of the pred tensors."""
    sc, ars = ssd_scales(size), ssd_aspect_ratios(size)
    parts = []
    for (h, w), s, ar in zip(shapes, sc, ars):
        cy = ((np.arange(h, dtype=F32) + F32(0.5)) * F32(size) / F32(h)).astype(F32)  # :328
        cx = ((np.arange(w, dtype=F32) + F32(0.5)) * F32(size) / F32(w)).astype(F32)
        parts.append(_grid_anchors(h, w, cy, cx, ssd_priors(s, ar)))
    return [np.concatenate([p[i] for p in parts], axis=0) for i in range(4)]


def retina_priors(size):
    """RetinaNet.py:344-348, ratio-major [1,1/2,2] x scale [2^0,2^(1/3),2^(2/3)]."""
    pri = []
    for r in [1, 1 / 2, 2]:
        for s in [2 ** 0, 2 ** (1 / 3), 2 ** (2 / 3)]:
            pri.append([s * size * (r ** 0.5), s * size / (r ** 0.5)])
    return np.array(pri, dtype=F32)


def retina_anchors(data_shape, shapes):
    """RetinaNet.py:328-355.  NB: the stride uses data_shape[1] (= W) for both axes."""
    parts = []
    for (h, w), size in zip(shapes, [32, 64, 128, 256, 512]):
        rate = F32(F32(data_shape[1]) / F32(h))  # :330-331
        cy = ((np.arange(h, dtype=F32) + F32(0.5)) * rate).astype(F32)
        cx = ((np.arange(w, dtype=F32) + F32(0.5)) * rate).astype(F32)
        parts.append(_grid_anchors(h, w, cy, cx, retina_priors(size)))
    return [np.concatenate([p[i] for p in parts], axis=0) for i in range(4)]


# -------------------------------------------------------------- NMS loop ----
def _per_class_nms(conf, boxes, anchor_idx, n_classes, score_thr, max_boxes, iou_thr):
    """The Python-unrolled per-class loop (SSD300.py:172-190 and siblings).
    Returns scores, boxes, class ids and, for tests, the candidate row index of
    every kept box (keep index)."""
    fmask = conf >= F32(score_thr)
    out_s, out_b, out_c, out_k = [], [], [], []
    for i in range(n_classes):
        m = fmask[:, i]
        si, bi, ki = conf[m, i], boxes[m], anchor_idx[m]
        sel = nms_c(bi, si, max_boxes, iou_thr)
        out_s.append(si[sel])
        out_b.append(bi[sel])
        out_c.append(np.full(len(sel), i, dtype=np.int32))
        out_k.append(ki[sel])
    return (np.concatenate(out_s).astype(F32), np.concatenate(out_b).astype(F32).reshape(-1, 4),
            np.concatenate(out_c), np.concatenate(out_k).astype(np.int32))


def softmax_scores_boxes(pconf, pyx, phw, ayx, ahw, num_fg):
    """Per-row quantities of SSD300.py:157-171 / RetinaNet.py:224-238 for ONE image, before any filtering:
    softmax probabilities [N,num_fg+1], boxes [N,4] (y1,x1,y2,x2) and the background filter (arg-max is a
    foreground class; background is the LAST class, first maximum wins)."""
    prob = T.softmax_lastdim(pconf)
    valid = np.argmax(prob, axis=-1) < num_fg
    dyx = (pyx * ahw).astype(F32)
    dyx = (dyx + ayx).astype(F32)
    dhw = (ahw * np.exp(phw.astype(F32)).astype(F32)).astype(F32)
    y1x1 = (dyx - dhw / F32(2.0)).astype(F32)
    y2x2 = (dyx + dhw / F32(2.0)).astype(F32)
    return prob, np.concatenate([y1x1, y2x2], axis=-1), valid


def softmax_tail(pconf, pyx, phw, ayx, ahw, num_fg, score_thr, max_boxes, iou_thr):
    """SSD300.py:157-190 / RetinaNet.py:224-256 for ONE image.
    pconf [N,21], pyx/phw [N,2], anchors [N,2]."""
    prob, boxes, keep = softmax_scores_boxes(pconf, pyx, phw, ayx, ahw, num_fg)
    idx = np.nonzero(keep)[0].astype(np.int32)
    return _per_class_nms(prob[keep][:, :num_fg], boxes[keep], idx, num_fg, score_thr, max_boxes, iou_thr)


def ssd_rows(preds, num_classes=21):
    """_get_pbbox (SSD300.py:316-321) + concat (:121-123): [B, N, C+4] rows."""
    b = preds[0].shape[0]
    return np.concatenate([p.reshape(b, -1, num_classes + 4) for p in preds], axis=1)


def ssd_detect(preds, size, score_thr, max_boxes, iou_thr, image=0, num_classes=21):
    rows = ssd_rows(preds, num_classes)[image]
    shapes = [(p.shape[1], p.shape[2]) for p in preds]
    _, _, ayx, ahw = ssd_anchors(size, shapes)
    return softmax_tail(rows[:, :num_classes], rows[:, num_classes:num_classes + 2],
                        rows[:, num_classes + 2:], ayx, ahw, num_classes - 1, score_thr, max_boxes,
                        iou_thr)


def retina_rows(heads, num_classes=21):
    """_get_pbbox (RetinaNet.py:321-326): rows [B,N,25] = 21 logits + (ty,tx,th,tw)."""
    b = heads[0][0].shape[0]
    cls = np.concatenate([c.reshape(b, -1, num_classes) for c, _ in heads], axis=1)
    reg = np.concatenate([r.reshape(b, -1, 4) for _, r in heads], axis=1)
    return np.concatenate([cls, reg], axis=2)


def retina_detect(heads, data_shape, score_thr, max_boxes, iou_thr, image=0, num_classes=21):
    rows = retina_rows(heads, num_classes)[image]
    shapes = [(c.shape[1], c.shape[2]) for c, _ in heads]
    _, _, ayx, ahw = retina_anchors(data_shape, shapes)
    return softmax_tail(rows[:, :num_classes], rows[:, num_classes:num_classes + 2],
                        rows[:, num_classes + 2:], ayx, ahw, num_classes - 1, score_thr, max_boxes,
                        iou_thr)


def yolo_rows(preds, num_classes=20, num_priors=3):
    b = preds[0].shape[0]
    return np.concatenate([p.reshape(b, -1, num_classes + 5) for p in preds], axis=1)


def yolo_scores_boxes(preds, priors, image=0, num_classes=20):
    """Per-row confidences [N,20] and boxes [N,4] of YOLOv3.py:320-351 for one image (no filtering).
    priors: config['priors'] (3x3x2 nested list).  Level k uses priors[k]/stride[k] with stride=[8,16,32] in
    *config order* (YOLOv3.py:38-41) and output multipliers 32,32,16 (:346-348)."""
    stride = [8.0, 16.0, 32.0]
    mult = [stride[-1], stride[-1], stride[-2]]
    confs, boxes = [], []
    for k, p in enumerate(preds):
        _, h, w, _ = p.shape
        r = p[image].reshape(h, w, 3, num_classes + 5).astype(F32)
        pri = (np.array(priors[k], dtype=F32) / F32(stride[k])).astype(F32).reshape(1, 1, 3, 2)
        cy = (np.arange(h, dtype=F32) + F32(0.5)).reshape(h, 1, 1)
        cx = (np.arange(w, dtype=F32) + F32(0.5)).reshape(1, w, 1)
        ayx = np.zeros((h, w, 3, 2), dtype=F32)
        ayx[..., 0] = cy
        ayx[..., 1] = cx
        cls = T.sigmoid(r[..., :num_classes])
        obj = T.sigmoid(r[..., num_classes + 4:])
        byx = (ayx + T.sigmoid(r[..., num_classes:num_classes + 2])).astype(F32)
        bhw = (pri + np.exp(r[..., num_classes + 2:num_classes + 4]).astype(F32)).astype(F32)
        box = np.concatenate([byx - bhw / F32(2.0), byx + bhw / F32(2.0)], axis=-1).astype(F32)
        boxes.append((box * F32(mult[k])).astype(F32).reshape(-1, 4))
        confs.append((cls * obj).astype(F32).reshape(-1, num_classes))
    return np.concatenate(confs, axis=0), np.concatenate(boxes, axis=0)


def yolo_detect(preds, priors, score_thr, max_boxes, iou_thr, image=0, num_classes=20):
    """YOLOv3.py:320-368 for one image."""
    conf, box = yolo_scores_boxes(preds, priors, image, num_classes)
    idx = np.arange(conf.shape[0], dtype=np.int32)
    return _per_class_nms(conf, box, idx, num_classes, score_thr, max_boxes, iou_thr)


def fcos_rows(heads):
    """Our 25-float row layout for FCOS: [cls(20), ctr, l, r, t, b] (raw, pre-exp)."""
    b = heads[0][0].shape[0]
    return np.concatenate(
        [np.concatenate([c.reshape(b, -1, c.shape[3]), ct.reshape(b, -1, 1), rg.reshape(b, -1, 4)],
                        axis=2) for c, ct, rg in heads], axis=1)


def fcos_scores_boxes(heads, image=0, num_classes=20):
    """Per-row confidences [N,20] and boxes [N,4] of FCOS.py:130-150,197-248 for one image (no filtering)."""
    strides = [8, 16, 32, 64, 128]
    confs, boxes = [], []
    for (cls, ctr, reg), s in zip(heads, strides):
        _, h, w, _ = cls.shape
        conf = (T.sigmoid(cls[image]) * T.sigmoid(ctr[image])).astype(F32).reshape(-1, num_classes)
        r = np.exp(reg[image].astype(F32)).astype(F32)  # (l, r, t, b)  :363
        gy = np.arange(h, dtype=F32).reshape(h, 1, 1)
        gx = np.arange(w, dtype=F32).reshape(1, w, 1)
        y1 = gy - r[..., 2:3]
        y2 = gy + r[..., 3:4]
        x1 = gx - r[..., 0:1]
        x2 = gx + r[..., 1:2]
        box = (np.concatenate([y1, x1, y2, x2], axis=-1).astype(F32).reshape(-1, 4) * F32(s)).astype(F32)
        confs.append(conf)
        boxes.append(box)
    return np.concatenate(confs, axis=0), np.concatenate(boxes, axis=0)


def fcos_detect(heads, score_thr, max_boxes, iou_thr, image=0, num_classes=20):
    """FCOS.py:130-150,197-264 for one image.  NMS visits num_classes-1 classes (:252)."""
    conf, box = fcos_scores_boxes(heads, image, num_classes)
    idx = np.arange(conf.shape[0], dtype=np.int32)
    return _per_class_nms(conf, box, idx, num_classes - 1, score_thr, max_boxes, iou_thr)


def rows_to_levels(rows, levels, kind):
    """Inverse of *_rows: candidate rows [B,N,25] -> the per-level tensors the *_detect functions take
    (`levels`: [(H,W,A)] in candidate order).  kind: 'yolo' -> preds [B,H,W,A*25]; 'fcos' -> (cls, ctr, reg)."""
    b, off, out = rows.shape[0], 0, []
    for h, w, a in levels:
        r = rows[:, off:off + h * w * a]
        if kind == "yolo":
            out.append(r.reshape(b, h, w, a * 25))
        else:
            r = r.reshape(b, h, w, 25)
            out.append((r[..., :20], r[..., 20:21], r[..., 21:25]))
        off += h * w * a
    assert off == rows.shape[1]
    return out
