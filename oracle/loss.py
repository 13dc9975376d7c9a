"""RetinaNet per-image training loss, forward only, restated on CPU (test oracle).
Follows RetinaNet.py:357-474 step by step in numpy float32.  Not product code."""
import numpy as np

from . import tfops as T

F32 = np.float32


def _smooth_l1(x):
    ax = np.abs(x)
    return np.where(ax < 1.0, F32(0.5) * x * x, ax - F32(0.5)).astype(F32)


def retina_image_loss(pconf, pyx, phw, a_y1x1, a_y2x2, a_yx, a_hw, gt, alpha=0.25, gamma=2.0):
    """pconf [N,21] logits, pyx/phw [N,2], anchors [N,2] x4, gt [G,5] (y,x,h,w,id) padded -1."""
    gt = np.asarray(gt, F32)
    cnt = int(np.argmin(gt, axis=0)[0])  # :359-360
    g = gt[:cnt]
    gyx, ghw = g[:, 0:2], g[:, 2:4]
    gy1x1 = (gyx - ghw / F32(2.0)).astype(F32)
    gy2x2 = (gyx + ghw / F32(2.0)).astype(F32)
    label = g[:, 4].astype(np.int32)
    # IoU[G,A]  :383-388
    i1 = np.maximum(a_y1x1[None], gy1x1[:, None])
    i2 = np.minimum(a_y2x2[None], gy2x2[:, None])
    inter = np.prod(np.maximum(i2 - i1, F32(0)), axis=-1).astype(F32)
    aarea = np.prod(a_hw, axis=-1).astype(F32)[None]
    garea = np.prod(ghw, axis=-1).astype(F32)[:, None]
    iou = (inter / (aarea + garea - inter)).astype(F32)
    best = np.argmax(iou, axis=1)  # per GT, first maximum  :390
    bestmask = np.zeros(iou.shape[1], bool)
    bestmask[best] = True
    other = ~bestmask
    o_iou = iou.T[other]  # [A', G]
    o_best = o_iou.max(axis=1)
    pos = o_best > F32(0.5)
    neg = o_best < F32(0.4)
    rg = np.argmax(o_iou, axis=1)
    o_conf, o_yx, o_hw = pconf[other], pyx[other], phw[other]
    o_ayx, o_ahw = a_yx[other], a_hw[other]
    pos_conf = np.concatenate([pconf[best], o_conf[pos]], 0)
    pos_label = np.concatenate([label, label[rg[pos]]], 0)
    pos_pyx = np.concatenate([pyx[best], o_yx[pos]], 0)
    pos_phw = np.concatenate([phw[best], o_hw[pos]], 0)
    pos_gyx = np.concatenate([gyx, gyx[rg[pos]]], 0)
    pos_ghw = np.concatenate([ghw, ghw[rg[pos]]], 0)
    pos_ayx = np.concatenate([a_yx[best], o_ayx[pos]], 0)
    pos_ahw = np.concatenate([a_hw[best], o_ahw[pos]], 0)
    neg_conf = o_conf[neg]
    # focal  :457-474 (softmax, same alpha for positives and negatives)
    pp = T.softmax_lastdim(pos_conf)[np.arange(len(pos_label)), pos_label]
    npb = T.softmax_lastdim(neg_conf)[:, pconf.shape[1] - 1] if len(neg_conf) else np.zeros(0, F32)
    pp = np.clip(pp, F32(1e-8), F32(1.0))
    npb = np.clip(npb, F32(1e-8), F32(1.0))
    posloss = -F32(alpha) * np.power(F32(1.0) - pp, F32(gamma)) * np.log(pp)
    negloss = -F32(alpha) * np.power(F32(1.0) - npb, F32(gamma)) * np.log(npb)
    conf_loss = (posloss.sum(dtype=np.float64) + negloss.sum(dtype=np.float64)) / len(posloss)
    # smooth-L1  :445-449
    tyx = ((pos_gyx - pos_ayx) / pos_ahw).astype(F32)
    thw = np.log((pos_ghw / pos_ahw).astype(F32)).astype(F32)
    per = _smooth_l1(pos_pyx - tyx).sum(-1) + _smooth_l1(pos_phw - thw).sum(-1)
    coord = per.mean(dtype=np.float64)
    return float(conf_loss + coord), dict(num_pos=len(posloss), num_neg=len(negloss))


def ssd_image_loss(pconf, pyx, phw, a_y1x1, a_y2x2, a_yx, a_hw, gt):
    """SSD per-image training loss, forward only (SSD300.py:345-453, SSD512.py same body).
    pconf [N,21] logits (background = last), pyx/phw [N,2], anchors [N,2] x4, gt [G,5]
    (y,x,h,w,id) padded with -1.  Matching: every GT's arg-max anchor is positive (duplicates
    kept), the other anchors are positive when their best IoU > 0.5 and NEGATIVE otherwise (no
    ignore band).  Hard-negative mining = tf.image.non_max_suppression over the negative
    ANCHOR boxes scored by their background cross-entropy, at most 3 x #positives, IoU 0.7."""
    from . import tails as OT
    gt = np.asarray(gt, F32)
    cnt = int(np.argmin(gt, axis=0)[0])  # :347-348
    g = gt[:cnt]
    gyx, ghw = g[:, 0:2], g[:, 2:4]
    gy1x1 = (gyx - ghw / F32(2.0)).astype(F32)
    gy2x2 = (gyx + ghw / F32(2.0)).astype(F32)
    label = g[:, 4].astype(np.int32)
    i1 = np.maximum(a_y1x1[None], gy1x1[:, None])
    i2 = np.minimum(a_y2x2[None], gy2x2[:, None])
    inter = np.prod(np.maximum(i2 - i1, F32(0)), axis=-1).astype(F32)
    aarea = np.prod(a_hw, axis=-1).astype(F32)[None]
    garea = np.prod(ghw, axis=-1).astype(F32)[:, None]
    iou = (inter / (aarea + garea - inter)).astype(F32)        # :372-377
    best = np.argmax(iou, axis=1)                                # :379
    bestmask = np.zeros(iou.shape[1], bool)
    bestmask[best] = True
    other = ~bestmask
    o_iou = iou.T[other]
    o_best = o_iou.max(axis=1)
    pos = o_best > F32(0.5)                                      # :404-405
    neg = ~pos
    rg = np.argmax(o_iou, axis=1)
    o_conf, o_yx, o_hw = pconf[other], pyx[other], phw[other]
    o_ayx, o_ahw = a_yx[other], a_hw[other]
    neg_conf = o_conf[neg]
    neg_ayx, neg_ahw = o_ayx[neg], o_ahw[neg]
    neg_box = np.concatenate([neg_ayx - neg_ahw / F32(2.0), neg_ayx + neg_ahw / F32(2.0)], -1).astype(F32)
    num_pos = cnt + int(pos.sum())
    num_neg = int(neg.sum())
    chosen = 3 * num_pos if num_neg > 3 * num_pos else num_neg   # :426

    def xent(logits, labels):  # sparse_softmax_cross_entropy: logsumexp(x) - x[label]
        m = logits.max(axis=1, keepdims=True)
        z = (logits - m).astype(F32)
        lse = np.log(np.exp(z).astype(F32).sum(axis=1, dtype=F32)).astype(F32)
        return (lse - z[np.arange(len(labels)), labels]).astype(F32)

    bg = pconf.shape[1] - 1
    neg_l = xent(neg_conf, np.full(num_neg, bg, np.int64))
    sel = OT.nms_c(neg_box, neg_l, chosen, 0.7)                  # :431-433
    neg_loss = neg_l[sel].mean(dtype=np.float64)
    pos_conf = np.concatenate([pconf[best], o_conf[pos]], 0)
    pos_label = np.concatenate([label, label[rg[pos]]], 0)
    pos_pyx = np.concatenate([pyx[best], o_yx[pos]], 0)
    pos_phw = np.concatenate([phw[best], o_hw[pos]], 0)
    pos_gyx = np.concatenate([gyx, gyx[rg[pos]]], 0)
    pos_ghw = np.concatenate([ghw, ghw[rg[pos]]], 0)
    pos_ayx = np.concatenate([a_yx[best], o_ayx[pos]], 0)
    pos_ahw = np.concatenate([a_hw[best], o_ahw[pos]], 0)
    pos_conf_loss = xent(pos_conf, pos_label).mean(dtype=np.float64)
    tyx = ((pos_gyx - pos_ayx) / pos_ahw).astype(F32)
    thw = np.log((pos_ghw / pos_ahw).astype(F32)).astype(F32)
    per = _smooth_l1(pos_pyx - tyx).sum(-1) + _smooth_l1(pos_phw - thw).sum(-1)
    coord = per.mean(dtype=np.float64)
    return float(neg_loss + pos_conf_loss + coord), dict(num_pos=num_pos, num_neg=num_neg, selected=len(sel))
