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


def _log_sigmoid(x):
    # tf.log_sigmoid(x) = -softplus(-x)
    return (-(np.maximum(-x, F32(0)) + np.log1p(np.exp(-np.abs(x))))).astype(F32)


def fcos_level_loss(cls, reg_raw, ctr, gt, stride):
    """One pyramid level of one image (FCOS.py:266-348).  cls [H,W,20] logits, reg_raw [H,W,4]
    pre-exp (l,r,t,b) (the head applies exp, :363), ctr [H,W] logit, gt [g,5] rows of THIS level."""
    h, w, nc = cls.shape
    s = F32(stride)
    gy, gx, gh, gw = [(gt[:, i] / s).astype(F32) for i in range(4)]
    cid = gt[:, 4].astype(np.int32)
    y1, y2 = (gy - gh / F32(2.0)).astype(F32), (gy + gh / F32(2.0)).astype(F32)
    x1, x2 = (gx - gw / F32(2.0)).astype(F32), (gx + gw / F32(2.0)).astype(F32)
    yy = np.arange(h, dtype=F32).reshape(h, 1, 1)
    xx = np.arange(w, dtype=F32).reshape(1, w, 1)
    dl, dr = (xx - x1).astype(F32) + np.zeros((h, 1, 1), F32), (x2 - xx).astype(F32) + np.zeros((h, 1, 1), F32)
    dt, db = (yy - y1).astype(F32) + np.zeros((1, w, 1), F32), (y2 - yy).astype(F32) + np.zeros((1, w, 1), F32)
    heat = ((dt > 0) & (db > 0) & (dl > 0) & (dr > 0)).astype(F32)       # :288-290
    dl, dr, dt, db = dl * heat, dr * heat, dt * heat, db * heat
    loc = heat.max(axis=-1)
    area = ((dl + dr) * (dt + db)).astype(F32)
    area_ = (area + (F32(1.0) - heat) * F32(1e8)).astype(F32)
    amin = area_.min(axis=-1, keepdims=True)
    dmask = (area == amin).astype(F32) * loc[..., None]                   # :299
    dl, dr, dt, db = [(d * dmask).max(axis=-1) for d in (dl, dr, dt, db)]
    p = np.exp(reg_raw.astype(F32)).astype(F32)
    pl, pr, pt, pb = p[..., 0], p[..., 1], p[..., 2], p[..., 3]
    iw = np.minimum(dl, pl) + np.minimum(dr, pr)
    ih = np.minimum(dt, pt) + np.minimum(db, pb)
    inter = (iw * ih).astype(F32)
    union = ((dl + dr) * (dt + db) + (pl + pr) * (pt + pb) - inter).astype(F32)
    iou = (inter / (union + F32(1e-12))).astype(F32)
    iou_loss = (-np.log(iou + F32(1e-12)) * loc).sum(dtype=np.float64)
    lrmin, tbmin = np.minimum(dl, dr), np.minimum(dt, db)
    lrmax, tbmax = np.maximum(dl, dr), np.maximum(dt, db)
    cgt = np.sqrt((lrmin * tbmin / (lrmax * tbmax + F32(1e-12))).astype(F32)).astype(F32)
    x = ctr.astype(F32)
    bce = (np.maximum(x, F32(0)) - x * cgt + np.log1p(np.exp(-np.abs(x)))).astype(F32)   # from_logits=True
    center_loss = bce.sum(dtype=np.float64)
    hgt = np.zeros((h, w, nc), F32)
    for c in range(nc):
        m = cid == c
        if m.any():
            hgt[..., c] = heat[..., m].max(axis=-1)
    sg = T.sigmoid(cls.astype(F32))
    ls = _log_sigmoid(cls.astype(F32))
    pos = (F32(-0.25) * np.power(F32(1.0) - sg, F32(2.0)) * ls * hgt).sum(dtype=np.float64)
    neg = (F32(-0.25) * np.power(sg, F32(2.0)) * (-cls.astype(F32) + ls) * (F32(1.0) - hgt)).sum(dtype=np.float64)
    return (iou_loss + pos + neg + center_loss) / hgt.sum(dtype=np.float64)


def fcos_image_loss(heads, gt, image=0):
    """heads: [(cls [B,H,W,20], ctr [B,H,W,1], reg_raw [B,H,W,4])] x 5; gt [G,5] padded with -1.
    GTs go to levels by sqrt(h*w) with the reference's inclusive, overlapping bounds (FCOS.py:158-164)."""
    gt = np.asarray(gt, F32)
    cnt = int(np.argmin(gt, axis=0)[0])
    g = gt[:cnt]
    size = np.sqrt((g[:, 2] * g[:, 3]).astype(F32)).astype(F32)
    sel = [size <= 64, (size >= 64) & (size <= 128), (size >= 128) & (size <= 256),
           (size >= 256) & (size <= 512), size >= 512]
    total = 0.0
    for (cls, ctr, reg), m, s in zip(heads, sel, [8, 16, 32, 64, 128]):
        if m.any():
            total += fcos_level_loss(cls[image], reg[image], ctr[image, ..., 0], g[m], s)
    return float(total)


def _sig_xent(labels, logits):
    # tf.nn.sigmoid_cross_entropy_with_logits: max(x,0) - x*z + log(1 + exp(-|x|))
    x, z = logits.astype(F32), labels.astype(F32)
    return (np.maximum(x, F32(0)) - x * z + np.log1p(np.exp(-np.abs(x)))).astype(F32)


def yolo_image_loss(preds, priors, gt, coord_scale=1.0, noobj_scale=1.0, obj_scale=5.0, class_scale=1.0,
                    image=0, num_classes=20):
    """YOLOv3 per-image training loss, forward only (YOLOv3.py:115-318, priors :37-41,419-433, GT
    normalisation :435-442).  preds: 3 tensors [B,H,W,3*25] (13, 26, 52 grids for 416), rows
    [cls(20), y, x, h, w, obj].  Quirks kept: level k normalises the GT with strides 32,16,8 but its
    priors are priors[k]/(8,16,32)[k]; the GT/anchor intersections are NOT clamped at 0; the
    "no-object" anchors are built from (yx - hw/2, yx + hw/2) re-used as (centre, size)."""
    nc = num_classes
    gt = np.asarray(gt, F32)
    cnt = int(np.argmin(gt, axis=0)[0])
    g = gt[:cnt]
    norm = [32.0, 16.0, 8.0]
    pstride = [8.0, 16.0, 32.0]
    lv = []
    for k, p in enumerate(preds):
        _, h, w, _ = p.shape
        r = p[image].reshape(h, w, 3, nc + 5).astype(F32)
        pri = (np.array(priors[k], dtype=F32) / F32(pstride[k])).astype(F32)      # [3,2] (h,w)
        gn = (g / np.array([norm[k]] * 4 + [1.0], F32).reshape(1, 5)).astype(F32)
        gyx, ghw, lab = gn[:, :2], gn[:, 2:4], gn[:, 4].astype(np.int32)
        fl = np.floor(gyx).astype(np.int64)
        ayx = (fl.astype(F32) + F32(0.5))[:, None, :]                                 # [G,1,2] cell centre
        a1 = (ayx - pri[None] / F32(2)).astype(F32)                                    # [G,3,2]
        a2 = (ayx + pri[None] / F32(2)).astype(F32)
        g1 = (gyx - ghw / F32(2.0)).astype(F32)[:, None, :]
        g2 = (gyx + ghw / F32(2.0)).astype(F32)[:, None, :]
        inter = np.prod(np.minimum(g2, a2) - np.maximum(g1, a1), axis=-1).astype(F32)  # no clamp (:171-173)
        garea = np.prod(g2 - g1, axis=-1).astype(F32)
        aarea = np.prod(pri, axis=-1).astype(F32)[None]
        iou = (inter / (aarea + garea - inter)).astype(F32)                            # [G,3]
        lv.append(dict(r=r, pri=pri, gyx=gyx, ghw=ghw, lab=lab, fl=fl, g1=g1[:, 0], g2=g2[:, 0],
                       idx=np.argmax(iou, axis=-1), mx=iou.max(axis=-1), h=h, w=w))
    m1 = (lv[0]["mx"] > lv[1]["mx"]) & (lv[0]["mx"] > lv[2]["mx"])
    m2 = (lv[1]["mx"] > lv[0]["mx"]) & (lv[1]["mx"] > lv[2]["mx"])
    m3 = ~(m1 | m2)
    coord = cls_l = obj_l = noobj = 0.0
    for L_, m in zip(lv, (m1, m2, m3)):
        r, pri = L_["r"], L_["pri"]
        for gi in np.nonzero(m)[0]:
            y, x = L_["fl"][gi]
            a = L_["idx"][gi]
            row = r[y, x, a]
            tyx = (L_["gyx"][gi] - np.floor(L_["gyx"][gi])).astype(F32)
            thw = np.log((L_["ghw"][gi] / pri[a]).astype(F32)).astype(F32)
            coord += float(_sig_xent(tyx, row[nc:nc + 2]).sum(dtype=np.float64))
            coord += 0.5 * float(np.square(row[nc + 2:nc + 4] - thw).astype(F32).sum(dtype=np.float64))
            onehot = np.zeros(nc, F32)
            onehot[L_["lab"][gi]] = 1.0
            cls_l += float(_sig_xent(onehot, row[:nc]).sum(dtype=np.float64))
            obj_l += float(_sig_xent(np.ones(1, F32), row[nc + 4:nc + 5]).sum(dtype=np.float64))
        # no-object term over the cells that hold no GT centre (:249-311)
        h, w = L_["h"], L_["w"]
        occ = np.zeros((h, w), bool)
        occ[L_["fl"][:, 0], L_["fl"][:, 1]] = True
        cy = (np.arange(h, dtype=F32) + F32(0.5)).reshape(h, 1, 1, 1)
        cx = (np.arange(w, dtype=F32) + F32(0.5)).reshape(1, w, 1, 1)
        ayx = np.concatenate([cy + np.zeros((1, w, 3, 1), F32), cx + np.zeros((h, 1, 3, 1), F32)], -1).astype(F32)
        yx_nb = (ayx - pri.reshape(1, 1, 3, 2) / F32(2.0)).astype(F32)      # really y1x1
        hw_nb = (ayx + pri.reshape(1, 1, 3, 2) / F32(2.0)).astype(F32)      # really y2x2
        b1 = (yx_nb - hw_nb / F32(2.0)).astype(F32)[..., None, :]             # [h,w,3,1,2]
        b2 = (yx_nb + hw_nb / F32(2.0)).astype(F32)[..., None, :]
        gg1, gg2 = L_["g1"].reshape(1, 1, 1, -1, 2), L_["g2"].reshape(1, 1, 1, -1, 2)
        inter = np.prod(np.minimum(gg2, b2) - np.maximum(gg1, b1), axis=-1).astype(F32)
        aarea = np.prod(b2 - b1, axis=-1).astype(F32)
        garea = np.prod(gg2 - gg1, axis=-1).astype(F32)
        iou = (inter / (aarea + garea - inter)).astype(F32).max(axis=-1)       # [h,w,3]
        mask = (iou <= F32(0.5)) & (~occ)[..., None]
        noobj += float((_sig_xent(np.zeros_like(r[..., nc + 4]), r[..., nc + 4]) * mask).sum(dtype=np.float64))
    ng = float(cnt)
    return (coord_scale * coord + class_scale * cls_l + obj_scale * obj_l) / ng + noobj_scale * noobj / ng
