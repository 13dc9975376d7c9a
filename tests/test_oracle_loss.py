"""Known answers of the training-loss oracles (oracle/loss.py), derived by hand from the reference
formulas on tiny inputs -- the pins the GPU loss kernels inherit.  CPU only.  (Parity unpinned against
TensorFlow itself: there is no TF here; these check that the restatement computes what the cited
reference lines say.)"""
import math

import numpy as np

F32 = np.float32


def _anchors(boxes):
    b = np.asarray(boxes, F32)           # rows y1,x1,y2,x2
    y1x1, y2x2 = b[:, :2], b[:, 2:]
    yx = (y1x1 / F32(2) + y2x2 / F32(2)).astype(F32)
    hw = (y2x2 - y1x1).astype(F32)
    return y1x1, y2x2, yx, hw


def test_ssd_loss_two_anchors_by_hand():
    """One GT identical to anchor 0 (IoU 1 -> the GT's best anchor), anchor 1 disjoint (negative).
    Uniform logits -> every cross-entropy is log(21); regression targets are 0, predictions 0 -> no box
    loss.  chosen negatives = min(1, 3*1) = 1.  loss = log 21 (neg) + log 21 (pos) + 0.  SSD300.py:345-453"""
    from oracle import loss as OL
    a1, a2, ayx, ahw = _anchors([[10, 10, 50, 50], [200, 200, 240, 240]])
    gt = np.full((4, 5), -1.0, F32)
    gt[0] = [30, 30, 40, 40, 7]
    pconf = np.zeros((2, 21), F32)
    loss, info = OL.ssd_image_loss(pconf, np.zeros((2, 2), F32), np.zeros((2, 2), F32), a1, a2, ayx, ahw, gt)
    assert info == dict(num_pos=1, num_neg=1, selected=1)
    assert abs(loss - 2 * math.log(21.0)) < 1e-5


def test_ssd_loss_hard_negative_mining_uses_nms_on_anchor_boxes():
    """Three negatives: two heavily overlapping anchors (IoU 0.81 > 0.7) and one far away.  The two
    overlapping ones carry the largest background losses; NMS keeps only the larger of them plus the far
    one, so with chosen = 3 the mean is over 2 boxes, not 3."""
    from oracle import loss as OL
    a1, a2, ayx, ahw = _anchors([[0, 0, 40, 40], [100, 100, 140, 140], [104, 100, 144, 140], [300, 300, 340, 340]])
    gt = np.full((3, 5), -1.0, F32)
    gt[0] = [20, 20, 40, 40, 0]
    pconf = np.zeros((4, 21), F32)
    pconf[1, 20], pconf[2, 20], pconf[3, 20] = -2.0, -3.0, -1.0   # background logit low -> high loss
    loss, info = OL.ssd_image_loss(pconf, np.zeros((4, 2), F32), np.zeros((4, 2), F32), a1, a2, ayx, ahw, gt)
    assert info == dict(num_pos=1, num_neg=3, selected=2)

    def ce_bg(bg):  # 20 zero logits + one bg logit
        return math.log(20.0 + math.exp(bg)) - bg
    neg = (ce_bg(-3.0) + ce_bg(-1.0)) / 2.0
    assert abs(loss - (neg + math.log(21.0))) < 1e-5


def test_fcos_level_loss_single_cell_by_hand():
    """2x2 grid, stride 8, one GT centred at (8, 8) px of size 12x12: in stride units y1=x1=.25,
    y2=x2=1.75 -> only cell (1,1) is strictly inside with l=t=.75, r=b=.75.  Predictions: exp(reg)=.75
    -> IoU 1 -> iou loss = -log(1+1e-12) ~ 0.  Centre target sqrt(.75*.75/(.75*.75)) = 1 at that cell, 0
    elsewhere; all logits 0 -> centre BCE = 4*log 2.  Focal: sigmoid = .5 everywhere: positive cell/class
    -0.25*.25*log(.5), each of the 79 negatives -0.25*.25*log(.5) as well -> 80 * 0.0625*log 2.
    total / #positive cells (1).  FCOS.py:266-348"""
    from oracle import loss as OL
    cls = np.zeros((2, 2, 20), F32)
    reg = np.full((2, 2, 4), math.log(0.75), F32)
    ctr = np.zeros((2, 2), F32)
    gt = np.asarray([[8, 8, 12, 12, 5]], F32)
    got = OL.fcos_level_loss(cls, reg, ctr, gt, 8)
    exp = 4 * math.log(2.0) + 80 * 0.0625 * math.log(2.0)
    assert abs(got - exp) < 1e-4


def test_fcos_level_assignment_bounds_are_inclusive_and_overlapping():
    """sqrt(h*w) == 64 feeds P3 AND P4 (FCOS.py:158-164): the image loss is the sum of both level losses."""
    from oracle import loss as OL
    rng = np.random.default_rng(0)
    heads = [(rng.standard_normal((1, h, h, 20)).astype(F32), rng.standard_normal((1, h, h, 1)).astype(F32),
              rng.standard_normal((1, h, h, 4)).astype(F32)) for h in (16, 8, 4, 2, 1)]
    gt = np.full((3, 5), -1.0, F32)
    gt[0] = [64, 64, 64, 64, 2]
    both = OL.fcos_image_loss(heads, gt)
    p3 = OL.fcos_level_loss(heads[0][0][0], heads[0][2][0], heads[0][1][0, ..., 0], gt[:1], 8)
    p4 = OL.fcos_level_loss(heads[1][0][0], heads[1][2][0], heads[1][1][0, ..., 0], gt[:1], 16)
    assert abs(both - (p3 + p4)) < 1e-9


def test_yolo_loss_single_gt_by_hand():
    """416-style strides on tiny grids (2x2, 4x4, 8x8 -> 64 px image), one GT centred at (40, 24) px of size
    32x32.  All logits 0.  With priors chosen so that level 0's first prior equals the GT exactly in level-0
    units ((32/32, 32/32) -> prior 8x8 px / 8), level 0 wins (IoU 1).  Positive terms: centre targets
    (40/32 - 1, 24/32 - 0) = (.25, .75) with logit 0 -> 2 log 2; size target log(1/1) = 0, prediction 0 -> 0;
    class: 20 log 2; objectness: log 2.  No-object: every (cell, prior) of the cells without the GT centre
    whose re-used-form anchor has IoU <= .5 contributes log 2 -- counted by brute force below."""
    from oracle import loss as OL
    priors = [[[8, 8], [3, 3], [2, 2]], [[20, 20], [24, 24], [28, 28]], [[40, 40], [48, 48], [56, 56]]]
    preds = [np.zeros((1, h, h, 75), F32) for h in (2, 4, 8)]
    gt = np.full((3, 5), -1.0, F32)
    gt[0] = [40, 24, 32, 32, 4]
    got = OL.yolo_image_loss(preds, priors, gt, coord_scale=1, noobj_scale=1, obj_scale=5, class_scale=1)
    ln2 = math.log(2.0)
    pos = (2 * ln2) + 20 * ln2 + 5 * ln2
    # brute-force count of the no-object entries
    n_noobj = 0
    for k, (h, norm, pst) in enumerate(zip((2, 4, 8), (32.0, 16.0, 8.0), (8.0, 16.0, 32.0))):
        gy, gx, gh, gw = 40 / norm, 24 / norm, 32 / norm, 32 / norm
        g1, g2 = (gy - gh / 2, gx - gw / 2), (gy + gh / 2, gx + gw / 2)
        for y in range(h):
            for x in range(h):
                if (y, x) == (int(gy), int(gx)):
                    continue
                for a in range(3):
                    ph, pw = priors[k][a][0] / pst, priors[k][a][1] / pst
                    c = (y + .5 - ph / 2, x + .5 - pw / 2)
                    s = (y + .5 + ph / 2, x + .5 + pw / 2)
                    b1 = (c[0] - s[0] / 2, c[1] - s[1] / 2)
                    b2 = (c[0] + s[0] / 2, c[1] + s[1] / 2)
                    inter = (min(g2[0], b2[0]) - max(g1[0], b1[0])) * (min(g2[1], b2[1]) - max(g1[1], b1[1]))
                    iou = inter / ((b2[0] - b1[0]) * (b2[1] - b1[1]) + gh * gw - inter)
                    n_noobj += iou <= 0.5
    assert abs(got - (pos + n_noobj * ln2)) < 1e-3 * (pos + n_noobj * ln2)
