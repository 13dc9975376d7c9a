"""Generates the committed golden fixtures tests/golden/*.npz.

The reference cannot run here (TensorFlow 1.13 absent; SSD300.py does not
parse), so these vectors are produced by the CPU oracle (oracle/tails.py) on
seeded synthetic head rows (logits ~N(0,2^2), box deltas
~N(0,0.5^2), seed 2) and frozen: the GPU path is compared against them on the
GPU box, and tests/test_oracle.py checks the oracle still reproduces them.
Run:  python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
for p in (ROOT, os.path.join(ROOT, "object-detection-tensorflow_b200")):
    if p not in sys.path:
        sys.path.insert(0, p)

from oracle import tails as OT  # noqa: E402

YOLO_PRIORS = [[[10, 13], [16, 30], [33, 23]], [[30, 61], [62, 45], [59, 119]],
               [[116, 90], [156, 198], [373, 326]]]

# name -> geometry (levels are (H, W, A)) and thresholds
CASES = {
    "tail_ssd": dict(kind="ssd", size=300,
                     levels=[(38, 38, 4), (19, 19, 6), (10, 10, 6), (5, 5, 6), (5, 5, 4), (3, 3, 4)],
                     score_thr=0.30, max_boxes=20, iou_thr=0.5),
    "tail_retina": dict(kind="retina", data_shape=[128, 128, 3],
                        levels=[(16, 16, 9), (8, 8, 9), (4, 4, 9), (2, 2, 9), (1, 1, 9)],
                        score_thr=0.25, max_boxes=10, iou_thr=0.45),
    "tail_yolo": dict(kind="yolo", levels=[(5, 5, 3), (10, 10, 3), (20, 20, 3)],
                      score_thr=0.45, max_boxes=10, iou_thr=0.45),
    "tail_fcos": dict(kind="fcos", levels=[(32, 32, 1), (16, 16, 1), (8, 8, 1), (4, 4, 1), (2, 2, 1)],
                      score_thr=0.40, max_boxes=10, iou_thr=0.45),
}


def make_rows(name, batch=1, seed=2):
    c = CASES[name]
    n = sum(h * w * a for h, w, a in c["levels"])
    rng = np.random.default_rng(seed)
    rows = np.empty((batch, n, 25), np.float32)
    if c["kind"] in ("ssd", "retina"):
        rows[..., :21] = rng.standard_normal((batch, n, 21)) * 2.0
        rows[..., 21:] = rng.standard_normal((batch, n, 4)) * 0.5
    elif c["kind"] == "yolo":
        rows[..., :20] = rng.standard_normal((batch, n, 20)) * 2.0
        rows[..., 20:24] = rng.standard_normal((batch, n, 4)) * 0.5
        rows[..., 24] = rng.standard_normal((batch, n)) * 2.0 + 1.0
    else:
        rows[..., :20] = rng.standard_normal((batch, n, 20)) * 2.0
        rows[..., 20] = rng.standard_normal((batch, n)) * 2.0 + 1.0
        rows[..., 21:] = rng.standard_normal((batch, n, 4)) * 0.5 + 1.0
    return rows


def run_case(name, rows, image=0):
    """Oracle tail on one image's rows [N,25] (or [B,N,25]) -> (scores, bbox, class_id, keep)."""
    c = CASES[name]
    rows = np.asarray(rows, np.float32)
    if rows.ndim == 2:
        rows = rows[None]
    b = rows.shape[0]

    def split(width):
        out, off = [], 0
        for h, w, a in c["levels"]:
            out.append(rows[:, off:off + h * w * a].reshape(b, h, w, a * width) if width else None)
            off += h * w * a
        return out

    if c["kind"] == "ssd":
        preds = split(25)
        return OT.ssd_detect(preds, c["size"], c["score_thr"], c["max_boxes"], c["iou_thr"], image)
    if c["kind"] == "retina":
        heads, off = [], 0
        for h, w, a in c["levels"]:
            r = rows[:, off:off + h * w * a]
            heads.append((r[..., :21].reshape(b, h, w, a * 21), r[..., 21:].reshape(b, h, w, a * 4)))
            off += h * w * a
        return OT.retina_detect(heads, c["data_shape"], c["score_thr"], c["max_boxes"], c["iou_thr"], image)
    if c["kind"] == "yolo":
        preds = split(25)
        return OT.yolo_detect(preds, YOLO_PRIORS, c["score_thr"], c["max_boxes"], c["iou_thr"], image)
    heads, off = [], 0
    for h, w, a in c["levels"]:
        r = rows[:, off:off + h * w]
        heads.append((r[..., :20].reshape(b, h, w, 20), r[..., 20:21].reshape(b, h, w, 1),
                      r[..., 21:].reshape(b, h, w, 4)))
        off += h * w
    return OT.fcos_detect(heads, c["score_thr"], c["max_boxes"], c["iou_thr"], image)


# ---------------------------------------------------------------- losses ----
# kind -> (tail case whose geometry / rows are re-used, image size that places the GT boxes)
LOSS_CASES = {"loss_ssd": ("tail_ssd", 300), "loss_retina": ("tail_retina", 128), "loss_yolo": ("tail_yolo", 160),
              "loss_fcos": ("tail_fcos", 256)}


def make_gt(name, batch=2, seed=7, G=12):
    _, size = LOSS_CASES[name]
    rng = np.random.default_rng(seed)
    gt = np.full((batch, G, 5), -1.0, np.float32)
    for b in range(batch):
        n = 2 + 3 * b
        gt[b, :n, 0:2] = rng.uniform(0.2 * size, 0.8 * size, (n, 2))
        gt[b, :n, 2:4] = rng.uniform(0.08 * size, 0.6 * size, (n, 2))
        gt[b, :n, 4] = rng.integers(0, 20, n)
    return gt


def run_loss_case(name, rows, gt, image=0):
    """Oracle loss of one image on candidate rows [B,N,25]."""
    from oracle import loss as OL
    tail, _ = LOSS_CASES[name]
    c = CASES[tail]
    rows = np.asarray(rows, np.float32)
    B = rows.shape[0]
    shapes = [(h, w) for h, w, _ in c["levels"]]
    r = rows[image]
    if name == "loss_ssd":
        a1, a2, ayx, ahw = OT.ssd_anchors(c["size"], shapes)
        return OL.ssd_image_loss(r[:, :21], r[:, 21:23], r[:, 23:], a1, a2, ayx, ahw, gt[image])[0]
    if name == "loss_retina":
        a1, a2, ayx, ahw = OT.retina_anchors(c["data_shape"], shapes)
        return OL.retina_image_loss(r[:, :21], r[:, 21:23], r[:, 23:], a1, a2, ayx, ahw, gt[image])[0]
    off, lv = 0, []
    for h, w, a in c["levels"]:
        blk = rows[:, off:off + h * w * a]
        if name == "loss_yolo":
            lv.append(blk.reshape(B, h, w, a * 25))
        else:
            q = blk.reshape(B, h, w, 25)
            lv.append((q[..., :20], q[..., 20:21], q[..., 21:25]))
        off += h * w * a
    if name == "loss_yolo":
        return OL.yolo_image_loss(lv, YOLO_PRIORS, gt[image], image=image)
    return OL.fcos_image_loss(lv, gt[image], image=image)


def main():
    for name, (tail, _) in LOSS_CASES.items():
        rows = make_rows(tail, batch=2, seed=5)
        gt = make_gt(name)
        losses = np.asarray([run_loss_case(name, rows, gt, b) for b in range(2)], np.float64)
        np.savez_compressed(os.path.join(HERE, name + ".npz"), rows=rows, gt=gt, loss=losses)
        print(name, "rows", rows.shape, "loss", losses)
    for name in CASES:
        rows = make_rows(name)[0]
        s, bx, cid, keep = run_case(name, rows)
        assert len(np.unique(s)) == len(s) or len(s) == 0, "score ties in %s" % name
        np.savez_compressed(os.path.join(HERE, name + ".npz"), rows=rows, scores=s, bbox=bx,
                            class_id=cid, keep=keep)
        print(name, "rows", rows.shape, "detections", len(s), "classes", len(np.unique(cid)))


if __name__ == "__main__":
    main()
