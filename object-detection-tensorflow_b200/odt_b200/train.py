"""Training step of SSD300 / SSD512 (first stage): forward in training mode, the reference's
per-image loss, backward, Momentum(0.9) + L2 weight decay -- what `train_one_epoch` runs per batch
(SSD300.py:129-155,345-453,473-484).

Scope and honesty: this is the "PyTorch autograd first" stage the survey plans.  The forward / backward convolutions
here are PyTorch's (library kernels), in fp32; the hand-written sm_100a kernels of this repository cover the
inference hot path and the loss FORWARD (csrc/loss.cu: the same value this module computes, used as a cross-check).
dgrad / wgrad kernels are the next stage.  Restated from the reference, not copied:

* graph: VGG-16 conv+bias+ReLU (no BN), conv6..conv11_2 / pred convs = conv -> BN -> (ReLU), BN in TRAINING mode
  (`tf.layers.batch_normalization(training=True)`: batch moments, eps 1e-3; moving statistics updated with momentum
  0.99 through UPDATE_OPS, the variance with Bessel's correction as the fused TF kernel reports it  [TF-sem]);
* loss per image (`_compute_one_image_loss`): every GT's arg-max anchor is positive, other anchors with best IoU > 0.5
  are positive, the rest negative; hard-negative mining = NonMaxSuppression over the negative ANCHOR boxes scored by
  their background cross-entropy (IoU 0.7, at most 3 x #positives, gradients flow through the selected scores only);
  loss = mean CE(positives) + mean CE(mined negatives) + mean smooth-L1(positives); batch mean over images;
* + weight_decay * sum(l2_loss(v)) over ALL trainable variables (kernels, biases, BN gamma / beta, the L2-norm scale),
  l2_loss(v) = sum(v^2) / 2;
* `tf.train.MomentumOptimizer(lr, 0.9)`: accum = 0.9 * accum + grad; v -= lr * accum.
"""
import numpy as np
import torch
import torch.nn.functional as F

from . import nets

BN_EPS = 1e-3
BN_MOMENTUM = 0.99


def _same_pad(size, k, stride, dil=1):
    out = -(-size // stride)
    total = max((out - 1) * stride + (k - 1) * dil + 1 - size, 0)
    return total // 2, total - total // 2


def conv_same(x, kernel_hwio, bias, stride=1, dil=1):
    """tf.layers.conv2d / tf.nn.conv2d with SAME padding on NCHW activations; kernel stored HWIO like TF."""
    kh, kw = kernel_hwio.shape[0], kernel_hwio.shape[1]
    pt, pb = _same_pad(x.shape[2], kh, stride, dil)
    pl, pr = _same_pad(x.shape[3], kw, stride, dil)
    if pt or pb or pl or pr:
        x = F.pad(x, (pl, pr, pt, pb))
    return F.conv2d(x, kernel_hwio.permute(3, 2, 0, 1), bias, stride=stride, dilation=dil)


def max_pool_same(x, k, stride):
    pt, pb = _same_pad(x.shape[2], k, stride)
    pl, pr = _same_pad(x.shape[3], k, stride)
    if pt or pb or pl or pr:
        x = F.pad(x, (pl, pr, pt, pb), value=float("-inf"))
    return F.max_pool2d(x, k, stride)


class SSDTrainer:
    """Holds the trainable variables of one SSD model as torch tensors (TF layout / names), the BN moving statistics
    and the Momentum slots.  `step(images, gt, lr)` = one `sess.run([train_op, loss])`."""

    def __init__(self, model, device=None, dtype=torch.float32):
        self.model = model
        self.size = model.input_size
        self.cfg = model.config
        self.weight_decay = float(model.weight_decay)
        self.device = torch.device(device if device is not None else ("cuda" if torch.cuda.is_available() else "cpu"))
        self.dtype = dtype
        w = model.get_weights()
        self.params, self.buffers, self.slots = {}, {}, {}
        for name, arr in w.items():
            t = torch.tensor(np.asarray(arr, np.float32), device=self.device, dtype=dtype)
            if name.endswith(("moving_mean", "moving_variance")):
                self.buffers[name] = t
            else:
                self.params[name] = t.requires_grad_(True)
                self.slots[name] = torch.zeros_like(t)
        self._anchors = None
        self.global_step = int(getattr(model, "global_step", 0))

    # ------------------------------------------------------------------ graph --
    def _bn_train(self, x, scope, stats):
        g, b = self.params[scope + "/gamma"], self.params[scope + "/beta"]
        mean = x.mean(dim=(0, 2, 3))
        var = x.var(dim=(0, 2, 3), unbiased=False)
        n = x.shape[0] * x.shape[2] * x.shape[3]
        stats.append((scope, mean.detach(), (var * (n / max(n - 1, 1))).detach()))
        inv = torch.rsqrt(var + BN_EPS) * g
        return x * inv.view(1, -1, 1, 1) + (b - mean * inv).view(1, -1, 1, 1)

    def forward_rows(self, images, stats=None):
        """images [B,H,W,3] float (RGB 0..255) -> candidate rows [B,N,25] (21 logits, ty, tx, th, tw), training-mode
        BN.  Layer / variable order follows nets.build_ssd (same names as the reference's checkpoints)."""
        P = self.params
        stats = [] if stats is None else stats
        x = torch.as_tensor(images, device=self.device, dtype=self.dtype)
        mean = torch.tensor([123.68, 116.779, 103.979], device=self.device, dtype=self.dtype)
        x = (x - mean).permute(0, 3, 1, 2)
        conv4_3 = None
        for item in nets._VGG:
            if item == "pool":
                x = max_pool_same(x, 2, 2)
                continue
            lname, kn, bnm, _ = item
            x = F.relu(conv_same(x, P["feature_extractor/" + kn], P["feature_extractor/" + bnm]))
            if lname == "conv4_3":
                conv4_3 = x
        x = max_pool_same(x, 3, 1)
        bn_idx = [0]

        def bn_scope(prefix):
            k = bn_idx[0]
            bn_idx[0] += 1
            return prefix + ("/batch_normalization" if k == 0 else "/batch_normalization_%d" % k)

        def cl(x, k, s, name, dil=1, act=True, prefix="feature_extractor"):
            y = conv_same(x, P["%s/%s/kernel" % (prefix, name)], P["%s/%s/bias" % (prefix, name)], s, dil)
            y = self._bn_train(y, bn_scope(prefix), stats)
            return F.relu(y) if act else y

        conv6 = cl(x, 3, 1, "conv6", dil=2)
        conv7 = cl(conv6, 1, 1, "conv7")
        conv8_2 = cl(cl(conv7, 1, 1, "conv8_1"), 3, 2, "conv8_2")
        conv9_2 = cl(cl(conv8_2, 1, 1, "conv9_1"), 3, 2, "conv9_2")
        conv10_2 = cl(cl(conv9_2, 1, 1, "conv10_1"), 3, 1, "conv10_2")
        conv11_2 = cl(cl(conv10_2, 1, 1, "conv11_1"), 3, 2, "conv11_2")
        feats = [conv4_3, conv7, conv8_2, conv9_2, conv10_2, conv11_2]
        if self.size == 512:
            feats.append(cl(cl(conv11_2, 1, 1, "conv12_1"), 3, 2, "conv12_2"))
        # conv4_3: x * rsqrt(max(sum_c x^2, 1e-12)) * gamma  (SSD300.py:74-83)
        f0 = feats[0]
        f0 = f0 * torch.rsqrt(torch.clamp((f0 * f0).sum(dim=1, keepdim=True), min=1e-12))
        feats[0] = f0 * P["feature_extractor/l2_norm_factor"].view(1, 1, 1, 1)
        bn_idx[0] = 0
        rows = []
        for i, f in enumerate(feats):
            p = cl(f, 3, 1, "pred%d" % (i + 1), act=False, prefix="regressor")   # conv -> BN, no activation
            B, C, H, W = p.shape
            rows.append(p.permute(0, 2, 3, 1).reshape(B, H * W * (C // 25), 25))
        self._shapes = [(f.shape[2], f.shape[3]) for f in feats]
        return torch.cat(rows, dim=1)

    def anchors(self):
        if self._anchors is None:
            scales, ratios = nets.ssd_scales(self.size), nets.ssd_ratios(self.size)
            y1x1, y2x2 = [], []
            for (h, w), s, ar in zip(self._shapes, scales, ratios):
                pri = [[s[0], s[0]], [s[1], s[1]]] + [[s[0] * (a ** 0.5), s[0] / (a ** 0.5)] for a in ar]
                pri = torch.tensor(pri, dtype=torch.float32).view(1, 1, -1, 2)
                cy = (torch.arange(h, dtype=torch.float32) + 0.5) * float(self.size) / float(h)
                cx = (torch.arange(w, dtype=torch.float32) + 0.5) * float(self.size) / float(w)
                yx = torch.stack(torch.meshgrid(cy, cx, indexing="ij"), dim=-1).view(h, w, 1, 2)
                y1x1.append((yx - pri / 2.0).reshape(-1, 2))
                y2x2.append((yx + pri / 2.0).reshape(-1, 2))
            a1 = torch.cat(y1x1).to(self.device, self.dtype)
            a2 = torch.cat(y2x2).to(self.device, self.dtype)
            self._anchors = (a1, a2, a1 / 2.0 + a2 / 2.0, a2 - a1)
        return self._anchors

    # ------------------------------------------------------------------- loss --
    @staticmethod
    def _smooth_l1(x):
        ax = x.abs()
        return torch.where(ax < 1.0, 0.5 * x * x, ax - 0.5)

    def image_loss(self, row, gt):
        """One image: row [N,25], gt [G,5] (y, x, h, w, id) padded with -1.  Returns (loss, #pos, #neg, #mined)."""
        from torchvision.ops import nms
        a1, a2, ayx, ahw = self.anchors()
        gt = torch.as_tensor(gt, device=self.device, dtype=self.dtype)
        cnt = int(torch.argmin(gt[:, 0]).item())              # first padded row (SSD300.py:347-348)
        g = gt[:cnt]
        gyx, ghw, label = g[:, 0:2], g[:, 2:4], g[:, 4].long()
        g1, g2 = gyx - ghw / 2.0, gyx + ghw / 2.0
        pconf, pyx, phw = row[:, :21], row[:, 21:23], row[:, 23:25]
        with torch.no_grad():
            i1 = torch.maximum(a1[None], g1[:, None])
            i2 = torch.minimum(a2[None], g2[:, None])
            inter = torch.clamp(i2 - i1, min=0).prod(dim=-1)
            iou = inter / (ahw.prod(dim=-1)[None] + ghw.prod(dim=-1)[:, None] - inter)
            best = torch.argmax(iou, dim=1)
            other = torch.ones(iou.shape[1], dtype=torch.bool, device=self.device)
            other[best] = False
            o_idx = torch.nonzero(other).squeeze(1)
            o_best, rg = iou[:, o_idx].max(dim=0)
            pos = o_best > 0.5
            pos_idx = torch.cat([best, o_idx[pos]])
            pos_g = torch.cat([torch.arange(cnt, device=self.device), rg[pos]])
            neg_idx = o_idx[~pos]
        num_pos, num_neg = int(pos_idx.numel()), int(neg_idx.numel())
        chosen = 3 * num_pos if num_neg > 3 * num_pos else num_neg
        neg_l = F.cross_entropy(pconf[neg_idx], torch.full((num_neg,), 20, device=self.device, dtype=torch.long),
                                reduction="none")
        with torch.no_grad():
            nbox = torch.cat([ayx[neg_idx] - ahw[neg_idx] / 2.0, ayx[neg_idx] + ahw[neg_idx] / 2.0], dim=1)
            sel = nms(nbox.float(), neg_l.detach().float(), 0.7)[:chosen]
        neg_loss = neg_l[sel].mean()
        pos_loss = F.cross_entropy(pconf[pos_idx], label[pos_g], reduction="none").mean()
        tyx = (gyx[pos_g] - ayx[pos_idx]) / ahw[pos_idx]
        thw = torch.log(ghw[pos_g] / ahw[pos_idx])
        coord = (self._smooth_l1(pyx[pos_idx] - tyx).sum(-1) + self._smooth_l1(phw[pos_idx] - thw).sum(-1)).mean()
        return neg_loss + pos_loss + coord, num_pos, num_neg, int(sel.numel())

    def total_loss(self, rows, ground_truth):
        per = [self.image_loss(rows[b], ground_truth[b])[0] for b in range(rows.shape[0])]
        data = torch.stack(per).sum() / rows.shape[0]
        l2 = sum((p * p).sum() for p in self.params.values()) * 0.5
        return data + self.weight_decay * l2, data

    # ------------------------------------------------------------------- step --
    def step(self, images, ground_truth, lr):
        stats = []
        rows = self.forward_rows(images, stats)
        loss, _ = self.total_loss(rows, ground_truth)
        grads = torch.autograd.grad(loss, list(self.params.values()))
        with torch.no_grad():
            for (name, p), gr in zip(self.params.items(), grads):
                acc = self.slots[name]
                acc.mul_(0.9).add_(gr)
                p.sub_(acc, alpha=float(lr))
            for scope, mean, var in stats:          # UPDATE_OPS: moving = moving * 0.99 + batch * 0.01
                self.buffers[scope + "/moving_mean"].mul_(BN_MOMENTUM).add_(mean, alpha=1 - BN_MOMENTUM)
                self.buffers[scope + "/moving_variance"].mul_(BN_MOMENTUM).add_(var, alpha=1 - BN_MOMENTUM)
        self.global_step += 1
        return float(loss.detach())

    def export(self):
        """name -> float32 ndarray of every variable (TF layout), for save_weight / the inference engines."""
        out = {k: v.detach().float().cpu().numpy() for k, v in self.params.items()}
        out.update({k: v.detach().float().cpu().numpy() for k, v in self.buffers.items()})
        return out
