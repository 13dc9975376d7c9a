"""Backbone + head graphs of the five detectors, restated on CPU (test oracle).

Each function walks the layers in the reference's creation order, pulling
variables from a dict keyed by TF1-style variable names, and
returns the per-level head tensors exactly as the reference's graph holds them
before the inference tail.  Not product code -- see oracle/__init__.py.
"""
from contextlib import contextmanager

import numpy as np

from . import tfops as T

F32 = np.float32


class Namer:
    """TF1 variable-scope naming: explicit names stay, default names are made
    unique per enclosing scope in creation order (conv2d, conv2d_1, ...)."""

    def __init__(self):
        self.stack = []
        self.counts = {}

    @contextmanager
    def scope(self, name):
        self.stack.append(name)
        try:
            yield
        finally:
            self.stack.pop()

    def prefix(self):
        return "/".join(self.stack)

    def unique(self, base):
        key = (self.prefix(), base)
        k = self.counts.get(key, 0)
        self.counts[key] = k + 1
        name = base if k == 0 else "%s_%d" % (base, k)
        return self.prefix() + "/" + name if self.stack else name

    def named(self, name):
        return self.prefix() + "/" + name if self.stack else name


class Ctx:
    def __init__(self, weights):
        self.w = weights
        self.n = Namer()

    # tf.layers.conv2d (default or explicit name) -- kernel HWIO + bias
    def conv(self, x, k, stride=1, dil=1, name=None):
        vs = self.n.named(name) if name else self.n.unique("conv2d")
        return T.conv2d_same(x, self.w[vs + "/kernel"], self.w[vs + "/bias"], stride, dil)

    def bn(self, x):
        vs = self.n.unique("batch_normalization")
        return T.batch_norm_inference(x, self.w[vs + "/gamma"], self.w[vs + "/beta"],
                                      self.w[vs + "/moving_mean"], self.w[vs + "/moving_variance"])

    def gn(self, x):
        vs = self.n.unique("GroupNorm")
        return T.group_norm(x, self.w[vs + "/gamma"], self.w[vs + "/beta"])


# ----------------------------------------------------------------- SSD ------
_VGG = [("conv1_1", "kernel_conv1_1", "bias_conv1_1"), ("conv1_2", "kernel_conv1_2", "bias_conv1_2"),
        "pool",
        ("conv2_1", "kenrel_conv2_1", "bias_conv2_1"), ("conv2_2", "kernel_conv2_2", "bias_conv2_2"),
        "pool",
        ("conv3_1", "kernel_conv3_1", "bias_conv_3_1"), ("conv3_2", "kernel_conv3_2", "bias_conv3_2"),
        ("conv3_3", "kernel_conv3_3", "bias_conv3_3"), "pool",
        ("conv4_1", "kernel_conv4_1", "bias_conv4_1"), ("conv4_2", "kernel_conv4_2", "bias_conv4_2"),
        ("conv4_3", "kernel_conv4_3", "bias_conv4_3"), "pool",
        ("conv5_1", "kernel_conv5_1", "bias_conv5_1"), ("conv5_2", "kernel_conv5_2", "bias_conv5_2"),
        ("conv5_3", "kernel_conv5_3", "bias_conv5_3")]


def ssd_heads(weights, images, size=300, num_classes=21):
    """SSD300.py:71-90,192-314 (SSD512.py:320-322 adds conv12): returns the list of
    pred tensors [B,H,W,A*(C+4)] (BN applied, no activation) per scale."""
    c = Ctx(weights)
    x = (images.astype(F32) - T.RGB_MEAN.reshape(1, 1, 1, 3)).astype(F32)  # :52-66
    feats = []
    with c.n.scope("feature_extractor"):
        for item in _VGG:  # :193-302  conv3x3 + bias + ReLU, no BN
            if item == "pool":
                x = T.max_pool_same(x, 2, 2)
                continue
            lname, kn, bn_ = item
            x = T.relu(T.conv2d_same(x, weights["feature_extractor/" + kn],
                                     weights["feature_extractor/" + bn_]))
            if lname == "conv4_3":
                conv4_3 = x
        x = T.max_pool_same(x, 3, 1)  # pool5 :303

        def cl(x, k, s, name, dil=1):  # _conv_layer :523-537  conv -> BN -> relu
            return T.relu(c.bn(c.conv(x, k, s, dil, name=name)))

        conv6 = cl(x, 3, 1, "conv6", dil=2)
        conv7 = cl(conv6, 1, 1, "conv7")
        conv8_2 = cl(cl(conv7, 1, 1, "conv8_1"), 3, 2, "conv8_2")
        conv9_2 = cl(cl(conv8_2, 1, 1, "conv9_1"), 3, 2, "conv9_2")
        conv10_2 = cl(cl(conv9_2, 1, 1, "conv10_1"), 3, 1, "conv10_2")
        conv11_2 = cl(cl(conv10_2, 1, 1, "conv11_1"), 3, 2, "conv11_2")
        feats = [conv4_3, conv7, conv8_2, conv9_2, conv10_2, conv11_2]
        if size == 512:
            feats.append(cl(cl(conv11_2, 1, 1, "conv12_1"), 3, 2, "conv12_2"))
        # :74-83 l2-normalise conv4_3 over channels, times learned scalar
        f1 = T.l2_normalize_channels(feats[0])
        feats[0] = (weights["feature_extractor/l2_norm_factor"].astype(F32)[0] * f1).astype(F32)
    preds = []
    with c.n.scope("regressor"):
        for i, f in enumerate(feats):  # :85-90 pred convs: conv -> BN, no activation
            preds.append(c.bn(c.conv(f, 3, 1, name="pred%d" % (i + 1))))
    return preds


# ------------------------------------------------------------- RetinaNet ----
def _bac(c, x, k, stride, norm="bn"):
    """_bn_activation_conv (RetinaNet.py:594-619 / FCOS.py:467-489): norm -> relu -> conv."""
    x = c.bn(x) if norm == "bn" else c.gn(x)
    return c.conv(T.relu(x), k, stride)


def _bottleneck(c, x, f, stride, scope, norm):
    """_residual_bottleneck (RetinaNet.py:634-643): shortcut is always a 3x3 conv."""
    with c.n.scope(scope):
        with c.n.scope("conv_branch"):
            y = _bac(c, x, 1, 1, norm)
            y = _bac(c, y, 3, stride, norm)
            y = _bac(c, y, 1, 1, norm)
        with c.n.scope("identity_branch"):
            sc = _bac(c, x, 3, stride, norm)
    return (y + sc).astype(F32)


def _resnet(c, x, filters, blocks, norm):
    """stem + pool + bottleneck stacks (RetinaNet.py:258-285, FCOS.py:72-97)."""
    y = c.conv(x, 7, 2)
    y = T.relu(c.bn(y) if norm == "bn" else c.gn(y))
    y = T.max_pool_same(y, 3, 2)
    ends = []
    for i in range(blocks[0]):
        y = _bottleneck(c, y, filters[0], 1, "block1_unit%d" % (i + 1), norm)
    ends.append(y)
    for i in range(1, len(blocks)):
        y = _bottleneck(c, y, filters[i], 2, "block%d_unit1" % (i + 1), norm)
        for j in range(1, blocks[i]):
            y = _bottleneck(c, y, filters[i], 1, "block%d_unit%d" % (i + 1, j + 1), norm)
        ends.append(y)
    return ends[-3], ends[-2], ends[-1]


def _pyramid(c, feat, top, norm):
    """_get_pyramid (RetinaNet.py:303-319)."""
    if top is None:
        return _bac(c, feat, 3, 1, norm), None
    f = _bac(c, feat, 1, 1, norm)
    t = T.resize_bilinear_legacy(top, f.shape[1], f.shape[2])
    total = (f + t).astype(F32)
    return _bac(c, total, 3, 1, norm), total


def retinanet_heads(weights, images, block_list=(3, 4, 6, 3), init_conv_filters=16,
                    num_classes=21, num_anchors=9):
    """RetinaNet.py:137-155: returns [(cls [B,H,W,9*21], reg [B,H,W,9*4])] for P3..P7.
    Backbone widths are 7*2^i (RetinaNet.py:27 uses init_conv_kernel_size)."""
    c = Ctx(weights)
    x = (images.astype(F32) - T.RGB_MEAN.reshape(1, 1, 1, 3)).astype(F32)
    filters = [7 * (2 ** i) for i in range(len(block_list))]
    with c.n.scope("feature_extractor"):
        # stem conv uses config['init_conv_filters'] output channels (:260-265)
        y = T.relu(c.bn(c.conv(x, 7, 2)))
        assert y.shape[3] == init_conv_filters
        y = T.max_pool_same(y, 3, 2)
        ends = []
        for i in range(block_list[0]):
            y = _bottleneck(c, y, filters[0], 1, "block1_unit%d" % (i + 1), "bn")
        ends.append(y)
        for i in range(1, len(block_list)):
            y = _bottleneck(c, y, filters[i], 2, "block%d_unit1" % (i + 1), "bn")
            for j in range(1, block_list[i]):
                y = _bottleneck(c, y, filters[i], 1, "block%d_unit%d" % (i + 1, j + 1), "bn")
            ends.append(y)
        f1, f2, f3 = ends[-3], ends[-2], ends[-1]
        p5, _ = _pyramid(c, f3, None, "bn")
        p4, td = _pyramid(c, f2, p5, "bn")
        p3, _ = _pyramid(c, f1, td, "bn")
        p6 = _bac(c, p5, 3, 2)
        p7 = _bac(c, p6, 3, 2)
    out = []
    with c.n.scope("regressor"):
        for p in (p3, p4, p5, p6, p7):  # :146-155, towers are NOT shared
            y = p
            for _ in range(4):
                y = _bac(c, y, 3, 1)
            cls = _bac(c, y, 3, 1)
            y = p
            for _ in range(4):
                y = _bac(c, y, 3, 1)
            reg = _bac(c, y, 3, 1)
            out.append((cls, reg))
    return out


# ---------------------------------------------------------------- YOLOv3 ----
def yolov3_heads(weights, images, num_classes=20, num_priors=3):
    """YOLOv3.py:81-88,387-417,485-507: returns [pred1(13x13), pred2, pred3] [B,H,W,75]."""
    c = Ctx(weights)
    x = (images.astype(F32) - T.RGB_MEAN.reshape(1, 1, 1, 3)).astype(F32)

    def cl(x, f, k, s, act=True):  # _conv_layer :494-507 conv -> BN -> leaky
        y = c.bn(c.conv(x, k, s))
        assert y.shape[3] == int(f)
        return T.leaky_relu(y) if act else y

    def block(x, f, n, scope):  # _darknet_block :485-492
        with c.n.scope(scope):
            y = cl(x, f, 3, 2)
            for _ in range(n):
                y = (y + cl(cl(y, f // 2, 1, 1), f, 3, 1)).astype(F32)
        return y

    with c.n.scope("backone"):
        y = cl(x, 32, 3, 1)
        b1 = block(y, 64, 1, "block1")
        b2 = block(b1, 128, 2, "block2")
        b3 = block(b2, 256, 8, "block3")
        b4 = block(b3, 512, 8, "block4")
        b5 = block(b4, 1024, 4, "block5")

    def header(bottom, f, scope, pyramid=None):  # _yolo3_header :396-417
        with c.n.scope(scope):
            if pyramid is not None:
                u = cl(pyramid, f, 1, 1, act=False)
                u = T.resize_nearest_legacy(u, bottom.shape[1], bottom.shape[2])
                y = np.concatenate([bottom, u], axis=3)
            else:
                y = bottom
            c1 = cl(y, f // 2, 1, 1)
            c2 = cl(c1, f, 3, 1)
            c3 = cl(c2, f // 2, 1, 1)
            c4 = cl(c3, f, 3, 1)
            c5 = cl(c4, f // 2, 1, 1)
            c6 = cl(c5, f, 3, 1)
            pred = cl(c6, (num_classes + 5) * num_priors, 1, 1)  # BN + leaky too (:416)
        return pred, c5

    with c.n.scope("head"):
        p1, td = header(b5, 1024, "pyd1")
        p2, td = header(b4, 256, "pyd2", td)
        p3, _ = header(b3, 128, "pyd3", td)
    return [p1, p2, p3]


# ------------------------------------------------------------------ FCOS ----
def fcos_heads(weights, images, num_classes=20, share_heads=True):
    """FCOS.py:70-107,350-364: returns [(cls [B,H,W,20], ctr [B,H,W,1], reg_raw [B,H,W,4])]
    for P3..P7; reg_raw is the conv output BEFORE tf.exp (:363)."""
    c = Ctx(weights)
    x = (images.astype(F32) - T.RGB_MEAN.reshape(1, 1, 1, 3)).astype(F32)
    filters = [16 * (2 ** i) for i in range(4)]
    with c.n.scope("backone"):
        e3, e4, e5 = _resnet(c, x, filters, [3, 4, 6, 3], "gn")
    with c.n.scope("pyramid"):
        c3 = _bac(c, e3, 1, 1, "gn")
        c4 = _bac(c, e4, 1, 1, "gn")
        c5 = _bac(c, e5, 1, 1, "gn")
        p5, _ = _pyramid(c, c5, None, "gn")
        p4, td = _pyramid(c, c4, p5, "gn")
        p3, _ = _pyramid(c, c3, td, "gn")
        p6 = _bac(c, p5, 3, 2, "gn")
        p7 = _bac(c, p6, 3, 2, "gn")
    out = []
    with c.n.scope("head"):
        for p in (p3, p4, p5, p6, p7):
            if share_heads:  # AUTO_REUSE: every level re-uses the first level's variables
                c.n.counts = {k: v for k, v in c.n.counts.items() if not k[0].startswith("head/")}
            with c.n.scope("classifier_head"):
                y = p
                for _ in range(4):
                    y = _bac(c, y, 3, 1, "gn")
                cls = _bac(c, y, 3, 1, "gn")
                ctr = _bac(c, y, 3, 1, "gn")
            with c.n.scope("regress_head"):
                y = p
                for _ in range(4):
                    y = _bac(c, y, 3, 1, "gn")
                reg = _bac(c, y, 3, 1, "gn")
            out.append((cls, ctr, reg))
    return out
