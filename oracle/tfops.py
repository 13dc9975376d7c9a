"""TensorFlow 1.13 op semantics used on the detection hot path, restated on CPU.

Every function names the reference call site whose arithmetic it stands for
(paths relative to the reference checkout) and the SURVEY.md Appendix A item
that specifies the TF behaviour.  All tensors are numpy float32, NHWC.
[TF-sem, unverifiable here]: see oracle/__init__.py.
"""
import math

import numpy as np
import torch
import torch.nn.functional as F

F32 = np.float32


def same_pad(size, k, stride, dil=1):
    """TF SAME geometry (App. A.1): out=ceil(in/stride); before=total//2, rest after."""
    out = -(-size // stride)
    total = max((out - 1) * stride + (k - 1) * dil + 1 - size, 0)
    return out, total // 2, total - total // 2


def _to_nchw(x):
    return torch.from_numpy(np.ascontiguousarray(x)).permute(0, 3, 1, 2)


def _to_nhwc(t):
    return np.ascontiguousarray(t.permute(0, 2, 3, 1).numpy())


def conv2d_same(x, kernel_hwio, bias=None, stride=1, dil=1):
    """tf.nn.conv2d / tf.layers.conv2d(padding='same') (SSD300.py:519,524; App. A.1-2).
    x [B,H,W,Cin] f32, kernel HWIO, zero padding, bias added after the conv."""
    kh, kw, _, _ = kernel_hwio.shape
    _, pt, pb = same_pad(x.shape[1], kh, stride, dil)
    _, pl, pr = same_pad(x.shape[2], kw, stride, dil)
    t = _to_nchw(x)
    t = F.pad(t, (pl, pr, pt, pb))
    w = torch.from_numpy(np.ascontiguousarray(kernel_hwio)).permute(3, 2, 0, 1).contiguous()
    b = torch.from_numpy(np.ascontiguousarray(bias)) if bias is not None else None
    y = F.conv2d(t, w, b, stride=stride, dilation=dil)
    return _to_nhwc(y)


def batch_norm_inference(x, gamma, beta, mean, var, eps=1e-3):
    """tf.layers.batch_normalization(training=False) (SSD300.py:506-512; App. A.3):
    y = (x - mean) * rsqrt(var + eps) * gamma + beta."""
    inv = (F32(1.0) / np.sqrt(var.astype(F32) + F32(eps))).astype(F32) * gamma.astype(F32)
    return (x * inv + (beta.astype(F32) - mean.astype(F32) * inv)).astype(F32)


def group_norm(x, gamma, beta, groups=8, eps=1e-6):
    """tf.contrib.layers.group_norm(groups=8) (FCOS.py:438-446; App. A.5): per
    (image, group) moments over (H, W, C/groups), biased variance;
    gain = rsqrt(var+eps)*gamma; offset = -mean*gain + beta; y = x*gain + offset."""
    b, h, w, c = x.shape
    xg = x.reshape(b, h, w, groups, c // groups).astype(F32)
    mean = xg.mean(axis=(1, 2, 4), keepdims=True, dtype=np.float64).astype(F32)
    var = ((xg - mean) ** 2).mean(axis=(1, 2, 4), keepdims=True, dtype=np.float64).astype(F32)
    gain = (F32(1.0) / np.sqrt(var + F32(eps))).astype(F32)
    g = gamma.astype(F32).reshape(1, 1, 1, groups, c // groups)
    bt = beta.astype(F32).reshape(1, 1, 1, groups, c // groups)
    gain_c = gain * g
    offset = -mean * gain_c + bt
    return (xg * gain_c + offset).reshape(b, h, w, c).astype(F32)


def max_pool_same(x, k, stride):
    """tf.layers.max_pooling2d(padding='same') (SSD300.py:539-547; App. A.1): pads ignored."""
    _, pt, pb = same_pad(x.shape[1], k, stride)
    _, pl, pr = same_pad(x.shape[2], k, stride)
    t = _to_nchw(x)
    t = F.pad(t, (pl, pr, pt, pb), value=float("-inf"))
    return _to_nhwc(F.max_pool2d(t, k, stride))


def l2_normalize_channels(x, eps=1e-12):
    """tf.nn.l2_normalize(x, axis=3) (SSD300.py:75; App. A.4)."""
    s = np.sum(x.astype(F32) * x.astype(F32), axis=3, keepdims=True, dtype=F32)
    return (x * (F32(1.0) / np.sqrt(np.maximum(s, F32(eps))))).astype(F32)


def relu(x):
    return np.maximum(x, F32(0.0))


def leaky_relu(x, alpha=0.1):
    """tf.nn.leaky_relu(x, 0.1) = max(x, alpha*x) (YOLOv3.py:506; App. A.7)."""
    return np.maximum(x, F32(alpha) * x).astype(F32)


def resize_bilinear_legacy(x, out_h, out_w):
    """tf.image.resize_bilinear, TF1 legacy (align_corners=False, no half-pixel)
    (RetinaNet.py:309; App. A.6): src = dst*scale; lo=floor, hi=min(ceil,in-1);
    top = tl + (tr-tl)*lx; bottom = bl + (br-bl)*lx; out = top + (bottom-top)*ly."""
    b, h, w, c = x.shape
    hs = F32(h) / F32(out_h)
    ws = F32(w) / F32(out_w)
    ys = (np.arange(out_h, dtype=F32) * hs).astype(F32)
    xs = (np.arange(out_w, dtype=F32) * ws).astype(F32)
    y0 = np.floor(ys).astype(np.int64)
    y1 = np.minimum(np.ceil(ys).astype(np.int64), h - 1)
    x0 = np.floor(xs).astype(np.int64)
    x1 = np.minimum(np.ceil(xs).astype(np.int64), w - 1)
    ly = (ys - y0.astype(F32)).astype(F32).reshape(1, out_h, 1, 1)
    lx = (xs - x0.astype(F32)).astype(F32).reshape(1, 1, out_w, 1)
    tl = x[:, y0][:, :, x0]
    tr = x[:, y0][:, :, x1]
    bl = x[:, y1][:, :, x0]
    br = x[:, y1][:, :, x1]
    top = (tl + (tr - tl) * lx).astype(F32)
    bot = (bl + (br - bl) * lx).astype(F32)
    return (top + (bot - top) * ly).astype(F32)


def resize_nearest_legacy(x, out_h, out_w):
    """tf.image.resize_nearest_neighbor, legacy: src=min(floor(dst*scale), in-1)
    (YOLOv3.py:406; App. A.6)."""
    b, h, w, c = x.shape
    hs = F32(h) / F32(out_h)
    ws = F32(w) / F32(out_w)
    yi = np.minimum(np.floor(np.arange(out_h, dtype=F32) * hs).astype(np.int64), h - 1)
    xi = np.minimum(np.floor(np.arange(out_w, dtype=F32) * ws).astype(np.int64), w - 1)
    return np.ascontiguousarray(x[:, yi][:, :, xi])


def softmax_lastdim(x):
    """tf.nn.softmax (SSD300.py:159; App. A.7): exp(x-max) * (1/sum)."""
    x = x.astype(F32)
    e = np.exp(x - x.max(axis=-1, keepdims=True)).astype(F32)
    s = np.zeros(e.shape[:-1] + (1,), dtype=F32)
    for i in range(e.shape[-1]):  # sequential fp32 sum, index order
        s = (s + e[..., i:i + 1]).astype(F32)
    return (e * (F32(1.0) / s)).astype(F32)


def sigmoid(x):
    """tf.sigmoid = 1/(1+exp(-x)) (YOLOv3.py:338; App. A.9)."""
    x = x.astype(F32)
    return (F32(1.0) / (F32(1.0) + np.exp(-x).astype(F32))).astype(F32)


PI_BIAS = F32(-math.log((1 - 0.01) / 0.01))  # RetinaNet.py:617, FCOS.py:487
RGB_MEAN = np.array([123.68, 116.779, 103.979], dtype=F32)  # SSD300.py:55
