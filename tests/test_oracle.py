"""CPU suite, part 1: pin the oracle against the hand-derived known answers of
SURVEY.md section 8(c) / Appendix A-B (the reference ships no golden vectors:
parity is otherwise unpinned, see oracle/__init__.py) and against the committed
fixtures in tests/golden/."""
import os

import numpy as np
import pytest

from oracle import tails as OT
from oracle import tfops as T

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def test_same_pad_examples():
    # App. A.1 worked examples
    assert T.same_pad(19, 3, 2) == (10, 1, 1)
    assert T.same_pad(10, 3, 2) == (5, 0, 1)
    assert T.same_pad(5, 3, 2) == (3, 1, 1)
    assert T.same_pad(800, 7, 2) == (400, 2, 3)
    assert T.same_pad(75, 2, 2) == (38, 0, 1)
    assert T.same_pad(400, 3, 2) == (200, 0, 1)
    assert T.same_pad(19, 3, 1, 2) == (19, 2, 2)


def test_ssd300_scales_and_anchor_known_answers():
    s = OT.ssd_scales(300)
    exp = [[60, 78.2304], [102, 121.1941], [144, 163.6582], [186, 205.932], [228, 248.1129],
           [270, 290.2413]]
    assert np.allclose(s, exp, atol=1e-3)
    shapes = [(38, 38), (19, 19), (10, 10), (5, 5), (5, 5), (3, 3)]
    y1x1, y2x2, yx, hw = OT.ssd_anchors(300, shapes)
    assert yx.shape[0] == 8828  # NOT 8732: conv10_2 is SAME-padded (App. C)
    np.testing.assert_allclose(y1x1[0], [-26.052631, -26.052631], rtol=0, atol=2e-6)
    np.testing.assert_allclose(y2x2[0], [33.94737, 33.94737], atol=4e-6)
    np.testing.assert_allclose(yx[0], [3.947369, 3.947369], atol=1e-6)
    np.testing.assert_allclose(hw[0], [60, 60], atol=1e-5)
    np.testing.assert_allclose(hw[2], [84.85281, 42.426407], atol=1e-5)  # tall first
    np.testing.assert_allclose(yx[4], [3.947369, 11.842106], atol=2e-6)
    np.testing.assert_allclose(yx[5775], [296.05264, 296.05264], atol=4e-5)
    np.testing.assert_allclose(hw[5775], [42.42639, 84.85283], atol=4e-5)


def test_ssd512_scales_and_count():
    exp = [[35.84, 52.4644], [76.8, 108.6116], [153.6, 188.1208], [230.4, 266.043],
           [307.2, 343.46], [384, 420.6509], [460.8, 497.7209]]
    assert np.allclose(OT.ssd_scales(512), exp, atol=1e-3)
    shapes = [(64, 64), (32, 32), (16, 16), (8, 8), (8, 8), (4, 4), (2, 2)]
    assert OT.ssd_anchors(512, shapes)[2].shape[0] == 24912


def test_retinanet_priors_and_count():
    p = OT.retina_priors(32)
    exp = [(32, 32), (40.317474, 40.317474), (50.796833, 50.796833), (22.627417, 45.254833),
           (28.508759, 57.017517), (35.918785, 71.83757), (45.254833, 22.627417),
           (57.017517, 28.508759), (71.83757, 35.918785)]
    np.testing.assert_allclose(p, exp, rtol=1e-6)
    shapes = [(100, 100), (50, 50), (25, 25), (13, 13), (7, 7)]
    assert OT.retina_anchors([800, 800, 3], shapes)[2].shape[0] == 120087


def test_nms_c_matches_python_statement():
    rng = np.random.default_rng(5)
    for n in (0, 1, 2, 17, 200):
        c = rng.uniform(0, 100, (n, 2)).astype(np.float32)
        hw = rng.uniform(5, 60, (n, 2)).astype(np.float32)
        boxes = np.concatenate([c - hw / 2, c + hw / 2], 1).astype(np.float32)
        scores = rng.permutation(n).astype(np.float32) / max(n, 1)
        for thr in (0.3, 0.5):
            a = OT.nms_c(boxes, scores, 20, thr)
            b = OT.nms_py(boxes, scores, 20, thr)
            assert np.array_equal(a, b)


def test_nms_semantics_edge_cases():
    # strict '>' : IoU exactly equal to the threshold does NOT suppress
    boxes = np.array([[0, 0, 2, 2], [0, 1, 2, 3]], np.float32)  # IoU = 2/6 = 1/3
    scores = np.array([0.9, 0.8], np.float32)
    iou = float(np.float32(2) / np.float32(6))
    assert list(OT.nms_c(boxes, scores, 10, iou)) == [0, 1]
    assert list(OT.nms_c(boxes, scores, 10, np.nextafter(np.float32(iou), np.float32(0)))) == [0]
    # zero-area boxes never suppress / are never suppressed
    boxes = np.array([[1, 1, 1, 5], [1, 1, 1, 5], [0, 0, 4, 4]], np.float32)
    assert list(OT.nms_c(boxes, np.array([0.5, 0.4, 0.3], np.float32), 10, 0.1)) == [0, 1, 2]
    # flipped corners are normalised
    boxes = np.array([[4, 4, 0, 0], [0, 0, 4, 4]], np.float32)
    assert list(OT.nms_c(boxes, np.array([0.9, 0.8], np.float32), 10, 0.5)) == [0]
    # max_output_size counts SELECTED boxes
    boxes = np.array([[0, 0, 1, 1], [10, 10, 11, 11], [20, 20, 21, 21]], np.float32)
    assert list(OT.nms_c(boxes, np.array([0.1, 0.3, 0.2], np.float32), 2, 0.5)) == [1, 2]


def test_resize_legacy():
    x = np.arange(4, dtype=np.float32).reshape(1, 2, 2, 1)
    y = T.resize_bilinear_legacy(x, 4, 4)[0, :, :, 0]
    # src = dst * 0.5 ; last row/col clamp to the edge (no half-pixel centres)
    exp = np.array([[0, .5, 1, 1], [1, 1.5, 2, 2], [2, 2.5, 3, 3], [2, 2.5, 3, 3]], np.float32)
    np.testing.assert_allclose(y, exp)
    z = T.resize_nearest_legacy(x, 4, 4)[0, :, :, 0]
    np.testing.assert_allclose(z, [[0, 0, 1, 1], [0, 0, 1, 1], [2, 2, 3, 3], [2, 2, 3, 3]])
    # odd target (RetinaNet 13 -> 25)
    x = np.random.default_rng(0).standard_normal((1, 13, 13, 2)).astype(np.float32)
    y = T.resize_bilinear_legacy(x, 25, 25)
    assert y.shape == (1, 25, 25, 2)
    np.testing.assert_allclose(y[0, 0, 0], x[0, 0, 0])


def test_softmax_and_argmax_conventions():
    x = np.array([[1, 1, 0], [0, 2, 2]], np.float32)
    p = T.softmax_lastdim(x)
    assert np.argmax(p[0]) == 0 and np.argmax(p[1]) == 1  # first maximum wins
    np.testing.assert_allclose(p.sum(-1), 1, atol=1e-6)


def test_group_norm_and_bn():
    rng = np.random.default_rng(1)
    x = rng.standard_normal((2, 5, 7, 16)).astype(np.float32)
    g = rng.uniform(.5, 1.5, 16).astype(np.float32)
    b = rng.standard_normal(16).astype(np.float32)
    y = T.group_norm(x, g, b)
    xg = x.reshape(2, 5, 7, 8, 2)
    ref = (xg - xg.mean((1, 2, 4), keepdims=True)) / np.sqrt(xg.var((1, 2, 4), keepdims=True) + 1e-6)
    ref = ref.reshape(x.shape) * g + b
    np.testing.assert_allclose(y, ref, atol=2e-5)
    m, v = rng.standard_normal(16).astype(np.float32), rng.uniform(.5, 1.5, 16).astype(np.float32)
    z = T.batch_norm_inference(x, g, b, m, v)
    np.testing.assert_allclose(z, (x - m) / np.sqrt(v + 1e-3) * g + b, atol=1e-5)


def test_conv_same_padding_asymmetric():
    # 3x3 stride-2 on 10 -> 5 pads (0,1): first output reads rows 0..2
    x = np.zeros((1, 10, 10, 1), np.float32)
    x[0, 0, 0, 0] = 1
    k = np.zeros((3, 3, 1, 1), np.float32)
    k[0, 0, 0, 0] = 7
    y = T.conv2d_same(x, k, None, stride=2)
    assert y.shape == (1, 5, 5, 1) and y[0, 0, 0, 0] == 7
    # dilation 2 on 19 pads (2,2)
    x = np.zeros((1, 19, 19, 1), np.float32)
    x[0, 0, 0, 0] = 1
    k = np.zeros((3, 3, 1, 1), np.float32)
    k[1, 1, 0, 0] = 3
    y = T.conv2d_same(x, k, None, dil=2)
    assert y.shape == (1, 19, 19, 1) and y[0, 0, 0, 0] == 3


@pytest.mark.parametrize("name", ["tail_ssd", "tail_retina", "tail_yolo", "tail_fcos"])
def test_oracle_reproduces_committed_golden(name):
    """Fixtures written by tests/golden/make_golden.py (oracle outputs frozen at
    commit time): guards the oracle itself against silent drift."""
    from golden import make_golden as mg
    f = os.path.join(GOLD, name + ".npz")
    assert os.path.exists(f), "run python tests/golden/make_golden.py"
    d = np.load(f)
    out = mg.run_case(name, d["rows"])
    np.testing.assert_array_equal(out[2], d["class_id"])
    np.testing.assert_array_equal(out[3], d["keep"])
    np.testing.assert_allclose(out[0], d["scores"], rtol=1e-6)
    np.testing.assert_allclose(out[1], d["bbox"], rtol=1e-5, atol=1e-4)


@pytest.mark.parametrize("name", ["loss_ssd", "loss_retina", "loss_yolo", "loss_fcos"])
def test_loss_oracle_reproduces_committed_golden(name):
    """Frozen loss values of the four training-loss oracles on seeded rows (tests/golden/make_golden.py)."""
    from golden import make_golden as mg
    d = np.load(os.path.join(GOLD, name + ".npz"))
    for b in range(d["rows"].shape[0]):
        got = mg.run_loss_case(name, d["rows"], d["gt"], b)
        assert abs(got - d["loss"][b]) <= 1e-9 * max(abs(d["loss"][b]), 1.0), (name, b, got, d["loss"][b])


# ---- independent implementations (not TensorFlow, but not ours either) ------------------------------------
def test_nms_agrees_with_torchvision_on_tie_free_boxes():
    """Greedy IoU suppression (strict >) is what torchvision.ops.nms implements too; on tie-free random boxes the
    first max_out survivors must coincide with the TF-semantics oracle (C and pure-python statements)."""
    import torch
    tv = pytest.importorskip("torchvision")
    rng = np.random.default_rng(42)
    for n, thr, max_out in ((50, 0.5, 20), (400, 0.45, 10), (1500, 0.3, 200), (7, 0.5, 20)):
        yx = rng.uniform(0, 250, (n, 2))
        hw = rng.uniform(8, 90, (n, 2))
        boxes = np.concatenate([yx, yx + hw], 1).astype(np.float32)          # (y1, x1, y2, x2)
        scores = rng.permutation(n).astype(np.float32) / n + rng.uniform(0, 1e-4, n).astype(np.float32)
        assert len(np.unique(scores)) == n
        keep = tv.ops.nms(torch.from_numpy(boxes[:, [1, 0, 3, 2]].copy()), torch.from_numpy(scores), thr).numpy()
        np.testing.assert_array_equal(OT.nms_c(boxes, scores, max_out, thr), keep[:max_out])
        np.testing.assert_array_equal(OT.nms_py(boxes, scores, max_out, thr), keep[:max_out])


def test_tfops_agree_with_torch_functional():
    import torch
    import torch.nn.functional as F
    rng = np.random.default_rng(43)
    x = rng.standard_normal((2, 6, 9, 16)).astype(np.float32)
    g, b = rng.uniform(.5, 1.5, 16).astype(np.float32), rng.standard_normal(16).astype(np.float32)
    xt = torch.from_numpy(x).permute(0, 3, 1, 2)
    ref = F.group_norm(xt, 8, torch.from_numpy(g), torch.from_numpy(b), eps=1e-6).permute(0, 2, 3, 1).numpy()
    np.testing.assert_allclose(T.group_norm(x, g, b), ref, atol=2e-5)
    np.testing.assert_allclose(T.softmax_lastdim(x), F.softmax(torch.from_numpy(x), -1).numpy(), atol=1e-6)
    np.testing.assert_allclose(T.sigmoid(x), torch.sigmoid(torch.from_numpy(x)).numpy(), atol=1e-6)
    np.testing.assert_allclose(T.leaky_relu(x), F.leaky_relu(torch.from_numpy(x), 0.1).numpy(), atol=0)
    np.testing.assert_allclose(T.l2_normalize_channels(x), F.normalize(torch.from_numpy(x), dim=-1, eps=1e-6).numpy(),
                               atol=1e-6)
    # even sizes, window 2 stride 2: SAME needs no padding, so it must equal the plain max-pool
    ref = F.max_pool2d(torch.from_numpy(x[:, :, :8]).permute(0, 3, 1, 2), 2, 2).permute(0, 2, 3, 1).numpy()
    np.testing.assert_array_equal(T.max_pool_same(x[:, :, :8], 2, 2), ref)
    # nearest x2 (YOLOv3's top-down path) has one definition
    ref = F.interpolate(xt, scale_factor=2, mode="nearest").permute(0, 2, 3, 1).numpy()
    np.testing.assert_array_equal(T.resize_nearest_legacy(x, 12, 18), ref)
    # TF1's legacy bilinear (no half-pixel centres) at integer x2 == align_corners=False of neither torch mode;
    # it does agree with torch where both sample grid points exactly: the even output positions are the inputs
    up = T.resize_bilinear_legacy(x, 12, 18)
    np.testing.assert_allclose(up[:, ::2, ::2], x, atol=1e-6)
