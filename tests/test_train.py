"""Training step of SSD (odt_b200/train.py) on the CPU: the differentiable restatement of the reference loss equals
the numpy oracle (oracle/loss.py::ssd_image_loss) on the same rows, its gradient matches finite differences, and one
Momentum step does what `tf.train.MomentumOptimizer(lr, 0.9)` + L2 + the BN update ops do."""
import numpy as np
import pytest
import torch

from helpers import model_cfg


def _gt(rng, size, B=2, G=8):
    gt = np.full((B, G, 5), -1.0, np.float32)
    for b in range(B):
        n = 2 + 2 * b
        gt[b, :n, 0:2] = rng.uniform(0.2 * size, 0.8 * size, (n, 2))
        gt[b, :n, 2:4] = rng.uniform(0.1 * size, 0.5 * size, (n, 2))
        gt[b, :n, 4] = rng.integers(0, 20, n)
    return gt


@pytest.fixture(scope="module")
def trainer():
    import SSD300
    from odt_b200.train import SSDTrainer
    m = SSD300.SSD300(model_cfg("ssd", bn_init="trained"), None)
    return m, SSDTrainer(m, "cpu")


def test_loss_matches_oracle_and_finite_differences(trainer):
    from oracle import loss as OL
    from oracle import tails as OT
    m, tr = trainer
    rng = np.random.default_rng(5)
    tr._shapes = [(38, 38), (19, 19), (10, 10), (5, 5), (5, 5), (3, 3)]
    rows = np.concatenate([rng.standard_normal((2, 8828, 21)) * 2.0, rng.standard_normal((2, 8828, 4)) * 0.5], -1)
    rows = rows.astype(np.float32)
    gt = _gt(rng, 300)
    a1, a2, ayx, ahw = OT.ssd_anchors(300, tr._shapes)
    got_a = [t.numpy() for t in tr.anchors()]
    for g, r in zip(got_a, (a1, a2, ayx, ahw)):
        np.testing.assert_allclose(g, r, rtol=1e-6, atol=1e-4)
    for b in range(2):
        ref, info = OL.ssd_image_loss(rows[b, :, :21], rows[b, :, 21:23], rows[b, :, 23:], a1, a2, ayx, ahw, gt[b])
        val, npos, nneg, nsel = tr.image_loss(torch.from_numpy(rows[b]), gt[b])
        assert (npos, nneg, nsel) == (info["num_pos"], info["num_neg"], info["selected"])
        assert abs(float(val) - ref) <= 2e-5 * max(abs(ref), 1.0)
    # gradient of the per-image loss w.r.t. a few row entries (float64 copy of the same graph)
    tr64 = type(tr)(m, "cpu", dtype=torch.float64)
    tr64._shapes = tr._shapes
    r64 = torch.tensor(rows[0], dtype=torch.float64, requires_grad=True)
    val = tr64.image_loss(r64, gt[0])[0]
    grad = torch.autograd.grad(val, r64)[0]
    nz = torch.nonzero(grad.abs() > 1e-6)
    assert len(nz) > 50
    for idx in nz[:: max(len(nz) // 12, 1)][:12]:
        i, j = int(idx[0]), int(idx[1])
        e = 1e-5
        rp, rm = r64.detach().clone(), r64.detach().clone()
        rp[i, j] += e
        rm[i, j] -= e
        fd = (float(tr64.image_loss(rp, gt[0])[0]) - float(tr64.image_loss(rm, gt[0])[0])) / (2 * e)
        assert abs(fd - float(grad[i, j])) <= 1e-5 + 1e-4 * abs(fd), (i, j, fd, float(grad[i, j]))


def test_one_momentum_step(trainer):
    """First step from zero slots: v1 = v0 - lr * (dL/dv + wd * v0); moving statistics move 1 % towards the batch's."""
    m, _ = trainer
    from odt_b200.train import SSDTrainer
    tr = SSDTrainer(m, "cpu")
    rng = np.random.default_rng(9)
    img = rng.integers(0, 256, (2, 300, 300, 3)).astype(np.float32)
    gt = _gt(rng, 300)
    before = {k: v.detach().clone() for k, v in tr.params.items()}
    mm0 = tr.buffers["feature_extractor/batch_normalization/moving_mean"].clone()
    stats = []
    rows = tr.forward_rows(img, stats)
    loss, data = tr.total_loss(rows, gt)
    grads = torch.autograd.grad(loss, list(tr.params.values()))
    l1 = tr.step(img, gt, 1e-3)
    assert abs(l1 - float(loss)) <= 1e-4 * abs(float(loss)) and np.isfinite(l1)
    for (name, p0), g in zip(before.items(), grads):
        np.testing.assert_allclose(tr.params[name].detach().numpy(), (p0 - 1e-3 * g).numpy(), rtol=1e-5, atol=1e-7)
        np.testing.assert_allclose(tr.slots[name].numpy(), g.numpy(), rtol=1e-6, atol=1e-9)
    # the weight-decay term is part of the loss (SSD300.py:149-151): gradient includes wd * v
    k = "feature_extractor/conv6/kernel"
    wd_part = m.weight_decay * before[k]
    assert float((grads[list(before).index(k)] - wd_part).abs().max()) > 0
    mm1 = tr.buffers["feature_extractor/batch_normalization/moving_mean"]
    np.testing.assert_allclose(mm1.numpy(), (0.99 * mm0 + 0.01 * stats[0][1]).numpy(), rtol=1e-5, atol=1e-7)
    l2 = tr.step(img, gt, 1e-3)
    l3 = tr.step(img, gt, 1e-3)
    assert np.isfinite(l2) and np.isfinite(l3) and tr.global_step == 3
    w = tr.export()
    assert set(w) == set(m.get_weights()) and w[k].dtype == np.float32


def test_train_one_epoch_surface(tmp_path):
    """`train_one_epoch(lr)` on a generator pair like utils.tfrecord_voc_utils.get_generator returns
    (ref SSD300.py:473-484, testSSD300.py:53-66)."""
    import SSD300
    rng = np.random.default_rng(2)

    class It:
        def __init__(self):
            self.n = 0

        def initialize(self):
            self.n = 0

        def get_next(self):
            self.n += 1
            return rng.integers(0, 256, (2, 300, 300, 3)).astype(np.float32), _gt(rng, 300)

    it = It()
    cfg = model_cfg("ssd", mode="train", batch_size=2)
    provider = {"data_shape": [300, 300, 3], "num_train": 4, "num_val": 0, "train_generator": (it.initialize, it),
                "val_generator": None}
    m = SSD300.SSD300(cfg, provider)
    m.device = "cpu"
    w0 = m.get_weights()["regressor/pred1/kernel"].copy()
    loss = m.train_one_epoch(1e-3)
    assert np.isfinite(loss) and it.n == 2 and m.global_step == 2
    assert not np.array_equal(m.get_weights()["regressor/pred1/kernel"], w0)
    out = m.save_weight("latest", str(tmp_path / "ck" / "model"))
    assert out.endswith("model-2")


# ---------------------------------------------------------------- all families through the graph executor ----------
FAMILIES = {"ssd300": ("ssd", {}, 300), "retinanet": ("retinanet", {}, 128), "yolov3": ("yolov3", {}, 64),
            "fcos": ("fcos", {}, 128)}


def _family_model(name, **over):
    import FCOS
    import RetinaNet
    import SSD300
    import YOLOv3
    cls = {"ssd300": SSD300.SSD300, "retinanet": RetinaNet.RetinaNet, "yolov3": YOLOv3.YOLOv3, "fcos": FCOS.FCOS}[name]
    return cls(model_cfg(FAMILIES[name][0], bn_init="trained", **over), None)


@pytest.mark.parametrize("name", ["ssd300", "retinanet", "yolov3", "fcos"])
def test_graph_executor_matches_oracle_forward(name):
    """The engine's layer list (nets.build_*), replayed by train.GraphTrainer with torch ops in INFERENCE mode, gives
    the rows of the independent oracle restatement (oracle/nets.py): pins the executor the training step uses."""
    from oracle import nets as ON
    from oracle import tails as OT
    from odt_b200.train import GraphTrainer
    m = _family_model(name)
    tr = GraphTrainer(m, "cpu")
    size = FAMILIES[name][2]
    img = np.random.default_rng(3).integers(0, 256, (1, size, size, 3)).astype(np.float32)
    with torch.no_grad():
        rows = tr.forward_rows(img, training=False).numpy()
    w = m.get_weights()
    if name == "ssd300":
        ref = OT.ssd_rows(ON.ssd_heads(w, img, 300))
    elif name == "retinanet":
        ref = OT.retina_rows(ON.retinanet_heads(w, img))
    elif name == "yolov3":
        ref = OT.yolo_rows(ON.yolov3_heads(w, img))
    else:
        ref = OT.fcos_rows(ON.fcos_heads(w, img))
    assert rows.shape == ref.shape
    assert np.abs(rows - ref).max() <= 2e-4 * np.abs(ref).max(), (name, np.abs(rows - ref).max(), np.abs(ref).max())


def _rows_and_gt(tr, rng, size, kind):
    N = sum(h * w * a for h, w, a in tr.net.levels)
    rows = np.empty((2, N, 25), np.float32)
    rows[..., :21] = rng.standard_normal((2, N, 21)) * 1.5
    rows[..., 21:] = rng.standard_normal((2, N, 4)) * 0.5
    if kind == "yolo":
        rows[..., 24] = rng.standard_normal((2, N)) + 0.5
    return rows, _gt(rng, size)


def test_retina_loss_matches_oracle():
    from oracle import loss as OL
    from oracle import tails as OT
    from odt_b200.train import GraphTrainer
    m = _family_model("retinanet")
    tr = GraphTrainer(m, "cpu")
    rows, gt = _rows_and_gt(tr, np.random.default_rng(11), 128, "retina")
    a1, a2, ayx, ahw = OT.retina_anchors([128, 128, 3], [(h, w) for h, w, _ in tr.net.levels])
    for b in range(2):
        ref, _ = OL.retina_image_loss(rows[b, :, :21], rows[b, :, 21:23], rows[b, :, 23:], a1, a2, ayx, ahw, gt[b])
        got = float(tr.image_loss(torch.from_numpy(rows[b]), gt[b]))
        assert abs(got - ref) <= 3e-5 * max(abs(ref), 1.0), (got, ref)


def test_fcos_loss_matches_oracle():
    from oracle import loss as OL
    from oracle import tails as OT
    from odt_b200.train import GraphTrainer
    m = _family_model("fcos", data_shape=[256, 256, 3])
    tr = GraphTrainer(m, "cpu")
    rng = np.random.default_rng(12)
    rows, _ = _rows_and_gt(tr, rng, 256, "fcos")
    rows[..., 21:] = rng.standard_normal(rows[..., 21:].shape) * 0.5 + 1.0
    gt = np.full((2, 10, 5), -1.0, np.float32)
    gt[0, :4] = [[60, 70, 40, 50, 3], [128, 128, 64, 64, 7], [120, 130, 200, 180, 11], [125, 125, 100, 90, 7]]
    gt[1, :3] = [[200, 40, 30, 60, 0], [100, 160, 150, 120, 19], [90, 150, 300, 290, 5]]
    heads = OT.rows_to_levels(rows, list(tr.net.levels), "fcos")
    for b in range(2):
        ref = OL.fcos_image_loss(heads, gt[b], image=b)
        got = float(tr.image_loss(torch.from_numpy(rows[b]), gt[b]))
        assert abs(got - ref) <= 3e-5 * max(abs(ref), 1.0), (got, ref)


def test_yolo_loss_matches_oracle():
    from helpers import YOLO_PRIORS
    from oracle import loss as OL
    from oracle import tails as OT
    from odt_b200.train import GraphTrainer
    m = _family_model("yolov3", data_shape=[160, 160, 3], coord_scale=2, noobj_scale=0.5, obj_scale=5, class_scale=1.5)
    tr = GraphTrainer(m, "cpu")
    rng = np.random.default_rng(13)
    rows, _ = _rows_and_gt(tr, rng, 160, "yolo")
    gt = np.full((2, 8, 5), -1.0, np.float32)
    for b in range(2):
        n = 3 + 2 * b
        gt[b, :n, 0:2] = rng.uniform(20, 140, (n, 2))
        gt[b, :n, 2:4] = rng.uniform(10, 120, (n, 2))
        gt[b, :n, 4] = rng.integers(0, 20, n)
    preds = OT.rows_to_levels(rows, list(tr.net.levels), "yolo")
    for b in range(2):
        ref = OL.yolo_image_loss(preds, YOLO_PRIORS, gt[b], coord_scale=2, noobj_scale=0.5, obj_scale=5, class_scale=1.5,
                                 image=b)
        got = float(tr.image_loss(torch.from_numpy(rows[b]), gt[b]))
        assert abs(got - ref) <= 3e-5 * max(abs(ref), 1.0), (got, ref)


@pytest.mark.parametrize("name,size", [("retinanet", 128), ("yolov3", 64), ("fcos", 128)])
def test_family_training_step(name, size):
    """One Momentum step per family: finite loss, every variable with a gradient moves by lr * grad from zero slots,
    BN statistics (where the family has BN) move 1 % towards the batch's."""
    from odt_b200.train import GraphTrainer
    m = _family_model(name, data_shape=[size, size, 3])
    tr = GraphTrainer(m, "cpu")
    rng = np.random.default_rng(21)
    img = rng.integers(0, 256, (2, size, size, 3)).astype(np.float32)
    gt = _gt(rng, size)
    before = {k: v.detach().clone() for k, v in tr.params.items()}
    stats = []
    loss, _ = tr.total_loss(tr.forward_rows(img, True, stats), gt)
    names = list(tr.params)
    grads = torch.autograd.grad(loss, [tr.params[k] for k in names], allow_unused=True)
    buf0 = {k: v.clone() for k, v in tr.buffers.items()}
    l1 = tr.step(img, gt, 1e-3)
    assert np.isfinite(l1) and abs(l1 - float(loss)) <= 1e-4 * max(abs(float(loss)), 1.0)
    moved = 0
    for k, g in zip(names, grads):
        if g is None:
            continue
        np.testing.assert_allclose(tr.params[k].detach().numpy(), (before[k] - 1e-3 * g).numpy(), rtol=2e-5, atol=1e-7)
        moved += 1
    assert moved >= 0.95 * len(names)
    if stats:
        scope, mean, _ = stats[0]
        np.testing.assert_allclose(tr.buffers[scope + "/moving_mean"].numpy(),
                                   (0.99 * buf0[scope + "/moving_mean"] + 0.01 * mean).numpy(), rtol=1e-5, atol=1e-7)
    assert np.isfinite(tr.step(img, gt, 1e-3))
