"""Shared test configuration (reference driver defaults)."""
BASE_CFG = {"mode": "test", "data_format": "channels_last", "num_classes": 20, "weight_decay": 1e-4,
            "keep_prob": 0.5, "batch_size": 1, "nms_score_threshold": 0.5, "nms_max_boxes": 20,
            "nms_iou_threshold": 0.5, "pretraining_weight": None}
YOLO_PRIORS = [[[10, 13], [16, 30], [33, 23]], [[30, 61], [62, 45], [59, 119]],
               [[116, 90], [156, 198], [373, 326]]]


def model_cfg(kind, **over):
    c = dict(BASE_CFG)
    if kind == "retinanet":
        c.update(data_shape=[128, 128, 3], is_bottleneck=True, residual_block_list=[3, 4, 6, 3],
                 init_conv_filters=16, is_pretraining=False, gamma=2.0, alpha=0.25,
                 nms_score_threshold=0.8, nms_max_boxes=10, nms_iou_threshold=0.45)
    elif kind == "yolov3":
        c.update(data_shape=[64, 64, 3], coord_scale=1, noobj_scale=1, obj_scale=5, class_scale=1,
                 num_priors=3, priors=YOLO_PRIORS, nms_max_boxes=10, nms_iou_threshold=0.45)
    elif kind == "fcos":
        c.update(data_shape=[128, 128, 3], nms_max_boxes=10, nms_iou_threshold=0.45)
    c.update(over)
    return c


def assert_boxes_close(got, exp, abs_tol=1e-4, rel_tol=1e-6, what="boxes"):
    """Box parity bar: 1e-4 absolute, plus rel_tol x the box's largest coordinate
    magnitude (random-weight boxes reach 1e7 px, where 1 fp32 ulp is ~1 px and
    y1 = cy - h/2 cancels catastrophically; in-image boxes see the bare 1e-4)."""
    import numpy as np
    got, exp = np.asarray(got, np.float32), np.asarray(exp, np.float32)
    assert got.shape == exp.shape, (got.shape, exp.shape)
    if got.size == 0:
        return
    mag = np.abs(exp).max(axis=-1, keepdims=True)
    bad = np.abs(got - exp) > abs_tol + rel_tol * mag
    assert not bad.any(), "%s: %d coords off, worst |err| %.4g at magnitude %.4g" % (
        what, int(bad.sum()), float(np.abs(got - exp)[bad].max()), float(mag[bad.any(-1)].max()))
