"""TFRecord / tf.train.Example / VOC record schema, CPU only.  PARITY UNPINNED (no
TensorFlow, no TF-written file): framing and protobuf layout are checked on hand-assembled bytes and
round trips."""
import struct

import numpy as np
import pytest


@pytest.fixture()
def tr():
    from odt_b200 import tfrecord
    return tfrecord


def _varint(v):
    out = bytearray()
    while v >= 0x80:
        out.append((v & 0x7F) | 0x80)
        v >>= 7
    out.append(v)
    return bytes(out)


def test_framing_bytes_and_corruption(tr, tmp_path):
    from odt_b200.tf_checkpoint import crc32c, mask_crc
    p = str(tmp_path / "a.tfrecord")
    tr.write_records(p, [b"hello", b"", b"x" * 1000])
    raw = open(p, "rb").read()
    # first record by hand: u64 5 | masked crc of those 8 bytes | payload | masked crc of payload
    head = struct.pack("<Q", 5)
    assert raw[:8] == head and raw[8:12] == struct.pack("<I", mask_crc(crc32c(head)))
    assert raw[12:17] == b"hello" and raw[17:21] == struct.pack("<I", mask_crc(crc32c(b"hello")))
    assert list(tr.read_records(p)) == [b"hello", b"", b"x" * 1000]
    bad = bytearray(raw)
    bad[14] ^= 1
    open(p, "wb").write(bytes(bad))
    with pytest.raises(tr.RecordError, match="payload checksum"):
        list(tr.read_records(p))


def test_example_protobuf_layout_by_hand(tr):
    """Example{features{feature{key:'k' value{bytes_list{value:'ab'}}}}} assembled byte by byte."""
    bytes_list = b"\x0a\x02ab"                       # field 1, len 2
    feature = b"\x0a" + _varint(len(bytes_list)) + bytes_list   # Feature.bytes_list = field 1
    entry = b"\x0a\x01k" + b"\x12" + _varint(len(feature)) + feature
    features = b"\x0a" + _varint(len(entry)) + entry
    example = b"\x0a" + _varint(len(features)) + features
    assert tr.parse_example(example) == {"k": [b"ab"]}
    assert tr.encode_example({"k": b"ab"}) == example
    # float_list (packed) and int64_list (packed varints, negative value)
    ex = tr.parse_example(tr.encode_example({"f": np.asarray([1.5, -2.0], np.float32), "i": np.asarray([3, -1])}))
    np.testing.assert_array_equal(ex["f"], np.asarray([1.5, -2.0], np.float32))
    np.testing.assert_array_equal(ex["i"], np.asarray([3, -1], np.int64))


def test_voc_records_through_get_generator(tr, tmp_path):
    import cv2
    from utils import tfrecord_voc_utils as voc_utils
    rng = np.random.default_rng(0)
    recs = []
    for i in range(5):
        h, w = 40 + 4 * i, 60
        img = np.zeros((h, w, 3), np.uint8)
        img[:, : w // 2] = (200, 30, 30)             # left half red (RGB), smooth -> JPEG-stable
        ok, enc = cv2.imencode(".jpg", cv2.cvtColor(img, cv2.COLOR_RGB2BGR))
        assert ok
        gt = np.asarray([[4, 24, 6, 36, i % 20], [10, 30, 20, 50, (i + 1) % 20]], np.float32)
        recs.append(tr.encode_voc_example(enc.tobytes(), [h, w, 3], gt))
    path = str(tmp_path / "voc_00001-of-00001.tfrecord")
    tr.write_records(path, recs)
    img0, gt0 = tr.decode_voc_example(next(tr.read_records(path)))
    assert img0.shape == (40, 60, 3) and img0[5, 5, 0] > 150 and img0[5, 55, 0] < 40   # RGB order kept
    cfg = {"data_format": "channels_last", "output_shape": [20, 30], "zoom_size": None, "crop_method": None,
           "flip_prob": None, "fill_mode": "BILINEAR", "keep_aspect_ratios": False, "constant_values": 0.,
           "color_jitter_prob": None, "rotate": None, "pad_truth_to": 6}
    init_op, it = voc_utils.get_generator([path], 2, 3, cfg)
    init_op()
    images, truth = it.get_next()
    assert images.shape == (2, 20, 30, 3) and images.dtype == np.float32
    assert truth.shape == (2, 6, 5) and np.all(truth[:, 2:] == -1)
    # the 40x60 record at half scale: box (ymin 4, ymax 24, xmin 6, xmax 36) -> centre (7, 10.5), size (10, 15)
    seen = False
    for _ in range(6):
        for t in truth.reshape(-1, 5):
            if t[4] == 0 and abs(t[2] - 10) < 1e-4:
                np.testing.assert_allclose(t[:4], [7.0, 10.5, 10.0, 15.0], atol=1e-4)
                seen = True
        images, truth = it.get_next()
    assert seen
    for _ in range(4):   # repeat(): the stream never ends
        it.get_next()


def test_dataset2tfrecord_from_voc_xml(tr, tmp_path):
    import cv2
    from utils import tfrecord_voc_utils as voc_utils
    xml_dir, img_dir, out_dir = tmp_path / "Annotations", tmp_path / "JPEGImages", tmp_path / "out"
    xml_dir.mkdir()
    img_dir.mkdir()
    for i in range(4):
        img = np.full((30, 50, 3), 40 * i, np.uint8)
        cv2.imwrite(str(img_dir / ("im%d.jpg" % i)), img)
        (xml_dir / ("im%d.xml" % i)).write_text(
            "<annotation><filename>im%d.jpg</filename><size><width>50</width><height>30</height><depth>3</depth></size>"
            "<object><name>dog</name><bndbox><xmin>5</xmin><ymin>6</ymin><xmax>25</xmax><ymax>16</ymax></bndbox></object>"
            "<object><name>person</name><bndbox><xmin>1</xmin><ymin>2</ymin><xmax>40</xmax><ymax>28</ymax></bndbox></object>"
            "</annotation>" % i)
    files = voc_utils.dataset2tfrecord(str(xml_dir), str(img_dir), str(out_dir), "voc", total_shards=2)
    assert [f.split("/")[-1] for f in files] == ["voc_00001-of-00002.tfrecord", "voc_00002-of-00002.tfrecord"]
    recs = [r for f in files for r in tr.read_records(f)]
    assert len(recs) == 4
    from utils.voc_classname_encoder import classname_to_ids
    img, gt = tr.decode_voc_example(recs[0])
    assert img.shape == (30, 50, 3)
    np.testing.assert_array_equal(gt, np.asarray([[6, 16, 5, 25, classname_to_ids["dog"]],
                                                  [2, 28, 1, 40, classname_to_ids["person"]]], np.float32))


def test_preprocess_zoom_crop_flip_clamp_by_hand(tr):
    """image_augmentor arithmetic on a 10x10 image: resize to zoom 8x8 (ratio .8), centre crop to 6x6 (offset 1),
    left-right flip (prob 1) with the reference's `- 1` offsets, clamping to [0, 5], centre filter, centre form."""
    img = np.arange(300, dtype=np.float32).reshape(10, 10, 3) % 251
    gt = np.asarray([[2, 6, 0, 10, 3],      # -> (1.6, 4.8, 0, 8) -> crop (.6, 3.8, -1, 7) -> flip x (-2, 6) -> clamp (0, 5)
                     [0, 1, 0, 1, 5]], np.float32)   # -> ends up with its centre on the border: dropped
    cfg = {"data_format": "channels_last", "output_shape": [6, 6], "zoom_size": [8, 8], "crop_method": "center",
           "flip_prob": [0.0, 1.0], "fill_mode": "BILINEAR", "keep_aspect_ratios": False, "constant_values": 0.,
           "color_jitter_prob": None, "rotate": None, "pad_truth_to": 3}
    out, box = tr.preprocess(img, gt, cfg, np.random.default_rng(0))
    assert out.shape == (6, 6, 3)
    np.testing.assert_allclose(box[0], [2.2, 2.5, 3.2, 5.0, 3.0], atol=1e-5)
    assert np.all(box[1:] == -1)
    # the image itself: aligned resize 10 -> 8 then rows/cols 1..6, then mirrored
    full = tr._resize_bilinear_aligned(img, 8, 8)
    np.testing.assert_allclose(out, full[1:7, 1:7][:, ::-1], atol=1e-5)
    # align_corners: the corners of the resized image are the corners of the source
    np.testing.assert_allclose(full[0, 0], img[0, 0])
    np.testing.assert_allclose(full[7, 7], img[9, 9])
    # aspect-preserving path: 10x20 image into an 8x8 canvas -> 4x8 content + constant padding
    wide = np.ones((10, 20, 3), np.float32) * 7
    cfg2 = dict(cfg, zoom_size=None, output_shape=[8, 8], keep_aspect_ratios=True, constant_values=0.5, flip_prob=None)
    o2, b2 = tr.preprocess(wide, np.asarray([[2, 8, 4, 16, 1]], np.float32), cfg2)
    assert np.all(o2[:4] == 7) and np.all(o2[4:] == 0.5)
    np.testing.assert_allclose(b2[0], [2.0, 4.0, 2.4, 4.8, 1.0], atol=1e-5)


def test_colour_adjustments_match_their_definitions(tr):
    import colorsys
    rng = np.random.default_rng(5)
    img = rng.uniform(0, 255, (6, 7, 3)).astype(np.float32)
    np.testing.assert_allclose(tr.adjust_brightness(img, 0.25), img + 0.25, rtol=0, atol=1e-5)
    mean = img.reshape(-1, 3).mean(0)
    np.testing.assert_allclose(tr.adjust_contrast(img, 1.15), (img - mean) * 1.15 + mean, rtol=1e-5, atol=1e-3)
    # hue rotation == HSV round trip with h shifted (value and saturation kept)
    for delta in (0.07, -0.1, 0.0):
        got = tr.adjust_hue(img, delta)
        for (y, x) in ((0, 0), (3, 4), (5, 6), (2, 1)):
            hh, ss, vv = colorsys.rgb_to_hsv(*(img[y, x].astype(np.float64) / 255.0))
            want = np.asarray(colorsys.hsv_to_rgb((hh + delta) % 1.0, ss, vv)) * 255.0
            np.testing.assert_allclose(got[y, x], want, rtol=0, atol=2e-3)
    # grey pixels have no hue: unchanged; min / max of every pixel are preserved
    grey = np.full((2, 2, 3), 77.0, np.float32)
    np.testing.assert_array_equal(tr.adjust_hue(grey, 0.05), grey)
    got = tr.adjust_hue(img, 0.09)
    np.testing.assert_allclose(got.max(-1), img.max(-1), rtol=0, atol=1e-4)
    np.testing.assert_allclose(got.min(-1), img.min(-1), rtol=0, atol=1e-4)


def test_rotation_image_and_boxes_agree(tr):
    rng = np.random.default_rng(6)
    img = rng.uniform(0, 255, (9, 9, 3)).astype(np.float32)
    np.testing.assert_allclose(tr.rotate_bilinear(img, 0.0), img, rtol=0, atol=1e-4)
    # a quarter turn is exact and counter-clockwise (tf.contrib.image.rotate's convention)
    np.testing.assert_allclose(tr.rotate_bilinear(img, np.pi / 2), np.rot90(img), rtol=0, atol=2e-3)
    # the box re-fit follows the pixels: pixel (r, c) lands on (n-1-c, r) under the quarter turn
    n = 9
    ymin, xmin, ymax, xmax = [np.asarray([v], np.float32) for v in (1.0, 2.0, 3.0, 6.0)]
    y0, x0, y1, x1 = tr.rotate_boxes(np.pi / 2, ymin, xmin, ymax, xmax, float(n), float(n))
    np.testing.assert_allclose([y0[0], x0[0], y1[0], x1[0]], [n - 1 - 6.0, 1.0, n - 1 - 2.0, 3.0], atol=1e-4)
    # small angle: a bright blob stays inside the re-fitted box of the box that framed it
    big = np.zeros((41, 61, 3), np.float32)
    big[10:15, 40:47] = 255.0
    ang = np.float32(4.0 * np.pi / 180)
    out = tr.rotate_bilinear(big, ang)
    yy, xx = np.nonzero(out[..., 0] > 128)
    b = tr.rotate_boxes(ang, *[np.asarray([v], np.float32) for v in (10.0, 40.0, 14.0, 46.0)], 41.0, 61.0)
    assert b[0][0] - 0.5 <= yy.min() and yy.max() <= b[2][0] + 0.5
    assert b[1][0] - 0.5 <= xx.min() and xx.max() <= b[3][0] + 0.5
    # pixels rotated in from outside read 0
    assert out[0, 0, 0] == 0.0 or out[-1, -1, 0] == 0.0


def test_preprocess_jitter_rotate_and_fallback(tr):
    rng = np.random.default_rng(7)
    img = rng.integers(0, 256, (50, 80, 3)).astype(np.uint8)
    gt = np.asarray([[10, 30, 20, 60, 3], [5, 45, 5, 75, 7]], np.float32)   # ymin ymax xmin xmax id
    base = dict(data_format="channels_last", output_shape=[40, 40], pad_truth_to=4)
    plain, box0 = tr.preprocess(img, gt, base, np.random.default_rng(0))
    cfg = dict(base, color_jitter_prob=1.0, rotate=[1.0, -5.0, 5.0])
    out, box = tr.preprocess(img, gt, cfg, np.random.default_rng(0))
    assert out.shape == (40, 40, 3) and box.shape == (4, 5) and out.dtype == np.float32
    assert np.abs(out - plain).max() > 1.0                       # the image really changed
    assert (box[:2, 4] == [3, 7]).all() and (box[2:] == -1).all()
    assert np.abs(box[:2, :4] - box0[:2, :4]).max() < 8.0        # a <=5 degree turn moves boxes a little
    assert (box[:2, 2:4] >= box0[:2, 2:4] - 1e-3).all()          # and the hull never shrinks
    # same seed -> same result; probabilities 0 -> identical to the plain path
    out2, box2 = tr.preprocess(img, gt, cfg, np.random.default_rng(0))
    np.testing.assert_array_equal(out, out2)
    np.testing.assert_array_equal(box, box2)
    off, boxo = tr.preprocess(img, gt, dict(base, color_jitter_prob=0.0, rotate=[0.0, -5.0, 5.0]), np.random.default_rng(0))
    np.testing.assert_array_equal(off, plain)
    np.testing.assert_array_equal(boxo, box0)
    with pytest.raises(ValueError):
        tr.preprocess(img, gt, dict(base, rotate=[0.5, 3.0, -3.0]), np.random.default_rng(0))
    # nearest-neighbour resampling picks source pixels only
    nn, _ = tr.preprocess(img, gt, dict(base, fill_mode="NEAREST_NEIGHBOR"), None)
    assert set(np.unique(nn)).issubset(set(np.unique(img).astype(np.float32)))
    assert nn[0, 0, 0] == img[0, 0, 0] and nn[-1, -1, 2] == img[-1, -1, 2]
    # no box keeps its centre inside after a crop -> the un-augmented, plainly resized example is returned
    far = np.asarray([[0, 4, 0, 6, 2]], np.float32)
    z, bz = tr.preprocess(img, far, dict(base, zoom_size=[100, 160], crop_method="center"), None)
    np.testing.assert_allclose(z, tr._resize_bilinear_legacy(img, 40, 40))
    np.testing.assert_allclose(bz[0], [2 * 0.8, 3 * 0.5, 4 * 0.8, 6 * 0.5, 2], rtol=1e-6)
    assert (bz[1:] == -1).all()


def test_corrupt_records_raise_record_error_only(tr, tmp_path):
    rng = np.random.default_rng(12)
    good = tr.encode_voc_example(b"\xff\xd8abc", np.asarray([3, 4, 3], np.int32), np.zeros((2, 5), np.float32))
    p = str(tmp_path / "a.tfrecord")
    tr.write_records(p, [b"hello", b"x" * 100])
    raw = open(p, "rb").read()

    def mutate(src):
        b = bytearray(src)
        for _ in range(int(rng.integers(1, 4))):
            b[int(rng.integers(len(b)))] = int(rng.integers(256))
        return bytes(b[:int(rng.integers(len(b)))] if rng.random() < 0.3 else b)

    for _ in range(400):
        ex = mutate(good)
        open(p, "wb").write(mutate(raw))
        for fn in (lambda: tr.parse_example(ex), lambda: tr.decode_voc_example(ex),
                   lambda: list(tr.read_records(p)), lambda: list(tr.read_records(p, verify=False))):
            try:
                fn()
            except tr.RecordError.__mro__[1]:      # CheckpointError (RecordError derives from it)
                pass


def test_bicubic_resize_restatement():
    """fill_mode='BICUBIC' (utils/image_augmentor.py:70-74,103-106): TF 1.13 ResizeBicubic, align_corners=True --
    Keys kernel A=-0.75 from a 1024-step table, border taps clamped.  Pins that need no TensorFlow: weights sum to 1,
    identity at equal size, exact hits on integer source positions, a constant stays constant, cubic accuracy on a
    smooth ramp, separability (x then y), and the call path through preprocess()."""
    from odt_b200 import tfrecord as R
    t = R._bicubic_table()
    i = np.arange(1025)
    np.testing.assert_allclose(t[i * 2 + 1] + t[i * 2] + t[(1024 - i) * 2] + t[(1024 - i) * 2 + 1], 1.0, atol=2e-6)
    assert t[0] == 1.0 and abs(t[1]) < 1e-7 and abs(t[2048]) < 1e-7          # distance 0 / 1 / 1
    assert abs(t[1024] - 0.59375) < 1e-7 and abs(t[1025] + 0.09375) < 1e-7   # x = 0.5 with A = -0.75: (19/32, -3/32)
    rng = np.random.default_rng(4)
    img = rng.uniform(0, 255, (9, 13, 3)).astype(np.float32)
    np.testing.assert_allclose(R._resize_bicubic_aligned(img, 9, 13), img, atol=1e-4)   # scale 1: offsets 0
    up = R._resize_bicubic_aligned(img, 17, 25)                                          # scale exactly 1/2
    np.testing.assert_allclose(up[::2, ::2], img, atol=1e-4)
    mid = up[2, 1::2]                                                                    # row 1, half-way columns
    c = np.clip(np.arange(12)[:, None] + np.array([-1, 0, 1, 2])[None], 0, 12)
    exp = (img[1][c] * np.array([-0.09375, 0.59375, 0.59375, -0.09375], np.float32)[None, :, None]).sum(1)
    np.testing.assert_allclose(mid, exp, atol=1e-3)
    flat = np.full((5, 7, 3), 42.0, np.float32)
    np.testing.assert_allclose(R._resize_bicubic_aligned(flat, 11, 4), 42.0, atol=1e-4)
    ramp = (np.arange(20, dtype=np.float32)[:, None, None] * 3.0 + np.arange(30, dtype=np.float32)[None, :, None] * 2.0)
    out = R._resize_bicubic_aligned(ramp, 33, 47)
    yy = np.arange(33, dtype=np.float32) * (19 / 32)
    xx = np.arange(47, dtype=np.float32) * (29 / 46)
    lin = yy[:, None] * 3.0 + xx[None] * 2.0
    # away from the clamped border taps the A = -0.75 kernel stays within ~7 % of a sample step of linear data
    # (only A = -0.5 reproduces it exactly): slopes 3 and 2 per sample
    np.testing.assert_allclose(out[3:-3, 3:-3, 0], lin[3:-3, 3:-3], atol=0.4)
    cfg = {"output_shape": [24, 24], "fill_mode": "BICUBIC"}
    gt = np.array([[2, 10, 3, 12, 5]], np.float32)
    a, _ = R.preprocess(img, gt, dict(cfg), np.random.default_rng(0))
    b, _ = R.preprocess(img, gt, dict(cfg, fill_mode="BILINEAR"), np.random.default_rng(0))
    assert a.shape == (24, 24, 3) and np.abs(a - b).max() > 1e-3      # a different resampler really ran
    np.testing.assert_allclose(a, R._resize_bicubic_aligned(img, 24, 24), atol=1e-5)
