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
