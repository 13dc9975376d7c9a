"""TF V2 checkpoint (TensorBundle) reader / writer: SURVEY §8(f) row 1.  CPU only.
PARITY UNPINNED (no TensorFlow, no TF-written checkpoint here): round trips through the
module's own writer, hand-assembled table blocks, and published known answers of the
building blocks (CRC-32C test vector, LevelDB crc mask, snappy literal/copy decoding)."""
import os
import struct

import numpy as np
import pytest

from helpers import model_cfg


@pytest.fixture()
def ck():
    from odt_b200 import tf_checkpoint
    return tf_checkpoint


def test_crc32c_known_answers(ck):
    assert ck.crc32c(b"123456789") == 0xE3069283            # the standard check value
    assert ck.crc32c(b"") == 0
    assert ck.crc32c(bytes(32)) == 0x8A9136AA               # RFC 3720 B.4: 32 zero bytes
    assert ck.crc32c(bytes([0xFF] * 32)) == 0x62A8AB43      # RFC 3720 B.4: 32 x 0xFF
    assert ck.crc32c(b"6789", ck.crc32c(b"12345")) == 0xE3069283   # continuation
    c = 0x8A9136AA
    assert ck.mask_crc(c) == (((c >> 15) | (c << 17)) + 0xA282EAD8) & 0xFFFFFFFF


def test_python_crc_fallback_matches_native(ck, monkeypatch):
    data = np.random.default_rng(0).integers(0, 256, 4099, dtype=np.uint8).tobytes()
    native = ck.crc32c(data)
    monkeypatch.setattr(ck, "_crc_native", False)
    assert ck.crc32c(data) == native


def test_round_trip_all_dtypes_and_many_blocks(ck, tmp_path):
    rng = np.random.default_rng(1)
    t = {"feature_extractor/kernel_conv1_1": rng.standard_normal((3, 3, 3, 64)).astype(np.float32),
         "feature_extractor/bias_conv1_1": rng.standard_normal(64).astype(np.float32),
         "global_step": np.asarray(1234, np.int64), "flags/b": np.asarray([True, False]),
         "h": rng.standard_normal((2, 5)).astype(np.float16), "empty": np.zeros((0, 4), np.float32)}
    for i in range(600):   # enough keys for several 4 KB data blocks with shared prefixes
        t["regressor/conv2d_%d/batch_normalization/moving_variance" % i] = rng.standard_normal(3).astype(np.float32)
    prefix = str(tmp_path / "run" / "model-77")
    ck.write_checkpoint(prefix, t)
    assert os.path.exists(prefix + ".index") and os.path.exists(prefix + ".data-00000-of-00001")
    r = ck.CheckpointReader(prefix)
    assert set(r.get_variable_to_shape_map()) == set(t)
    assert r.get_variable_to_shape_map()["feature_extractor/kernel_conv1_1"] == [3, 3, 3, 64]
    for k, v in t.items():
        got = r.get_tensor(k, verify=True)
        assert got.dtype == v.dtype and got.shape == v.shape
        np.testing.assert_array_equal(got, v)
    assert ck.latest_checkpoint(str(tmp_path / "run")) == prefix
    with pytest.raises(ck.CheckpointError, match="not found"):
        r.get_tensor("nope")


def test_corruption_is_detected(ck, tmp_path):
    prefix = str(tmp_path / "m")
    ck.write_checkpoint(prefix, {"a": np.arange(10, dtype=np.float32), "b": np.ones(3, np.float32)})
    raw = bytearray(open(prefix + ".index", "rb").read())
    raw[10] ^= 0x40
    open(prefix + ".index", "wb").write(bytes(raw))
    with pytest.raises(ck.CheckpointError, match="checksum"):
        ck.CheckpointReader(prefix)
    ck.write_checkpoint(prefix, {"a": np.arange(10, dtype=np.float32)})
    d = bytearray(open(prefix + ".data-00000-of-00001", "rb").read())
    d[3] ^= 1
    open(prefix + ".data-00000-of-00001", "wb").write(bytes(d))
    with pytest.raises(ck.CheckpointError, match="payload checksum"):
        ck.CheckpointReader(prefix).get_tensor("a", verify=True)


def test_hand_assembled_table_with_restarts_and_snappy(ck, tmp_path):
    """A table built byte by byte here (not by the module's writer): prefix compression
    across a restart point, and a snappy-compressed data block."""
    def varint(v):
        out = bytearray()
        while v >= 0x80:
            out.append((v & 0x7F) | 0x80)
            v >>= 7
        out.append(v)
        return bytes(out)

    def entry(shared, suffix, value):
        return varint(shared) + varint(len(suffix)) + varint(len(value)) + suffix + value

    blk = entry(0, b"apple", b"1") + entry(3, b"ly", b"22")       # "apple", "apply"
    r1 = len(blk)
    blk += entry(0, b"banana", b"333")                             # restart point
    blk += struct.pack("<III", 0, r1, 2)
    # snappy: uncompressed length, one literal of the whole block
    comp = varint(len(blk)) + bytes([(len(blk) - 1) << 2]) + blk if len(blk) <= 60 else None
    assert comp is not None
    f = bytearray()

    def emit(body, ctype):
        off = len(f)
        f.extend(body)
        f.append(ctype)
        f.extend(struct.pack("<I", ck.mask_crc(ck.crc32c(bytes(body) + bytes([ctype])))))
        return varint(off) + varint(len(body))

    h_data = emit(comp, 1)
    h_meta = emit(struct.pack("<II", 0, 1), 0)
    idx = entry(0, b"banana", h_data) + struct.pack("<II", 0, 1)
    h_idx = emit(idx, 0)
    footer = h_meta + h_idx
    f.extend(footer + b"\x00" * (40 - len(footer)) + struct.pack("<Q", ck.TABLE_MAGIC))
    p = tmp_path / "t.index"
    p.write_bytes(bytes(f))
    assert ck._read_table(str(p)) == [(b"apple", b"1"), (b"apply", b"22"), (b"banana", b"333")]
    # snappy copies (overlapping run): "ab" + copy(offset 2, len 6) -> "abababab"
    s = varint(8) + bytes([(2 - 1) << 2]) + b"ab" + bytes([((6 - 4) << 2) | 1, 2])
    assert ck._snappy_decompress(s) == b"abababab"


def test_v1_checkpoint_is_rejected_with_a_clear_message(ck, tmp_path):
    p = tmp_path / "vgg_16.ckpt"
    ck._write_table(str(p), [(b"", b"x")])           # any SSTable without a .index sibling looks like V1
    with pytest.raises(ck.CheckpointError, match="V1 checkpoint"):
        ck.CheckpointReader(str(p))
    with pytest.raises(ck.CheckpointError, match="no checkpoint"):
        ck.CheckpointReader(str(tmp_path / "missing"))


def test_model_save_load_and_vgg_pretraining_names(ck, tmp_path):
    """save_weight writes `<path>-<global_step>` as a V2 bundle with the reference's variable
    names (SSD300.py:490-504); load_weight restores it; a vgg_16-named bundle feeds the
    `pretraining_weight` constructor argument (SSD300.py:31,195-301)."""
    import SSD300
    from odt_b200.api import VGG16_CKPT_NAMES
    cfg = model_cfg("ssd300")
    m = SSD300.SSD300(cfg, None)
    w = m.get_weights()
    out = m.save_weight("latest", str(tmp_path / "ssd" / "model"))
    assert out.endswith("model-0") and ck.is_v2_checkpoint(out)
    r = ck.CheckpointReader(out)
    assert r.has_tensor("global_step") and r.has_tensor("feature_extractor/kenrel_conv2_1")
    # the reference's variable is int32 (SSD300.py:43); Saver.restore checks the dtype
    assert r.get_tensor("global_step").dtype == np.int32
    m2 = SSD300.SSD300(dict(cfg), None)
    m2.seed = 99                                      # different random init, then restore
    m2.load_weight(out)
    for k, v in w.items():
        np.testing.assert_array_equal(m2.get_weights()[k], v)
    # VGG-16 classification checkpoint names -> the 13 backbone convs
    rng = np.random.default_rng(5)
    vgg = {}
    for ck_name, (kvar, bvar) in VGG16_CKPT_NAMES.items():
        vgg["vgg_16/%s/weights" % ck_name] = rng.standard_normal(w[kvar].shape).astype(np.float32)
        vgg["vgg_16/%s/biases" % ck_name] = rng.standard_normal(w[bvar].shape).astype(np.float32)
    vgg["vgg_16/fc8/weights"] = np.zeros((1, 1, 8, 4), np.float32)   # ignored extras
    ck.write_checkpoint(str(tmp_path / "vgg_16.ckpt"), vgg)
    m3 = SSD300.SSD300(dict(cfg, pretraining_weight=str(tmp_path / "vgg_16.ckpt")), None)
    w3 = m3.get_weights()
    np.testing.assert_array_equal(w3["feature_extractor/kernel_conv1_1"], vgg["vgg_16/conv1/conv1_1/weights"])
    np.testing.assert_array_equal(w3["feature_extractor/bias_conv_3_1"], vgg["vgg_16/conv3/conv3_1/biases"])
    np.testing.assert_array_equal(w3["feature_extractor/kernel_conv5_3"], vgg["vgg_16/conv5/conv5_3/weights"])


# ---------------------------------------------------------------- V1 files ----
def _pb(field, wt, payload):
    def varint(v):
        out = bytearray()
        v &= (1 << 64) - 1
        while v >= 0x80:
            out.append((v & 0x7F) | 0x80)
            v >>= 7
        out.append(v)
        return bytes(out)
    if wt == 0:
        return varint((field << 3) | 0) + varint(payload)
    return varint((field << 3) | 2) + varint(len(payload)) + payload


def _shape_pb(shape):
    return b"".join(_pb(2, 2, _pb(1, 0, d)) for d in shape)


def _v1_file(ck, path, tensors, split=None):
    """Hand-assembled V1 checkpoint: {name: array}; `split` = name whose first axis is saved as two slices."""
    meta = b""
    entries = []
    for name in sorted(tensors):
        a = tensors[name]
        dt = {np.dtype(np.float32): 1, np.dtype(np.int64): 9}[a.dtype]
        full_slice = _pb(1, 2, b"") * a.ndim        # extent without start/length = whole dimension
        meta += _pb(1, 2, _pb(1, 2, name.encode()) + _pb(2, 2, _shape_pb(a.shape)) + _pb(3, 0, dt) + _pb(4, 2, full_slice))
        parts = [(0, a.shape[0])] if (a.ndim and name != split) else ([(0, 1), (1, a.shape[0] - 1)] if a.ndim else [None])
        for i, pr in enumerate(parts):
            if pr is None:
                sub, ext = a, b""
            else:
                sub = a[pr[0]:pr[0] + pr[1]]
                first = _pb(1, 2, (_pb(1, 0, pr[0]) if pr[0] else b"") + _pb(2, 0, pr[1])) if name == split else _pb(1, 2, b"")
                ext = first + _pb(1, 2, b"") * (a.ndim - 1)
            if a.dtype == np.float32 and name.endswith("weights"):
                body = _pb(4, 2, sub.astype("<f4").tobytes())               # tensor_content
            elif a.dtype == np.float32:
                body = _pb(5, 2, sub.astype("<f4").tobytes())               # packed float_val
            else:
                body = _pb(10, 2, b"".join(_pb(1, 0, int(x))[1:] for x in sub.reshape(-1)))  # packed int64_val varints
            tp = _pb(1, 0, dt) + _pb(2, 2, _shape_pb(sub.shape)) + body
            saved = _pb(2, 2, _pb(1, 2, name.encode()) + _pb(2, 2, ext) + _pb(3, 2, tp))
            entries.append((b"\x00" + name.encode() + b"\x00\x01" + bytes([i]), saved))
    items = [(b"", _pb(1, 2, meta))] + sorted(entries)
    ck._write_table(path, items)


def test_v1_single_file_checkpoint(ck, tmp_path):
    rng = np.random.default_rng(3)
    t = {"vgg_16/conv1/conv1_1/weights": rng.standard_normal((3, 3, 3, 64)).astype(np.float32),
         "vgg_16/conv1/conv1_1/biases": rng.standard_normal(64).astype(np.float32),
         "vgg_16/fc8/biases": rng.standard_normal(10).astype(np.float32),
         "global_step": np.asarray(30000, np.int64)}
    p = str(tmp_path / "vgg_16.ckpt")
    _v1_file(ck, p, t, split="vgg_16/fc8/biases")
    assert ck.is_checkpoint(p) and not ck.is_v2_checkpoint(p)
    r = ck.open_checkpoint(p)
    assert isinstance(r, ck.CheckpointReaderV1)
    assert r.get_variable_to_shape_map()["vgg_16/conv1/conv1_1/weights"] == [3, 3, 3, 64]
    for k, v in t.items():
        got = r.get_tensor(k)
        assert got.dtype == v.dtype and got.shape == v.shape
        np.testing.assert_array_equal(got, v)
    assert set(ck.read_checkpoint(p)) == set(t)


def test_v1_vgg_checkpoint_feeds_the_ssd_constructor(ck, tmp_path):
    """The slim `vgg_16.ckpt` the reference's SSD constructors read is a V1 file (SSD300.py:31,195-301)."""
    import SSD300
    from odt_b200.api import VGG16_CKPT_NAMES
    cfg = model_cfg("ssd300")
    w = SSD300.SSD300(cfg, None).get_weights()
    rng = np.random.default_rng(9)
    vgg = {}
    for ck_name, (kvar, bvar) in VGG16_CKPT_NAMES.items():
        vgg["vgg_16/%s/weights" % ck_name] = rng.standard_normal(w[kvar].shape).astype(np.float32)
        vgg["vgg_16/%s/biases" % ck_name] = rng.standard_normal(w[bvar].shape).astype(np.float32)
    p = str(tmp_path / "vgg_16.ckpt")
    _v1_file(ck, p, vgg)
    m = SSD300.SSD300(dict(cfg, pretraining_weight=p), None)
    w2 = m.get_weights()
    np.testing.assert_array_equal(w2["feature_extractor/kernel_conv4_3"], vgg["vgg_16/conv4/conv4_3/weights"])
    np.testing.assert_array_equal(w2["feature_extractor/bias_conv1_2"], vgg["vgg_16/conv1/conv1_2/biases"])


# ------------------------------------------------------------ property tests ----
def _snappy_compress_simple(data):
    """Tiny greedy snappy encoder (test-only): literals + 2-byte-offset copies of 4..64 bytes."""
    def varint(v):
        out = bytearray()
        while v >= 0x80:
            out.append((v & 0x7F) | 0x80)
            v >>= 7
        out.append(v)
        return bytes(out)

    out = bytearray(varint(len(data)))
    lit = bytearray()

    def flush():
        nonlocal lit
        i = 0
        while i < len(lit):
            chunk = lit[i:i + 60]
            out.append((len(chunk) - 1) << 2)
            out.extend(chunk)
            i += 60
        lit = bytearray()

    table, i, n = {}, 0, len(data)
    while i < n:
        key = bytes(data[i:i + 4])
        j = table.get(key)
        if len(key) == 4 and j is not None and 0 < i - j < 65536:
            ln = 4
            while ln < 64 and i + ln < n and data[j + ln] == data[i + ln]:
                ln += 1
            flush()
            out.append(((ln - 1) << 2) | 2)
            out.extend(struct.pack("<H", i - j))
            for k in range(i, i + ln):
                table[bytes(data[k:k + 4])] = k
            i += ln
        else:
            table[key] = i
            lit.append(data[i])
            i += 1
    flush()
    return bytes(out)


def test_snappy_decoder_against_a_test_encoder(ck):
    hyp = pytest.importorskip("hypothesis")
    from hypothesis import given, settings, strategies as st

    @settings(max_examples=60, deadline=None)
    @given(st.binary(max_size=600), st.integers(1, 6))
    def run(seed_bytes, rep):
        data = (seed_bytes * rep)[:3000]            # repetition creates back-references, incl. overlapping runs
        assert ck._snappy_decompress(_snappy_compress_simple(data)) == data
    run()
    assert ck._snappy_decompress(_snappy_compress_simple(b"a" * 500)) == b"a" * 500


def test_bundle_round_trip_property(ck, tmp_path):
    hyp = pytest.importorskip("hypothesis")
    from hypothesis import given, settings, strategies as st
    names = st.text(alphabet="abcdefgh/_0123456789", min_size=1, max_size=24)
    shapes = st.lists(st.integers(0, 5), min_size=0, max_size=3)

    @settings(max_examples=25, deadline=None)
    @given(st.dictionaries(names, shapes, min_size=1, max_size=12), st.integers(0, 2 ** 31 - 1))
    def run(spec, seed):
        rng = np.random.default_rng(seed)
        t = {k: rng.standard_normal(s).astype(np.float32) for k, s in spec.items()}
        prefix = str(tmp_path / ("p%d" % seed))
        ck.write_checkpoint(prefix, t)
        got = ck.read_checkpoint(prefix, verify=True)
        assert set(got) == set(t)
        for k in t:
            np.testing.assert_array_equal(got[k], t[k])
    run()


def test_corrupt_files_raise_checkpoint_error_only(ck, tmp_path):
    """Byte flips / truncations of the index (checks off), of a V1 file and of snappy blocks must surface as
    CheckpointError, never as a raw IndexError / struct.error / UnicodeDecodeError."""
    rng = np.random.default_rng(11)
    t = {"a/kernel": rng.standard_normal((3, 3, 3, 8)).astype(np.float32), "global_step": np.asarray(7, np.int64)}
    pre = str(tmp_path / "m-7")
    ck.write_checkpoint(pre, t)
    good = open(pre + ".index", "rb").read()
    p1 = str(tmp_path / "v1.ckpt")
    _v1_file(ck, p1, {"x/w": rng.standard_normal((4, 6)).astype(np.float32)}, split="x/w")
    good1 = open(p1, "rb").read()

    def mutate(raw):
        b = bytearray(raw)
        for _ in range(int(rng.integers(1, 4))):
            b[int(rng.integers(len(b)))] = int(rng.integers(256))
        return bytes(b[:int(rng.integers(len(b)))] if rng.random() < 0.2 else b)

    for _ in range(300):
        open(pre + ".index", "wb").write(mutate(good))
        open(p1, "wb").write(mutate(good1))
        for fn in (lambda: [r.get_tensor(n, verify=True) for r in [ck.CheckpointReader(pre, verify_index=False)]
                            for n in r.get_variable_to_shape_map()],
                   lambda: [r.get_tensor(n) for r in [ck.CheckpointReaderV1(p1, verify=False)]
                            for n in r.get_variable_to_shape_map()],
                   lambda: ck._snappy_decompress(bytes(rng.integers(0, 256, int(rng.integers(1, 40))).astype(np.uint8)))):
            try:
                fn()
            except ck.CheckpointError:
                pass


def test_saver_semantics_of_load_weight_and_pretraining_loader(ck, tmp_path):
    """`load_weight` = Saver() over ALL variables: a file lacking one is an error (NotFoundError in TF).
    `load_pretraining_weight` = Saver(trainable backbone variables): restores kernels / biases / gamma / beta of
    the backbone only -- not the heads, not the BN moving statistics (YOLOv3.py:376-378,481-483)."""
    import YOLOv3
    cfg = dict(mode="test", data_format="channels_last", num_classes=20, weight_decay=1e-4, keep_prob=0.5,
               batch_size=1, nms_score_threshold=0.5, nms_max_boxes=10, nms_iou_threshold=0.45,
               data_shape=[416, 416, 3], coord_scale=1, noobj_scale=1, obj_scale=5., class_scale=1., num_priors=3,
               priors=[[[10, 13], [16, 30], [33, 23]], [[30, 61], [62, 45], [59, 119]], [[116, 90], [156, 198], [373, 326]]])
    m = YOLOv3.YOLOv3(cfg, None)
    w0 = {k: v.copy() for k, v in m.get_weights().items()}
    names = m.pretraining_variables()
    assert names and all(n.startswith("backone/") for n in names)
    assert not any(n.endswith(("moving_mean", "moving_variance")) for n in names)
    assert "backone/batch_normalization/gamma" in names and "backone/conv2d/kernel" in names
    rng = np.random.default_rng(3)
    donor = {k: rng.standard_normal(v.shape).astype(np.float32) for k, v in w0.items()}
    full = str(tmp_path / "full")
    ck.write_checkpoint(full, donor)
    m.load_pretraining_weight(full)
    w1 = m.get_weights()
    for k in w0:
        want = donor[k] if k in names else w0[k]                    # heads and moving statistics untouched
        np.testing.assert_array_equal(w1[k], want, err_msg=k)
    # a backbone-only file is fine for the pretraining loader, an error for load_weight
    part = str(tmp_path / "backbone")
    ck.write_checkpoint(part, {k: donor[k] for k in names})
    m.load_pretraining_weight(part)
    with pytest.raises(ck.CheckpointError, match="lacks"):
        m.load_weight(part)
    missing_one = dict(donor)
    missing_one.pop(names[5])
    p2 = str(tmp_path / "holed")
    ck.write_checkpoint(p2, missing_one)
    with pytest.raises(ck.CheckpointError, match=names[5].replace("/", "/")):
        m.load_pretraining_weight(p2)
    m.load_weight(full)                                             # everything present: restores heads too
    np.testing.assert_array_equal(m.get_weights()["head/conv2d/kernel"] if "head/conv2d/kernel" in donor
                                  else m.get_weights()[sorted(donor)[0]],
                                  donor["head/conv2d/kernel"] if "head/conv2d/kernel" in donor else donor[sorted(donor)[0]])
