"""Reader / writer for TensorFlow V2 checkpoints (TensorBundle: `<prefix>.index` +
`<prefix>.data-SSSSS-of-NNNNN`), the format `tf.train.Saver().save / .restore` and
`NewCheckpointReader` use in the reference (SSD300.py:31,195-301,490-504;
RetinaNet.py:505-557; YOLOv3.py:376-385; FCOS.py:384-436).  No TensorFlow needed.

Format restated from the public definitions (nothing copied):
* `.index` is an SSTable in the LevelDB table format: data blocks of prefix-compressed
  (shared, non_shared, value_len varint32 + key suffix + value) entries with a restart
  array, each followed by a 5-byte trailer (compression type, masked CRC-32C of block +
  type); a metaindex block, an index block (last key of a data block -> BlockHandle) and a
  48-byte footer (two BlockHandles, padding, magic 0xdb4775248b80fb57).
* key "" holds a BundleHeaderProto {1: num_shards, 2: endianness, 3: version{1: producer}};
  every other key is a tensor name with a BundleEntryProto {1: dtype, 2: shape{2: dim{1:
  size}}, 3: shard_id, 4: offset, 5: size, 6: fixed32 masked crc32c of the bytes}.
* tensor bytes are raw little-endian, row-major, at [offset, offset+size) of the shard.

PARITY UNPINNED: no TensorFlow and no TF-written checkpoint exists in this environment; the
reader is exercised against this module's own writer plus hand-assembled blocks
(tests/test_checkpoint.py).

V1 checkpoints (ONE SSTable file, usually snappy-compressed blocks, e.g. the 2016 slim `vgg_16.ckpt`
the SSD constructors read) are supported read-only: key "" holds SavedTensorSlices{1: meta{1: tensor
SavedSliceMeta{1: name, 2: shape, 3: type, 4: slices}}}, every other entry SavedTensorSlices{2: data
SavedSlice{1: name, 2: slice TensorSliceProto{1: extent{1: start, 2: length}}, 3: data TensorProto{1:
dtype, 4: tensor_content | 5: float_val | 6: double_val | 7: int_val | 10: int64_val | 11: bool_val |
13: half_val}}}.  Slices are pasted into the full tensor by their extents.
"""
import os
import struct

import numpy as np

TABLE_MAGIC = 0xDB4775248B80FB57
_MASK_DELTA = 0xA282EAD8

# tensorflow DataType enum <-> numpy
_DTYPES = {1: np.float32, 2: np.float64, 3: np.int32, 4: np.uint8, 5: np.int16, 6: np.int8, 9: np.int64,
           10: np.bool_, 17: np.uint16, 19: np.float16, 22: np.uint32, 23: np.uint64}
_DTYPE_OF = {np.dtype(v): k for k, v in _DTYPES.items()}


class CheckpointError(RuntimeError):
    pass


class _malformed:
    """Context manager: whatever a parser trips over in corrupt input (a scalar where bytes were expected, a
    bad UTF-8 name, an extent outside its tensor, a missing shard ...) surfaces as CheckpointError."""

    def __init__(self, what):
        self.what = what

    def __enter__(self):
        return self

    def __exit__(self, et, ev, tb):
        if et is not None and not issubclass(et, CheckpointError) and issubclass(
                et, (ValueError, TypeError, AttributeError, IndexError, KeyError, OverflowError, struct.error, OSError,
                     MemoryError)):
            raise CheckpointError("%s: malformed or unreadable (%s: %s)" % (self.what, et.__name__, ev)) from None
        return False


# ------------------------------------------------------------------ crc32c ---
_crc_table = None
_crc_native = None


def crc32c(data, crc=0):
    """CRC-32C (Castagnoli).  Uses the C-ABI helper of libodt_b200.so when it is built,
    a table-driven Python loop otherwise (index blocks are small)."""
    global _crc_table, _crc_native
    if _crc_native is None:
        try:
            from . import lib as L
            fn = L.load().odt_crc32c
            _crc_native = fn
        except Exception:  # library not built: host-only fallback
            _crc_native = False
    mv = data if isinstance(data, bytes) else bytes(data)
    if _crc_native:
        return int(_crc_native(crc, mv, len(mv)))
    if _crc_table is None:
        t = []
        for i in range(256):
            c = i
            for _ in range(8):
                c = (c >> 1) ^ 0x82F63B78 if c & 1 else c >> 1
            t.append(c)
        _crc_table = t
    c = crc ^ 0xFFFFFFFF
    for b in bytes(mv):
        c = _crc_table[(c ^ b) & 0xFF] ^ (c >> 8)
    return c ^ 0xFFFFFFFF


def mask_crc(c):
    return (((c >> 15) | (c << 17)) + _MASK_DELTA) & 0xFFFFFFFF


# ----------------------------------------------------------------- varints ---
def _put_varint(out, v):
    v &= (1 << 64) - 1
    while v >= 0x80:
        out.append((v & 0x7F) | 0x80)
        v >>= 7
    out.append(v)


def _get_varint(buf, pos):
    shift, v = 0, 0
    while True:
        if pos >= len(buf):
            raise CheckpointError("truncated varint")
        b = buf[pos]
        pos += 1
        v |= (b & 0x7F) << shift
        if b < 0x80:
            return v, pos
        shift += 7
        if shift > 63:
            raise CheckpointError("varint too long")


# ---------------------------------------------------------------- protobuf ---
def _pb_fields(buf):
    """Yield (field number, wire type, value) of one message; value = int or bytes.  Malformed input
    (truncated field, length past the end, group wire types) raises CheckpointError."""
    if not isinstance(buf, (bytes, bytearray, memoryview)):
        raise CheckpointError("protobuf: a length-delimited field was expected, got a scalar")
    pos, n = 0, len(buf)
    while pos < n:
        tag, pos = _get_varint(buf, pos)
        f, wt = tag >> 3, tag & 7
        if wt == 0:
            v, pos = _get_varint(buf, pos)
        elif wt == 1 or wt == 5:
            size = 8 if wt == 1 else 4
            if pos + size > n:
                raise CheckpointError("protobuf: truncated fixed%d field" % (8 * size))
            v = struct.unpack_from("<Q" if wt == 1 else "<I", buf, pos)[0]
            pos += size
        elif wt == 2:
            ln, pos = _get_varint(buf, pos)
            if pos + ln > n:
                raise CheckpointError("protobuf: field of %d bytes runs past the end of its message" % ln)
            v = bytes(buf[pos:pos + ln])
            pos += ln
        else:
            raise CheckpointError("unsupported protobuf wire type %d" % wt)
        yield f, wt, v


def _pb_varint(out, field, v):
    _put_varint(out, (field << 3) | 0)
    _put_varint(out, v)


def _pb_bytes(out, field, b):
    _put_varint(out, (field << 3) | 2)
    _put_varint(out, len(b))
    out.extend(b)


def _signed64(v):
    return v - (1 << 64) if v >= (1 << 63) else v


def _parse_shape(buf):
    dims = []
    for f, _, v in _pb_fields(buf):
        if f == 2:  # Dim
            size = 0
            for f2, _, v2 in _pb_fields(v):
                if f2 == 1:
                    size = _signed64(v2)
            dims.append(size)
    return tuple(dims)


def _parse_entry(buf):
    e = {"dtype": 0, "shape": (), "shard_id": 0, "offset": 0, "size": 0, "crc32c": None, "sliced": False}
    wire = {1: 0, 2: 2, 3: 0, 4: 0, 5: 0, 6: 5, 7: 2}
    for f, wt, v in _pb_fields(buf):
        if f in wire and wt != wire[f]:
            raise CheckpointError("BundleEntryProto field %d has wire type %d" % (f, wt))
        if f == 1:
            e["dtype"] = v
        elif f == 2:
            e["shape"] = _parse_shape(v)
        elif f == 3:
            e["shard_id"] = v
        elif f == 4:
            e["offset"] = v
        elif f == 5:
            e["size"] = v
        elif f == 6:
            e["crc32c"] = v
        elif f == 7:
            e["sliced"] = True
    return e


def _encode_entry(dtype, shape, shard_id, offset, size, crc):
    out = bytearray()
    _pb_varint(out, 1, dtype)
    sh = bytearray()
    for d in shape:
        dim = bytearray()
        _pb_varint(dim, 1, int(d))
        _pb_bytes(sh, 2, dim)
    _pb_bytes(out, 2, sh)
    if shard_id:
        _pb_varint(out, 3, shard_id)
    if offset:
        _pb_varint(out, 4, offset)
    _pb_varint(out, 5, size)
    _put_varint(out, (6 << 3) | 5)
    out.extend(struct.pack("<I", crc))
    return bytes(out)


def _encode_header(num_shards):
    out = bytearray()
    _pb_varint(out, 1, num_shards)
    # endianness LITTLE = 0 is the default and omitted; version {producer: 1}
    ver = bytearray()
    _pb_varint(ver, 1, 1)
    _pb_bytes(out, 3, ver)
    return bytes(out)


# ------------------------------------------------------------------ snappy ---
def _snappy_decompress(buf):
    """Raw snappy block format (index blocks written with compression enabled)."""
    with _malformed("snappy block"):
        return _snappy_body(buf)


def _snappy_body(buf):
    n, pos = _get_varint(buf, 0)
    out = bytearray()
    while pos < len(buf):
        tag = buf[pos]
        pos += 1
        kind = tag & 3
        if kind == 0:
            ln = tag >> 2
            if ln >= 60:
                nb = ln - 59
                ln = int.from_bytes(buf[pos:pos + nb], "little")
                pos += nb
            ln += 1
            if pos + ln > len(buf) or len(out) + ln > n:
                raise CheckpointError("corrupt snappy block")
            out.extend(buf[pos:pos + ln])
            pos += ln
            continue
        if kind == 1:
            ln = ((tag >> 2) & 7) + 4
            off = ((tag >> 5) << 8) | buf[pos]
            pos += 1
        elif kind == 2:
            ln = (tag >> 2) + 1
            off = buf[pos] | (buf[pos + 1] << 8)
            pos += 2
        else:
            ln = (tag >> 2) + 1
            off = int.from_bytes(buf[pos:pos + 4], "little")
            pos += 4
        if off == 0 or off > len(out) or len(out) + ln > n:
            raise CheckpointError("corrupt snappy block")
        for _ in range(ln):  # may overlap
            out.append(out[-off])
    if len(out) != n:
        raise CheckpointError("snappy length mismatch")
    return bytes(out)


# ------------------------------------------------------------------- table ---
def _read_block(f, offset, size, verify=True):
    f.seek(offset)
    raw = f.read(size + 5)
    if len(raw) != size + 5:
        raise CheckpointError("truncated table block")
    body, ctype = raw[:size], raw[size]
    stored = struct.unpack_from("<I", raw, size + 1)[0]
    if verify and mask_crc(crc32c(raw[:size + 1])) != stored:
        raise CheckpointError("table block checksum mismatch at offset %d" % offset)
    if ctype == 0:
        return body
    if ctype == 1:
        return _snappy_decompress(body)
    raise CheckpointError("unknown block compression %d" % ctype)


def _block_entries(block):
    if len(block) < 4:
        raise CheckpointError("bad block")
    num_restarts = struct.unpack_from("<I", block, len(block) - 4)[0]
    limit = len(block) - 4 - 4 * num_restarts
    if limit < 0:
        raise CheckpointError("bad restart array")
    pos, key = 0, b""
    while pos < limit:
        shared, pos = _get_varint(block, pos)
        non_shared, pos = _get_varint(block, pos)
        vlen, pos = _get_varint(block, pos)
        if shared > len(key) or pos + non_shared + vlen > limit:
            raise CheckpointError("bad table block entry")
        key = key[:shared] + bytes(block[pos:pos + non_shared])
        pos += non_shared
        yield key, bytes(block[pos:pos + vlen])
        pos += vlen


def _read_table(path, verify=True):
    """All (key, value) pairs of an SSTable file, in key order."""
    with open(path, "rb") as f:
        f.seek(0, os.SEEK_END)
        size = f.tell()
        if size < 48:
            raise CheckpointError("%s: too short for a table" % path)
        f.seek(size - 48)
        footer = f.read(48)
        if struct.unpack_from("<Q", footer, 40)[0] != TABLE_MAGIC:
            raise CheckpointError("%s: not an SSTable (bad magic)" % path)
        pos = 0
        _, pos = _get_varint(footer, pos)       # metaindex handle
        _, pos = _get_varint(footer, pos)
        ioff, pos = _get_varint(footer, pos)    # index handle
        isz, pos = _get_varint(footer, pos)
        out = []
        for _, handle in _block_entries(_read_block(f, ioff, isz, verify)):
            boff, p2 = _get_varint(handle, 0)
            bsz, _ = _get_varint(handle, p2)
            out.extend(_block_entries(_read_block(f, boff, bsz, verify)))
    return out


class _BlockBuilder:
    def __init__(self, restart_interval=16):
        self.buf, self.restarts, self.count, self.last = bytearray(), [0], 0, b""
        self.interval = restart_interval

    def add(self, key, value):
        shared = 0
        if self.count % self.interval == 0 and self.count:
            self.restarts.append(len(self.buf))
        elif self.count:
            m = min(len(key), len(self.last))
            while shared < m and key[shared] == self.last[shared]:
                shared += 1
        _put_varint(self.buf, shared)
        _put_varint(self.buf, len(key) - shared)
        _put_varint(self.buf, len(value))
        self.buf.extend(key[shared:])
        self.buf.extend(value)
        self.last = key
        self.count += 1

    def finish(self):
        out = bytearray(self.buf)
        for r in self.restarts:
            out.extend(struct.pack("<I", r))
        out.extend(struct.pack("<I", len(self.restarts)))
        return bytes(out)


def _write_table(path, items, block_size=4096):
    """items: list of (key bytes, value bytes) sorted by key."""
    with open(path, "wb") as f:
        index = _BlockBuilder(restart_interval=1)

        def emit(block):
            off = f.tell()
            f.write(block)
            f.write(b"\x00")
            f.write(struct.pack("<I", mask_crc(crc32c(block + b"\x00"))))
            return off, len(block)

        def handle(off, size):
            h = bytearray()
            _put_varint(h, off)
            _put_varint(h, size)
            return bytes(h)

        bb = _BlockBuilder()
        for key, value in items:
            bb.add(key, value)
            if len(bb.buf) >= block_size:
                off, size = emit(bb.finish())
                index.add(bb.last, handle(off, size))
                bb = _BlockBuilder()
        if bb.count:
            off, size = emit(bb.finish())
            index.add(bb.last, handle(off, size))
        moff, msize = emit(_BlockBuilder().finish())   # empty metaindex
        ioff, isize = emit(index.finish())
        footer = bytearray(handle(moff, msize) + handle(ioff, isize))
        footer.extend(b"\x00" * (40 - len(footer)))
        footer.extend(struct.pack("<Q", TABLE_MAGIC))
        f.write(footer)


# ------------------------------------------------------------------ public ---
def is_v2_checkpoint(prefix):
    return os.path.exists(prefix + ".index")


def _shard_path(prefix, shard, num):
    return "%s.data-%05d-of-%05d" % (prefix, shard, num)


class CheckpointReader:
    """`tf.train.NewCheckpointReader(prefix)` for V2 bundles: has_tensor / get_tensor /
    get_variable_to_shape_map (SSD300.py:31,195-301 call sites)."""

    def __init__(self, prefix, verify_index=True):
        self.prefix = prefix
        if not os.path.exists(prefix + ".index"):
            if os.path.isfile(prefix):
                with open(prefix, "rb") as f:
                    f.seek(0, os.SEEK_END)
                    n = f.tell()
                    magic = 0
                    if n >= 48:
                        f.seek(n - 8)
                        magic = struct.unpack("<Q", f.read(8))[0]
                if magic == TABLE_MAGIC:
                    raise CheckpointError("%s is a V1 checkpoint (one SavedTensorSlices table): open it with "
                                          "open_checkpoint() / CheckpointReaderV1" % prefix)
            raise CheckpointError("no checkpoint at %r (expected %s.index)" % (prefix, prefix))
        self.entries, self.num_shards = {}, 1
        with _malformed(prefix + ".index"):
            for key, value in _read_table(prefix + ".index", verify_index):
                if key == b"":
                    for f, _, v in _pb_fields(value):
                        if f == 1:
                            self.num_shards = int(v)
                        elif f == 2 and v != 0:
                            raise CheckpointError("big-endian bundles are not supported")
                else:
                    self.entries[key.decode("utf-8")] = _parse_entry(value)

    def has_tensor(self, name):
        return name in self.entries

    def get_variable_to_shape_map(self):
        return {k: list(e["shape"]) for k, e in self.entries.items()}

    def get_tensor(self, name, verify=False):
        e = self.entries.get(name)
        if e is None:
            raise CheckpointError("tensor %r not found in %s" % (name, self.prefix))
        if e["sliced"]:
            raise CheckpointError("partitioned variable %r is not supported" % name)
        if e["dtype"] not in _DTYPES:
            raise CheckpointError("dtype %d of %r is not supported" % (e["dtype"], name))
        dt = np.dtype(_DTYPES[e["dtype"]])
        with _malformed("tensor %r of %s" % (name, self.prefix)):
            count = 1
            for d in e["shape"]:
                if d < 0:
                    raise CheckpointError("negative dimension in the shape of %r" % name)
                count *= int(d)
            if count * dt.itemsize != e["size"]:
                raise CheckpointError("size of %r does not match its shape" % name)
            shard = _shard_path(self.prefix, e["shard_id"], self.num_shards)
            if e["offset"] + e["size"] > os.path.getsize(shard):
                raise CheckpointError("truncated data shard for %r" % name)
            with open(shard, "rb") as f:
                f.seek(e["offset"])
                raw = f.read(e["size"])
            if len(raw) != e["size"]:
                raise CheckpointError("truncated data shard for %r" % name)
            if verify and e["crc32c"] is not None and mask_crc(crc32c(raw)) != e["crc32c"]:
                raise CheckpointError("payload checksum mismatch for %r" % name)
            return np.frombuffer(raw, dtype=dt.newbyteorder("<")).astype(dt).reshape(e["shape"])


def _parse_extents(buf):
    """TensorSliceProto -> [(start, length or None)] per dimension."""
    ext = []
    for f, _, v in _pb_fields(buf):
        if f == 1:
            start, length = 0, None
            for f2, _, v2 in _pb_fields(v):
                if f2 == 1:
                    start = _signed64(v2)
                elif f2 == 2:
                    length = _signed64(v2)
            ext.append((start, length))
    return ext


def _parse_tensor_proto(buf):
    """TensorProto -> (dtype enum, shape, flat numpy array)."""
    dtype, shape, content = 0, (), None
    packed = {5: ("<f4", []), 6: ("<f8", []), 7: ("varint", []), 10: ("varint", []), 11: ("varint", []), 13: ("varint", [])}
    for f, wt, v in _pb_fields(buf):
        if f == 1:
            dtype = v
        elif f == 2:
            shape = _parse_shape(v)
        elif f == 4:
            content = v
        elif f in packed:
            kind, acc = packed[f]
            if wt == 2 and kind != "varint":
                acc.append(np.frombuffer(v, kind))
            elif wt == 2:
                pos, vals = 0, []
                while pos < len(v):
                    x, pos = _get_varint(v, pos)
                    vals.append(_signed64(x))
                acc.append(np.asarray(vals, np.int64))
            elif wt == 5:
                acc.append(np.frombuffer(struct.pack("<I", v), "<f4"))
            elif wt == 1:
                acc.append(np.frombuffer(struct.pack("<Q", v), "<f8"))
            else:
                acc.append(np.asarray([_signed64(v)], np.int64))
    if dtype not in _DTYPES:
        raise CheckpointError("dtype %d is not supported" % dtype)
    dt = np.dtype(_DTYPES[dtype])
    if content is not None:
        flat = np.frombuffer(content, dt.newbyteorder("<")).astype(dt)
    else:
        field = {np.dtype(np.float32): 5, np.dtype(np.float64): 6, np.dtype(np.int64): 10, np.dtype(np.bool_): 11,
                 np.dtype(np.float16): 13}.get(dt, 7)
        parts = packed[field][1]
        flat = np.concatenate(parts) if parts else np.zeros(0)
        if field == 13:  # half_val carries the raw 16 bits in an int
            flat = flat.astype(np.uint16).view(np.float16)
        flat = flat.astype(dt)
    return dtype, shape, flat


class CheckpointReaderV1:
    """One-file V1 checkpoint (tensorflow::checkpoint::TensorSliceReader format), read-only."""

    def __init__(self, path, verify=True):
        self.prefix = path
        self.meta, self._values = {}, {}
        with _malformed(path):
            self._load(path, verify)
        if not self.meta:
            raise CheckpointError("%s: no SavedTensorSliceMeta entry (not a V1 checkpoint?)" % path)

    def _load(self, path, verify):
        for key, value in _read_table(path, verify):
            for f, _, v in _pb_fields(value):
                if key == b"" and f == 1:  # SavedTensorSliceMeta
                    for f2, _, t in _pb_fields(v):
                        if f2 != 1:
                            continue
                        name, shape, dtype = None, (), 0
                        for f3, _, x in _pb_fields(t):
                            if f3 == 1:
                                name = x.decode("utf-8")
                            elif f3 == 2:
                                shape = _parse_shape(x)
                            elif f3 == 3:
                                dtype = x
                        self.meta[name] = (shape, dtype)
                elif key != b"" and f == 2:  # SavedSlice
                    name, ext, tp = None, [], None
                    for f2, _, x in _pb_fields(v):
                        if f2 == 1:
                            name = x.decode("utf-8")
                        elif f2 == 2:
                            ext = _parse_extents(x)
                        elif f2 == 3:
                            tp = x
                    if name is not None and tp is not None:
                        self._values.setdefault(name, []).append((ext, tp))

    def has_tensor(self, name):
        return name in self.meta

    def get_variable_to_shape_map(self):
        return {k: list(v[0]) for k, v in self.meta.items()}

    @property
    def entries(self):
        return self.meta

    def get_tensor(self, name, verify=False):
        if name not in self.meta:
            raise CheckpointError("tensor %r not found in %s" % (name, self.prefix))
        shape, dtype = self.meta[name]
        if dtype not in _DTYPES:
            raise CheckpointError("dtype %d of %r is not supported" % (dtype, name))
        with _malformed("tensor %r of %s" % (name, self.prefix)):
            return self._assemble(name, shape, dtype)

    def _assemble(self, name, shape, dtype):
        if any(d < 0 for d in shape) or int(np.prod(shape, dtype=np.float64)) > (1 << 34):
            raise CheckpointError("implausible shape %s for %r" % (list(shape), name))
        out = np.zeros(shape, _DTYPES[dtype])
        filled = 0
        for ext, tp in self._values.get(name, []):
            _, _, flat = _parse_tensor_proto(tp)
            idx, sub = [], []
            for d, size in enumerate(shape):
                start, length = ext[d] if d < len(ext) else (0, None)
                length = size - start if length is None else length
                if start < 0 or length < 0 or start + length > size:
                    raise CheckpointError("slice of %r lies outside the tensor" % name)
                idx.append(slice(start, start + length))
                sub.append(length)
            if int(np.prod(sub)) != flat.size:
                raise CheckpointError("slice of %r does not match its extents" % name)
            out[tuple(idx)] = flat.reshape(sub)
            filled += flat.size
        if filled != out.size:
            raise CheckpointError("slices of %r do not cover the tensor" % name)
        return out


def open_checkpoint(path):
    """V2 bundle prefix or V1 single file -> reader with has_tensor / get_tensor / get_variable_to_shape_map."""
    if os.path.exists(path + ".index"):
        return CheckpointReader(path)
    if os.path.isfile(path):
        return CheckpointReaderV1(path)
    raise CheckpointError("no checkpoint at %r (expected %s.index or a V1 file)" % (path, path))


def is_checkpoint(path):
    if os.path.exists(path + ".index"):
        return True
    if os.path.isfile(path) and os.path.getsize(path) >= 48:
        with open(path, "rb") as f:
            f.seek(-8, os.SEEK_END)
            return struct.unpack("<Q", f.read(8))[0] == TABLE_MAGIC
    return False


def read_checkpoint(prefix, verify=False):
    r = open_checkpoint(prefix)
    return {k: r.get_tensor(k, verify) for k in r.entries}


def write_checkpoint(prefix, tensors):
    """Write {name: array} as a one-shard V2 bundle + update the `checkpoint` state file
    the way `Saver.save` does (SSD300.py:499)."""
    d = os.path.dirname(prefix)
    if d:
        os.makedirs(d, exist_ok=True)
    items = [(b"", _encode_header(1))]
    offset = 0
    with open(_shard_path(prefix, 0, 1), "wb") as f:
        for name in sorted(tensors, key=lambda s: s.encode("utf-8")):
            a = np.asarray(tensors[name], order="C")  # (ascontiguousarray would turn scalars into shape (1,))
            if a.dtype not in _DTYPE_OF:
                raise CheckpointError("dtype %s of %r cannot be stored" % (a.dtype, name))
            raw = a.astype(a.dtype.newbyteorder("<")).tobytes()
            f.write(raw)
            items.append((name.encode("utf-8"),
                          _encode_entry(_DTYPE_OF[a.dtype], a.shape, 0, offset, len(raw), mask_crc(crc32c(raw)))))
            offset += len(raw)
    _write_table(prefix + ".index", items)
    base = os.path.basename(prefix)
    with open(os.path.join(d or ".", "checkpoint"), "w") as f:
        f.write('model_checkpoint_path: "%s"\nall_model_checkpoint_paths: "%s"\n' % (base, base))
    return prefix


def latest_checkpoint(directory):
    """tf.train.latest_checkpoint: the prefix named by the `checkpoint` state file."""
    p = os.path.join(directory, "checkpoint")
    if not os.path.exists(p):
        return None
    with open(p) as f:
        for ln in f:
            if ln.startswith("model_checkpoint_path:"):
                return os.path.join(directory, ln.split(":", 1)[1].strip().strip('"'))
    return None
