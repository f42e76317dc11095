"""TensorFlow checkpoint (tensor-bundle V2) reader / writer in pure Python + numpy -- no TensorFlow needed.

The reference saves and restores its variables with `tf.train.Saver` / `tf.train.Supervisor` (trainer.py:107-123 restore
of the latest checkpoint in `model_dir`, trainer.py:291-292 / trainer3.py:178-179 `saver.save(sess, model_dir/model.ckpt,
global_step=step)`).  What lands on disk is TensorFlow's tensor bundle:

    <dir>/checkpoint                               CheckpointState text proto: model_checkpoint_path: "model.ckpt-<step>"
    <dir>/model.ckpt-<step>.index                  a leveldb-format table:  ""   -> BundleHeaderProto
                                                                           name -> BundleEntryProto (dtype, shape, shard, offset, size, crc32c)
    <dir>/model.ckpt-<step>.data-00000-of-00001    the tensors' raw little-endian bytes, back to back in key order

TensorFlow 1.15 is not installable here, so this module restates the published formats (leveldb `table_format.md`: blocks
of prefix-compressed entries + restart array, 5-byte trailer = compression type + masked CRC-32C, index block, 48-byte
footer with magic 0xdb4775248b80fb57; `tensor_bundle.proto`; RFC 3720 CRC-32C) and is checked against their known-answer
vectors (tests/test_tf_checkpoint.py) -- it has NOT been checked against a file written by TensorFlow itself.
Variable names and layouts need no conversion: the engines already keep the reference's names (`G/3_conv/weights`) and TF
layouts (HWIO / DHWIO convolutions, [in, out] FC), and Adam's slots are `<var>/Adam`, `<var>/Adam_1`, `beta1_power`,
`beta2_power` as tf.train.AdamOptimizer names them."""
import os
import struct

import numpy as np

# ---------------------------------------------------------------------------------------------- CRC-32C (Castagnoli)
_POLY = 0x82F63B78


def _make_table():
    t = np.zeros(256, dtype=np.uint32)
    for i in range(256):
        c = i
        for _ in range(8):
            c = (c >> 1) ^ _POLY if c & 1 else c >> 1
        t[i] = c
    return t


_TABLE = _make_table()
_TABLE_L = [int(v) for v in _TABLE]


def _raw_update(state, data):
    """CRC register after feeding `data` (bytes) into register value `state` (no init / final xor)."""
    tab = _TABLE_L
    for b in data:
        state = tab[(state ^ b) & 0xFF] ^ (state >> 8)
    return state


def _zeros_operator(nbytes):
    """32 x 32 GF(2) matrix (as 32 column words) that advances the CRC register through `nbytes` zero bytes."""
    def times(mat, vec):
        s, i = 0, 0
        while vec:
            if vec & 1:
                s ^= mat[i]
            vec >>= 1
            i += 1
        return s

    def square(mat):
        return [times(mat, mat[i]) for i in range(32)]
    # operator for one zero BIT: register -> (register >> 1) ^ (POLY if lsb)
    op = [_POLY] + [1 << (i - 1) for i in range(1, 32)]
    result = [1 << i for i in range(32)]           # identity
    n = nbytes * 8
    while n:
        if n & 1:
            result = [times(op, result[i]) for i in range(32)]
        op = square(op)
        n >>= 1
    return result, times


def crc32c(data):
    """CRC-32C of a bytes-like object.  Large inputs are split into lanes that numpy advances in lock step; the lane
    registers are then chained with the 'advance through n zero bytes' operator (CRC is linear over GF(2))."""
    buf = np.frombuffer(memoryview(data).cast("B"), dtype=np.uint8)
    n = buf.size
    lanes = 2048
    if n < lanes * 64:
        return _raw_update(0xFFFFFFFF, buf.tobytes()) ^ 0xFFFFFFFF
    L = n // lanes
    body = buf[:L * lanes].reshape(lanes, L)
    reg = np.zeros(lanes, dtype=np.uint32)
    for j in range(L):
        reg = _TABLE[(reg ^ body[:, j]) & 0xFF] ^ (reg >> np.uint32(8))
    op, times = _zeros_operator(L)
    state = 0xFFFFFFFF
    for r in reg.tolist():
        state = times(op, state) ^ r
    state = _raw_update(state, buf[L * lanes:].tobytes())
    return state ^ 0xFFFFFFFF


_MASK_DELTA = 0xA282EAD8


def mask_crc(crc):
    """leveldb / TensorFlow store CRCs 'masked' (crc32c.h: rotate right by 15, add a constant)"""
    return ((((crc >> 15) | (crc << 17)) & 0xFFFFFFFF) + _MASK_DELTA) & 0xFFFFFFFF


def unmask_crc(masked):
    rot = (masked - _MASK_DELTA) & 0xFFFFFFFF
    return ((rot >> 17) | (rot << 15)) & 0xFFFFFFFF


# ---------------------------------------------------------------------------------------------- varints / protobuf wire
def _varint(n):
    out = bytearray()
    n &= (1 << 64) - 1
    while True:
        b = n & 0x7F
        n >>= 7
        if n:
            out.append(b | 0x80)
        else:
            out.append(b)
            return bytes(out)


def _read_varint(buf, pos):
    shift = result = 0
    while True:
        b = buf[pos]
        pos += 1
        result |= (b & 0x7F) << shift
        if not b & 0x80:
            return result, pos
        shift += 7


def _pb_fields(buf):
    """yield (field number, wire type, value) of a serialized message; value = int (varint / fixed) or bytes"""
    pos, n = 0, len(buf)
    while pos < n:
        key, pos = _read_varint(buf, pos)
        f, wt = key >> 3, key & 7
        if wt == 0:
            v, pos = _read_varint(buf, pos)
        elif wt == 1:
            v = struct.unpack_from("<Q", buf, pos)[0]
            pos += 8
        elif wt == 2:
            ln, pos = _read_varint(buf, pos)
            v = bytes(buf[pos:pos + ln])
            pos += ln
        elif wt == 5:
            v = struct.unpack_from("<I", buf, pos)[0]
            pos += 4
        else:
            raise ValueError("unsupported protobuf wire type %d" % wt)
        yield f, wt, v


def _pb_varint_field(f, v):
    return _varint((f << 3) | 0) + _varint(v)


def _pb_bytes_field(f, b):
    return _varint((f << 3) | 2) + _varint(len(b)) + b


# TensorFlow DataType enum (types.proto) <-> numpy
_DT = {1: np.dtype("<f4"), 2: np.dtype("<f8"), 3: np.dtype("<i4"), 4: np.dtype("u1"), 5: np.dtype("<i2"), 6: np.dtype("i1"),
       9: np.dtype("<i8"), 10: np.dtype("?"), 17: np.dtype("<u2"), 19: np.dtype("<f2"), 22: np.dtype("<u4"), 23: np.dtype("<u8")}
_DT_BFLOAT16 = 14
_NP2DT = {v: k for k, v in _DT.items()}


def _encode_shape(shape):                       # TensorShapeProto: repeated Dim dim = 2 { int64 size = 1 }
    return b"".join(_pb_bytes_field(2, _pb_varint_field(1, int(d))) for d in shape)


def _decode_shape(buf):
    dims = []
    for f, _, v in _pb_fields(buf):
        if f == 2:
            size = 0
            for ff, _, vv in _pb_fields(v):
                if ff == 1:
                    size = vv
            dims.append(size)
        elif f == 3 and v:
            raise ValueError("tensor of unknown rank in checkpoint")
    return tuple(dims)


def _encode_entry(dtype_enum, shape, shard, offset, size, crc_masked):       # BundleEntryProto
    out = _pb_varint_field(1, dtype_enum) + _pb_bytes_field(2, _encode_shape(shape))
    if shard:
        out += _pb_varint_field(3, shard)
    if offset:
        out += _pb_varint_field(4, offset)
    out += _pb_varint_field(5, size)
    out += _varint((6 << 3) | 5) + struct.pack("<I", crc_masked)
    return out


def _decode_entry(buf):
    e = {"dtype": 0, "shape": (), "shard_id": 0, "offset": 0, "size": 0, "crc32c": None, "slices": 0}
    for f, _, v in _pb_fields(buf):
        if f == 1:
            e["dtype"] = v
        elif f == 2:
            e["shape"] = _decode_shape(v)
        elif f == 3:
            e["shard_id"] = v
        elif f == 4:
            e["offset"] = v
        elif f == 5:
            e["size"] = v
        elif f == 6:
            e["crc32c"] = v
        elif f == 7:
            e["slices"] += 1
    return e


def _encode_header(num_shards=1):                # BundleHeaderProto: num_shards = 1, endianness = 2 (LITTLE = 0), version = 3
    version = _pb_varint_field(1, 1)             # VersionDef.producer = kTensorBundleVersion = 1
    return _pb_varint_field(1, num_shards) + _pb_bytes_field(3, version)


def _decode_header(buf):
    h = {"num_shards": 1, "endianness": 0}
    for f, _, v in _pb_fields(buf):
        if f == 1:
            h["num_shards"] = v
        elif f == 2:
            h["endianness"] = v
    return h


# ---------------------------------------------------------------------------------------------- leveldb table
_MAGIC = 0xDB4775248B80FB57
_RESTART_INTERVAL = 16
_BLOCK_SIZE = 262144


class _BlockBuilder(object):
    def __init__(self):
        self.buf = bytearray()
        self.restarts = [0]
        self.count = 0
        self.last_key = b""

    def add(self, key, value):
        shared = 0
        if self.count and self.count % _RESTART_INTERVAL == 0:
            self.restarts.append(len(self.buf))
        elif self.count:
            m = min(len(key), len(self.last_key))
            while shared < m and key[shared] == self.last_key[shared]:
                shared += 1
        self.buf += _varint(shared) + _varint(len(key) - shared) + _varint(len(value)) + key[shared:] + value
        self.last_key = key
        self.count += 1

    def size(self):
        return len(self.buf) + 4 * len(self.restarts) + 4

    def finish(self):
        return bytes(self.buf) + b"".join(struct.pack("<I", r) for r in self.restarts) + struct.pack("<I", len(self.restarts))


def _shortest_separator(start, limit):
    m = min(len(start), len(limit))
    d = 0
    while d < m and start[d] == limit[d]:
        d += 1
    if d < m and start[d] < 0xFF and start[d] + 1 < limit[d]:
        return start[:d] + bytes([start[d] + 1])
    return start


def _short_successor(key):
    for i, b in enumerate(key):
        if b != 0xFF:
            return key[:i] + bytes([b + 1])
    return key


def _write_table(path, items):
    """items: sorted list of (key bytes, value bytes)"""
    out = bytearray()

    def emit(block):
        handle = _varint(len(out)) + _varint(len(block))
        out.extend(block)
        out.extend(b"\x00" + struct.pack("<I", mask_crc(crc32c(block + b"\x00"))))      # no compression
        return handle

    index = _BlockBuilder()
    data = _BlockBuilder()
    pending = None                                    # (last key of the finished block, its handle)
    for key, value in items:
        if pending is not None:
            index.add(_shortest_separator(pending[0], key), pending[1])
            pending = None
        data.add(key, value)
        if data.size() >= _BLOCK_SIZE:
            pending = (data.last_key, emit(data.finish()))
            data = _BlockBuilder()
    if data.count:
        pending = (data.last_key, emit(data.finish()))
    if pending is not None:
        index.add(_short_successor(pending[0]), pending[1])
    meta_handle = emit(_BlockBuilder().finish())
    index_handle = emit(index.finish())
    footer = meta_handle + index_handle
    footer += b"\x00" * (40 - len(footer)) + struct.pack("<Q", _MAGIC)
    out.extend(footer)
    with open(path, "wb") as f:
        f.write(out)


def _snappy_decompress(buf):
    n, pos = _read_varint(buf, 0)
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
            out += buf[pos:pos + ln]
            pos += ln
            continue
        if kind == 1:
            ln = ((tag >> 2) & 7) + 4
            off = ((tag >> 5) << 8) | buf[pos]
            pos += 1
        elif kind == 2:
            ln = (tag >> 2) + 1
            off = int.from_bytes(buf[pos:pos + 2], "little")
            pos += 2
        else:
            ln = (tag >> 2) + 1
            off = int.from_bytes(buf[pos:pos + 4], "little")
            pos += 4
        if off == 0 or off > len(out):
            raise ValueError("corrupt snappy block")
        for _ in range(ln):                           # overlapping copies are allowed
            out.append(out[-off])
    if len(out) != n:
        raise ValueError("corrupt snappy block (length)")
    return bytes(out)


def _read_block(buf, offset, size, verify=True):
    block = bytes(buf[offset:offset + size])
    ctype = buf[offset + size]
    stored = struct.unpack_from("<I", buf, offset + size + 1)[0]
    if verify and unmask_crc(stored) != crc32c(block + bytes([ctype])):
        raise ValueError("table block checksum mismatch at offset %d" % offset)
    if ctype == 1:
        block = _snappy_decompress(block)
    elif ctype != 0:
        raise ValueError("unknown block compression type %d" % ctype)
    return block


def _block_entries(block):
    nrestarts = struct.unpack_from("<I", block, len(block) - 4)[0]
    end = len(block) - 4 - 4 * nrestarts
    pos, key = 0, b""
    while pos < end:
        shared, pos = _read_varint(block, pos)
        unshared, pos = _read_varint(block, pos)
        vlen, pos = _read_varint(block, pos)
        key = key[:shared] + block[pos:pos + unshared]
        pos += unshared
        yield key, block[pos:pos + vlen]
        pos += vlen


def _read_table(path, verify=True):
    with open(path, "rb") as f:
        buf = f.read()
    if len(buf) < 48 or struct.unpack_from("<Q", buf, len(buf) - 8)[0] != _MAGIC:
        raise ValueError("%s is not a TensorFlow checkpoint index (bad table magic)" % path)
    footer = buf[len(buf) - 48:]
    _, pos = _read_varint(footer, 0)                  # metaindex handle (unused)
    _, pos = _read_varint(footer, pos)
    ioff, pos = _read_varint(footer, pos)
    isize, pos = _read_varint(footer, pos)
    items = []
    for _, handle in _block_entries(_read_block(buf, ioff, isize, verify)):
        off, p2 = _read_varint(handle, 0)
        size, _ = _read_varint(handle, p2)
        items.extend(_block_entries(_read_block(buf, off, size, verify)))
    return items


# ---------------------------------------------------------------------------------------------- public API
def _data_name(prefix, shard, num_shards):
    return "%s.data-%05d-of-%05d" % (prefix, shard, num_shards)


def write_checkpoint(prefix, tensors):
    """Write {name: numpy array} as `<prefix>.index` + `<prefix>.data-00000-of-00001` (what tf.train.Saver.save produces
    for `prefix`, minus the .meta graph file).  Keys are written in byte order, tensors back to back."""
    names = sorted(tensors, key=lambda k: k.encode("utf-8"))
    items = [(b"", _encode_header(1))]
    offset = 0
    os.makedirs(os.path.dirname(os.path.abspath(prefix)), exist_ok=True)
    with open(_data_name(prefix, 0, 1), "wb") as f:
        for name in names:
            if name == "":
                raise ValueError("the empty name is reserved for the bundle header")
            a = np.asarray(tensors[name])
            dt = a.dtype.newbyteorder("<") if a.dtype.byteorder == ">" else a.dtype
            if dt not in _NP2DT:
                raise TypeError("dtype %s of %r has no TensorFlow checkpoint encoding here" % (a.dtype, name))
            raw = np.ascontiguousarray(a, dtype=dt).tobytes()
            f.write(raw)
            items.append((name.encode("utf-8"), _encode_entry(_NP2DT[dt], a.shape, 0, offset, len(raw), mask_crc(crc32c(raw)))))
            offset += len(raw)
    _write_table(prefix + ".index", items)


def list_variables(prefix):
    """[(name, shape, numpy dtype)] of a checkpoint, in key order (tf.train.list_variables)."""
    out = []
    for key, value in _read_table(prefix + ".index"):
        if key == b"":
            continue
        e = _decode_entry(value)
        out.append((key.decode("utf-8"), e["shape"], _DT.get(e["dtype"], "bfloat16" if e["dtype"] == _DT_BFLOAT16 else None)))
    return out


def read_checkpoint(prefix, names=None, verify=True):
    """{name: numpy array} of `<prefix>.index` / `.data-*` (all variables, or `names`).  bfloat16 tensors come back as
    float32.  Raises ValueError on checksum mismatch, sliced (partitioned) variables or big-endian bundles."""
    items = _read_table(prefix + ".index", verify)
    header = {"num_shards": 1, "endianness": 0}
    entries = {}
    for key, value in items:
        if key == b"":
            header = _decode_header(value)
        else:
            entries[key.decode("utf-8")] = _decode_entry(value)
    if header["endianness"] != 0:
        raise ValueError("big-endian tensor bundles are not supported")
    want = list(entries) if names is None else list(names)
    shards = {}
    out = {}
    for name in want:
        if name not in entries:
            raise KeyError("variable %r not in checkpoint %s" % (name, prefix))
        e = entries[name]
        if e["slices"]:
            raise ValueError("partitioned variable %r is not supported" % name)
        if e["shard_id"] not in shards:
            shards[e["shard_id"]] = np.memmap(_data_name(prefix, e["shard_id"], header["num_shards"]), dtype=np.uint8, mode="r")
        raw = shards[e["shard_id"]][e["offset"]:e["offset"] + e["size"]]
        if raw.size != e["size"]:
            raise ValueError("data shard too short for %r" % name)
        if verify and e["crc32c"] is not None and unmask_crc(e["crc32c"]) != crc32c(raw):
            raise ValueError("tensor checksum mismatch for %r" % name)
        if e["dtype"] == _DT_BFLOAT16:
            a = (np.frombuffer(raw.tobytes(), dtype="<u2").astype(np.uint32) << 16).view(np.float32)
        elif e["dtype"] in _DT:
            a = np.frombuffer(raw.tobytes(), dtype=_DT[e["dtype"]])
        else:
            raise TypeError("TensorFlow dtype enum %d of %r is not supported" % (e["dtype"], name))
        out[name] = a.reshape(e["shape"]).copy()
    return out


def update_checkpoint_state(directory, basename):
    """the `checkpoint` file tf.train.Saver maintains beside the bundles (CheckpointState text proto)"""
    path = os.path.join(directory, "checkpoint")
    older = []
    if os.path.exists(path):
        for line in open(path):
            if line.startswith("all_model_checkpoint_paths:"):
                older.append(line.split(":", 1)[1].strip().strip('"'))
    allp = [p for p in older if p != basename] + [basename]
    with open(path, "w") as f:
        f.write('model_checkpoint_path: "%s"\n' % basename)
        for p in allp:
            f.write('all_model_checkpoint_paths: "%s"\n' % p)


def latest_checkpoint(directory):
    """tf.train.latest_checkpoint: prefix named by `<directory>/checkpoint`, or None"""
    path = os.path.join(directory, "checkpoint")
    if not os.path.exists(path):
        return None
    for line in open(path):
        if line.startswith("model_checkpoint_path:"):
            p = line.split(":", 1)[1].strip().strip('"')
            p = p if os.path.isabs(p) else os.path.join(directory, p)
            return p if os.path.exists(p + ".index") else None
    return None


# ---------------------------------------------------------------------------------------------- trainer glue
def adam_state_to_tf(variables, m, v, adam_t, beta1, beta2):
    """Adam slots as tf.train.AdamOptimizer names them: `<var>/Adam` (m), `<var>/Adam_1` (v) and the two power
    accumulators, which TF initialises to beta and multiplies by beta after every step (so beta^(t+1) after t steps)."""
    out = {}
    for k in variables:
        out[k + "/Adam"] = np.asarray(m[k], dtype=np.float32)
        out[k + "/Adam_1"] = np.asarray(v[k], dtype=np.float32)
    out["beta1_power"] = np.float32(beta1 ** (adam_t + 1))
    out["beta2_power"] = np.float32(beta2 ** (adam_t + 1))
    return out


def adam_t_from_tf(beta1_power, beta1, beta2_power=None, beta2=None):
    """Adam step count t from the saved power accumulators (beta^(t+1)).  beta2_power resolves t far longer than
    beta1_power (0.5^128 underflows float32, 0.999^t only beyond t ~ 87 000); once both have underflowed the bias
    corrections are exactly 1 in TensorFlow as well, which any large t reproduces."""
    import math
    for power, beta in ((beta2_power, beta2), (beta1_power, beta1)):
        if power is not None and beta is not None and 0.0 < beta < 1.0 and float(power) > 0.0:
            return max(0, int(round(math.log(float(power)) / math.log(beta))) - 1)
    return 10 ** 6
