"""TensorFlow tensor-bundle checkpoint reader / writer (deepfluids_b200/tf_checkpoint.py): known-answer vectors of the
published formats (RFC 3720 CRC-32C, leveldb crc32c_test / table format) and round trips.  TensorFlow itself is not
installable here, so there is no cross-check against a TF-written file (stated in the module header)."""
import os
import struct

import numpy as np
import pytest

from deepfluids_b200 import tf_checkpoint as tfc


def test_crc32c_known_answers():
    # RFC 3720 B.4 / leveldb util/crc32c_test.cc "StandardResults"
    assert tfc.crc32c(b"\x00" * 32) == 0x8A9136AA
    assert tfc.crc32c(b"\xff" * 32) == 0x62A8AB43
    assert tfc.crc32c(bytes(range(32))) == 0x46DD794E
    assert tfc.crc32c(bytes(range(31, -1, -1))) == 0x113FDB5C
    assert tfc.crc32c(b"123456789") == 0xE3069283
    iscsi_read = bytes([0x01, 0xC0, 0x00, 0x00, 0x00, 0x00, 0x00, 0x00, 0x00, 0x00, 0x00, 0x00, 0x00, 0x00, 0x00, 0x00,
                        0x14, 0x00, 0x00, 0x00, 0x00, 0x00, 0x04, 0x00, 0x00, 0x00, 0x00, 0x14, 0x00, 0x00, 0x00, 0x18,
                        0x28, 0x00, 0x00, 0x00, 0x00, 0x00, 0x00, 0x00, 0x02, 0x00, 0x00, 0x00, 0x00, 0x00, 0x00, 0x00])
    assert tfc.crc32c(iscsi_read) == 0xD9963A56


def test_crc32c_lane_path_equals_bytewise():
    rng = np.random.default_rng(0)
    for n in (2048 * 64, 2048 * 64 + 1, 1_000_003):
        data = rng.integers(0, 256, n, dtype=np.uint8).tobytes()
        assert tfc.crc32c(data) == tfc._raw_update(0xFFFFFFFF, data) ^ 0xFFFFFFFF


def test_crc_mask_roundtrip():
    crc = tfc.crc32c(b"foo")
    assert tfc.mask_crc(crc) != crc and tfc.mask_crc(tfc.mask_crc(crc)) != crc
    assert tfc.unmask_crc(tfc.mask_crc(crc)) == crc
    assert tfc.unmask_crc(tfc.unmask_crc(tfc.mask_crc(tfc.mask_crc(crc)))) == crc


def test_bundle_roundtrip_and_file_layout(tmp_path):
    rng = np.random.default_rng(1)
    t = {"G/0_fc/weights": rng.standard_normal((3, 640)).astype(np.float32),
         "G/1_conv/weights": rng.standard_normal((3, 3, 3, 8, 16)).astype(np.float32),
         "G/1_conv/biases": np.zeros(16, np.float32),
         "step": np.int32(1234), "g_lr": np.float32(2.5e-5), "flag": np.array([True, False]),
         "i64": np.arange(5, dtype=np.int64), "half": rng.standard_normal(7).astype(np.float16), "empty": np.zeros((0, 4), np.float32)}
    prefix = str(tmp_path / "model.ckpt-1234")
    tfc.write_checkpoint(prefix, t)
    assert sorted(os.listdir(tmp_path)) == ["model.ckpt-1234.data-00000-of-00001", "model.ckpt-1234.index"]
    idx = open(prefix + ".index", "rb").read()
    assert struct.unpack("<Q", idx[-8:])[0] == 0xDB4775248B80FB57            # leveldb table magic
    data = open(prefix + ".data-00000-of-00001", "rb").read()
    assert len(data) == sum(np.asarray(v).nbytes for v in t.values())          # tensors back to back
    names = [n for n, _, _ in tfc.list_variables(prefix)]
    assert names == sorted(t, key=lambda k: k.encode())                        # byte-ordered keys
    first = names[0]
    assert data[:np.asarray(t[first]).nbytes] == np.asarray(t[first]).tobytes()
    back = tfc.read_checkpoint(prefix)
    assert set(back) == set(t)
    for k, v in t.items():
        assert back[k].dtype == np.asarray(v).dtype and back[k].shape == np.asarray(v).shape
        np.testing.assert_array_equal(back[k], np.asarray(v))
    only = tfc.read_checkpoint(prefix, ["step"])
    assert list(only) == ["step"] and int(only["step"]) == 1234
    with pytest.raises(KeyError):
        tfc.read_checkpoint(prefix, ["nope"])


def test_many_keys_span_blocks_and_restarts(tmp_path, monkeypatch):
    """> 16 keys (restart interval) with long shared prefixes; a small block size makes the index span many data blocks
    (index block with shortest-separator keys) the way a 256 KB block size does for a very large model"""
    monkeypatch.setattr(tfc, "_BLOCK_SIZE", 4096)
    t = {"scope/%s/layer_%05d/weights" % ("x" * 60, i): np.full((2,), i, np.float32) for i in range(3000)}
    prefix = str(tmp_path / "big")
    tfc.write_checkpoint(prefix, t)
    assert os.path.getsize(prefix + ".index") > 20 * 4096
    back = tfc.read_checkpoint(prefix)
    assert len(back) == 3000
    for i in (0, 15, 16, 17, 1499, 2999):
        k = "scope/%s/layer_%05d/weights" % ("x" * 60, i)
        np.testing.assert_array_equal(back[k], t[k])


def test_corruption_is_detected(tmp_path):
    prefix = str(tmp_path / "c")
    tfc.write_checkpoint(prefix, {"a": np.arange(100, dtype=np.float32)})
    raw = bytearray(open(prefix + ".data-00000-of-00001", "rb").read())
    raw[17] ^= 0x40
    open(prefix + ".data-00000-of-00001", "wb").write(raw)
    with pytest.raises(ValueError, match="checksum"):
        tfc.read_checkpoint(prefix)
    assert tfc.read_checkpoint(prefix, verify=False)["a"].shape == (100,)
    idx = bytearray(open(prefix + ".index", "rb").read())
    idx[3] ^= 0x01
    open(prefix + ".index", "wb").write(idx)
    with pytest.raises(ValueError, match="checksum"):
        tfc.list_variables(prefix)
    open(prefix + ".index", "wb").write(b"not a table" * 10)
    with pytest.raises(ValueError, match="magic"):
        tfc.list_variables(prefix)


def test_snappy_block_decoder():
    # literal "abcd" + copy(offset 4, len 8) -> "abcdabcdabcd": 1-byte-offset copy tag = (len-4)<<2 | 1, offset low byte
    comp = bytes([12, (4 - 1) << 2]) + b"abcd" + bytes([((8 - 4) << 2) | 1, 4])
    assert tfc._snappy_decompress(comp) == b"abcdabcdabcd"


def test_checkpoint_state_file(tmp_path):
    d = str(tmp_path)
    assert tfc.latest_checkpoint(d) is None
    for step in (10, 20):
        tfc.write_checkpoint(os.path.join(d, "model.ckpt-%d" % step), {"step": np.int32(step)})
        tfc.update_checkpoint_state(d, "model.ckpt-%d" % step)
    txt = open(os.path.join(d, "checkpoint")).read().splitlines()
    assert txt == ['model_checkpoint_path: "model.ckpt-20"', 'all_model_checkpoint_paths: "model.ckpt-10"',
                   'all_model_checkpoint_paths: "model.ckpt-20"']
    assert tfc.latest_checkpoint(d) == os.path.join(d, "model.ckpt-20")


def test_adam_slot_naming_and_power_accumulators():
    var = {"G/0_fc/weights": (2, 3)}
    m = {"G/0_fc/weights": np.ones((2, 3), np.float32)}
    v = {"G/0_fc/weights": np.full((2, 3), 2, np.float32)}
    out = tfc.adam_state_to_tf(var, m, v, adam_t=7, beta1=0.5, beta2=0.999)
    assert set(out) == {"G/0_fc/weights/Adam", "G/0_fc/weights/Adam_1", "beta1_power", "beta2_power"}
    # tf.train.AdamOptimizer: beta1_power starts at beta1 and is multiplied by beta1 after every apply -> beta1^(t+1)
    assert out["beta1_power"] == np.float32(0.5 ** 8) and out["beta2_power"] == np.float32(0.999 ** 8)
    for t in (0, 1, 7, 40):
        assert tfc.adam_t_from_tf(0.5 ** (t + 1), 0.5) == t
    for t in (0, 130, 4321, 60000):       # beta1_power underflows float32 beyond t ~ 126: beta2_power carries the count
        assert tfc.adam_t_from_tf(float(np.float32(0.5 ** (t + 1))), 0.5, float(np.float32(0.999 ** (t + 1))), 0.999) == t
    assert tfc.adam_t_from_tf(0.0, 0.5, 0.0, 0.999) >= 10 ** 6


def test_trainer_glue_roundtrip_on_cpu(tmp_path):
    """save_tf / load_tf of the Trainer on a CPU stand-in (FlatParams works on any device; no kernels are called)."""
    import torch
    from collections import OrderedDict
    from deepfluids_b200.engine import FlatParams
    from deepfluids_b200.trainer import Trainer

    class Eng(object):
        def __init__(self):
            self.params = FlatParams(OrderedDict([("G/0_fc/weights", (3, 64)), ("G/0_fc/biases", (64,)),
                                                  ("G/1_conv/weights", (3, 3, 4, 5)), ("G/1_conv/biases", (5,))]), "cpu")
            self.adam_t = 0
            self.repacked = 0

        def repack(self):
            self.repacked += 1

    def make(seed):
        tr = Trainer.__new__(Trainer)
        tr.engine, tr.optimizer, tr.beta1, tr.beta2, tr.step, tr.g_lr = Eng(), "adam", 0.5, 0.999, 0, 1e-4
        g = torch.Generator().manual_seed(seed)
        for buf in (tr.engine.params.data, tr.engine.params.m, tr.engine.params.v):
            buf.copy_(torch.randn(buf.shape, generator=g))
        return tr

    a, b = make(1), make(2)
    a.step, a.g_lr, a.engine.adam_t = 4321, 3.25e-5, 4321
    prefix = a.save_tf(str(tmp_path))
    assert prefix.endswith("model.ckpt-4321") and tfc.latest_checkpoint(str(tmp_path)) == prefix
    names = {n for n, _, _ in tfc.list_variables(prefix)}
    assert {"G/1_conv/weights", "G/1_conv/weights/Adam", "G/1_conv/weights/Adam_1", "beta1_power", "beta2_power", "step", "g_lr"} <= names
    b.load_tf(prefix)
    for k in a.engine.params.table:
        assert torch.equal(a.engine.params.p(k), b.engine.params.p(k))
        assert torch.equal(a.engine.params._view(a.engine.params.m, k), b.engine.params._view(b.engine.params.m, k))
        assert torch.equal(a.engine.params._view(a.engine.params.v, k), b.engine.params._view(b.engine.params.v, k))
    assert b.step == 4321 and abs(b.g_lr - 3.25e-5) < 1e-12 and b.engine.repacked == 1
    assert b.engine.adam_t == 4321        # from beta2_power (beta1^4322 underflows float32)
    # a weights-only checkpoint (what a converter would write) restores variables and leaves the optimizer fresh
    tfc.write_checkpoint(str(tmp_path / "w" / "model.ckpt-0"), {k: a.engine.params.p(k).numpy() for k in a.engine.params.table})
    c = make(3)
    m_before = c.engine.params.m.clone()
    c.load_tf(str(tmp_path / "w" / "model.ckpt-0"))
    assert torch.equal(c.engine.params.p("G/1_conv/weights"), a.engine.params.p("G/1_conv/weights"))
    assert torch.equal(c.engine.params.m, m_before)
    with pytest.raises(ValueError, match="shape"):
        tfc.write_checkpoint(str(tmp_path / "bad"), {k: np.zeros((1,) + tuple(s), np.float32) for k, s in a.engine.params.table.items()})
        c.load_tf(str(tmp_path / "bad"))


def test_reader_parses_a_hand_assembled_table(tmp_path):
    """A leveldb table assembled byte by byte from the format description (table_format.md), NOT through the module's
    writer: one data block with prefix-compressed keys and two restart points, an index block, an empty metaindex block, the
    48-byte footer.  Guards against a symmetric mistake in writer + reader."""
    def entry(shared, key_delta, value):
        return bytes([shared, len(key_delta), len(value)]) + key_delta + value          # all lengths < 128: 1-byte varints

    e1 = entry(0, b"abc", b"v1")
    e2 = entry(2, b"d", b"value-2")              # "abd" shares "ab"
    e3 = entry(0, b"b", b"")                     # restart point (shared = 0), empty value
    data = e1 + e2 + e3 + struct.pack("<III", 0, len(e1) + len(e2), 2)                 # restarts [0, off(e3)], count 2

    def with_trailer(block):
        return block + b"\x00" + struct.pack("<I", tfc.mask_crc(tfc.crc32c(block + b"\x00")))

    blob = with_trailer(data)
    meta_off = len(blob)
    meta = struct.pack("<II", 0, 1)                                                     # no entries, one restart at 0
    blob += with_trailer(meta)
    index_off = len(blob)
    handle = bytes([0, len(data)])                                                      # offset 0, size (varints < 128)
    index = entry(0, b"c", handle) + struct.pack("<II", 0, 1)                           # separator key "c" >= last key "b"
    blob += with_trailer(index)
    footer = bytes([meta_off, len(meta)]) + bytes([index_off, len(index)])
    assert meta_off < 128 and index_off < 128
    blob += footer + b"\x00" * (40 - len(footer)) + struct.pack("<Q", 0xDB4775248B80FB57)
    path = str(tmp_path / "hand.index")
    open(path, "wb").write(blob)
    assert tfc._read_table(path) == [(b"abc", b"v1"), (b"abd", b"value-2"), (b"b", b"")]
    # and the writer's output for the same items has the same logical content and a valid structure
    tfc._write_table(str(tmp_path / "w.index"), [(b"abc", b"v1"), (b"abd", b"value-2"), (b"b", b"")])
    assert tfc._read_table(str(tmp_path / "w.index")) == [(b"abc", b"v1"), (b"abd", b"value-2"), (b"b", b"")]
    w = open(str(tmp_path / "w.index"), "rb").read()
    assert w[:len(e1) + len(e2)] == e1 + e2                                             # same prefix compression as by hand


def test_bundle_protos_encode_as_specified():
    """BundleHeaderProto / BundleEntryProto bytes against the protobuf wire format written out by hand (tensor_bundle.proto:
    header num_shards=1, version.producer=1; entry dtype=1 (DT_FLOAT), shape [3,4], offset 48, size 48, crc32c fixed32)."""
    assert tfc._encode_header(1) == bytes([0x08, 0x01, 0x1A, 0x02, 0x08, 0x01])
    e = tfc._encode_entry(1, (3, 4), 0, 48, 48, 0xDEADBEEF)
    want = bytes([0x08, 0x01,                                   # dtype = DT_FLOAT
                  0x12, 0x08, 0x12, 0x02, 0x08, 0x03, 0x12, 0x02, 0x08, 0x04,   # shape { dim { size: 3 } dim { size: 4 } }
                  0x20, 0x30,                                   # offset = 48
                  0x28, 0x30,                                   # size = 48
                  0x35, 0xEF, 0xBE, 0xAD, 0xDE])                # crc32c (fixed32, little endian)
    assert e == want
    d = tfc._decode_entry(want)
    assert d["dtype"] == 1 and d["shape"] == (3, 4) and d["offset"] == 48 and d["size"] == 48 and d["crc32c"] == 0xDEADBEEF
