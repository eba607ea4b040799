"""TensorFlow tensor-bundle reader / writer -- host code, no GPU, no TensorFlow.  Parity unpinned
against TF itself (no TF-written file exists here): the reader is checked against hand-assembled LevelDB blocks, the
format's own CRCs and the module's writer."""
import os
import struct

import numpy as np
import pytest

import sr_wavenet_b200  # noqa: F401
from sr_wavenet_b200 import nsynth, synth, tf_checkpoint as tfc


def test_block_prefix_compression_by_hand():
    """LevelDB block: entries share a prefix with the previous key (varint32 shared | non_shared | value_len | suffix |
    value), restart array + count at the end."""
    e = lambda shared, suffix, value: bytes([shared, len(suffix), len(value)]) + suffix + value
    body = e(0, b"Decoder/causal_conv_Bias", b"A") + e(20, b"Kernel", b"BB") + e(8, b"conv1d/bias", b"")
    block = body + struct.pack("<II", 0, 1)
    assert tfc._block_entries(block) == [(b"Decoder/causal_conv_Bias", b"A"), (b"Decoder/causal_conv_Kernel", b"BB"),
                                         (b"Decoder/conv1d/bias", b"")]
    assert tfc._block_entries(tfc._build_block([(b"a", b"1"), (b"ab", b"2"), (b"b", b"3")])) == [(b"a", b"1"), (b"ab", b"2"), (b"b", b"3")]


def test_bundle_entry_by_hand():
    """BundleEntryProto{dtype=1 DT_FLOAT, shape{dim{size:2} dim{size:3}}, offset=24, size=24, crc32c}."""
    raw = (b"\x08\x01" + b"\x12\x08" + b"\x12\x02\x08\x02" + b"\x12\x02\x08\x03" + b"\x20\x18" + b"\x28\x18" +
           b"\x35" + struct.pack("<I", 0xDEADBEEF))
    e = tfc._parse_entry(raw)
    assert (e["dtype"], e["shape"], e["offset"], e["size"], e["crc32c"], e["shard_id"]) == (1, [2, 3], 24, 24, 0xDEADBEEF, 0)
    assert tfc._parse_entry(tfc._enc_entry(1, (2, 3), 24, 24, 0xDEADBEEF)) == e


def test_roundtrip_many_blocks_and_corruption(tmp_path):
    w = synth.make_teacher_weights(synth.DEFAULT_DILATIONS)          # 246 variables: several 4 KB index blocks
    w["global_step"] = np.array(1234, dtype=np.int64)
    prefix = os.path.join(str(tmp_path), "model.ckpt-7")
    tfc.write_checkpoint(prefix, w)
    assert os.path.exists(prefix + ".index") and os.path.exists(prefix + ".data-00000-of-00001")
    raw = open(prefix + ".index", "rb").read()
    assert struct.unpack("<Q", raw[-8:])[0] == 0xdb4775248b80fb57
    back = tfc.read_checkpoint(prefix)
    assert sorted(back) == sorted(w)
    for k in w:
        assert back[k].dtype == np.asarray(w[k]).dtype and back[k].shape == np.asarray(w[k]).shape
        np.testing.assert_array_equal(back[k], w[k])
    names = [n for n, _, _ in tfc.list_variables(prefix)]
    assert names == sorted(names, key=lambda s: s.encode()) and len(names) == len(w)
    # a flipped byte in a tensor is caught by the per-tensor CRC, one in the index by the block CRC
    data = bytearray(open(prefix + ".data-00000-of-00001", "rb").read())
    data[100] ^= 0x40
    open(prefix + ".data-00000-of-00001", "wb").write(data)
    with pytest.raises(IOError):
        tfc.read_checkpoint(prefix)
    assert len(tfc.read_checkpoint(prefix, verify=False)) == len(w)
    idx = bytearray(raw)
    idx[50] ^= 0x01
    open(prefix + ".index", "wb").write(idx)
    with pytest.raises(IOError):
        tfc.read_table(prefix + ".index")
    open(prefix + ".index", "wb").write(raw[:-1])
    with pytest.raises(IOError):
        tfc.read_table(prefix + ".index")


def test_latest_checkpoint(tmp_path):
    d = str(tmp_path)
    assert tfc.latest_checkpoint(d) is None
    open(os.path.join(d, "checkpoint"), "w").write('model_checkpoint_path: "model.ckpt-42"\nall_model_checkpoint_paths: "model.ckpt-42"\n')
    assert tfc.latest_checkpoint(d) == os.path.join(d, "model.ckpt-42")
