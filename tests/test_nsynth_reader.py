"""The TFRecord / tf.train.Example reader that stands in for nsynth.py:5-52 (host I/O, no GPU, no TensorFlow)."""
import os
import struct

import numpy as np
import pytest

import sr_wavenet_b200  # noqa: F401
from sr_wavenet_b200 import nsynth


def test_crc32c_known_answers():
    assert nsynth.crc32c(b"123456789") == 0xE3069283           # the CRC-32C check value
    assert nsynth.crc32c(b"") == 0
    assert nsynth.crc32c(bytes(32)) == 0x8A9136AA               # RFC 3720 B.4: 32 bytes of zeros


def _example_classes():
    """tf.train.Example / Features / Feature rebuilt with the protobuf runtime from the field numbers of TensorFlow's
    example.proto / feature.proto: an independent serializer to check the hand-written parser against."""
    from google.protobuf import descriptor_pb2, descriptor_pool, message_factory
    fd = descriptor_pb2.FileDescriptorProto(name="tf_example_test.proto", package="tftest", syntax="proto3")
    T = descriptor_pb2.FieldDescriptorProto

    def msg(name, fields):
        m = fd.message_type.add(name=name)
        for fname, num, ftype, label, tname, packed in fields:
            f = m.field.add(name=fname, number=num, type=ftype, label=label)
            if tname:
                f.type_name = ".tftest." + tname
            if packed is not None:
                f.options.packed = packed
        return m
    msg("BytesList", [("value", 1, T.TYPE_BYTES, T.LABEL_REPEATED, None, None)])
    msg("FloatList", [("value", 1, T.TYPE_FLOAT, T.LABEL_REPEATED, None, True)])
    msg("Int64List", [("value", 1, T.TYPE_INT64, T.LABEL_REPEATED, None, True)])
    feat = msg("Feature", [("bytes_list", 1, T.TYPE_MESSAGE, T.LABEL_OPTIONAL, "BytesList", None),
                           ("float_list", 2, T.TYPE_MESSAGE, T.LABEL_OPTIONAL, "FloatList", None),
                           ("int64_list", 3, T.TYPE_MESSAGE, T.LABEL_OPTIONAL, "Int64List", None)])
    feat.oneof_decl.add(name="kind")
    for f in feat.field:
        f.oneof_index = 0
    feats = msg("Features", [("feature", 1, T.TYPE_MESSAGE, T.LABEL_REPEATED, "Features.FeatureEntry", None)])
    entry = feats.nested_type.add(name="FeatureEntry")
    entry.options.map_entry = True
    entry.field.add(name="key", number=1, type=T.TYPE_STRING, label=T.LABEL_OPTIONAL)
    v = entry.field.add(name="value", number=2, type=T.TYPE_MESSAGE, label=T.LABEL_OPTIONAL)
    v.type_name = ".tftest.Feature"
    msg("Example", [("features", 1, T.TYPE_MESSAGE, T.LABEL_OPTIONAL, "Features", None)])
    pool = descriptor_pool.DescriptorPool()
    pool.Add(fd)
    get = getattr(message_factory, "GetMessageClass", None)
    if get is None:
        return message_factory.MessageFactory(pool).GetPrototype(pool.FindMessageTypeByName("tftest.Example"))
    return get(pool.FindMessageTypeByName("tftest.Example"))


def test_parser_against_the_protobuf_runtime():
    Example = _example_classes()
    rng = np.random.default_rng(0)
    audio = rng.normal(0, 0.3, size=700).astype(np.float32)
    ex = Example()
    ex.features.feature["audio"].float_list.value.extend(audio.tolist())
    ex.features.feature["pitch"].int64_list.value.append(60)
    ex.features.feature["qualities"].int64_list.value.extend([0, 1, 0, 0, 1, 0, 0, 0, 0, -3])
    ex.features.feature["note_str"].bytes_list.value.append(b"bass_synthetic_033-060-100")
    got = nsynth.parse_example(ex.SerializeToString())
    np.testing.assert_array_equal(got["audio"], audio)
    np.testing.assert_array_equal(got["pitch"], [60])
    np.testing.assert_array_equal(got["qualities"], [0, 1, 0, 0, 1, 0, 0, 0, 0, -3])
    assert got["note_str"] == [b"bass_synthetic_033-060-100"]
    # and the writer: what serialize_example emits parses back with the protobuf runtime
    ex2 = Example()
    ex2.ParseFromString(nsynth.serialize_example({"audio": audio, "pitch": np.array([61]), "note_str": b"x"}))
    np.testing.assert_array_equal(np.asarray(ex2.features.feature["audio"].float_list.value, dtype=np.float32), audio)
    assert list(ex2.features.feature["pitch"].int64_list.value) == [61]
    assert list(ex2.features.feature["note_str"].bytes_list.value) == [b"x"]


def _write(tmp_path, n, length=512):
    rng = np.random.default_rng(1)
    clips = [rng.uniform(-1, 1, size=length).astype(np.float32) for _ in range(n)]
    path = os.path.join(str(tmp_path), "clips.tfrecord")
    nsynth.write_tfrecord(path, [{"audio": c, "pitch": np.array([40 + i])} for i, c in enumerate(clips)])
    return path, clips


def test_reader_contract(tmp_path):
    path, clips = _write(tmp_path, 5)
    r = nsynth.NsynthDataReader(path, 2, num_samples=300, shuffle=False, repeat=True, audio_max_length=512)
    x, y = r.next()
    assert x.shape == (2, 300) and x.dtype == np.float32 and y.shape == (2, 128)
    np.testing.assert_array_equal(x, np.stack([clips[0][:300], clips[1][:300]]))
    assert y[0].argmax() == 40 and y[1].argmax() == 41 and y.sum() == 2
    r.next()
    x, y = r.next()                                              # batches span the end of a pass when repeating
    np.testing.assert_array_equal(x, np.stack([clips[4][:300], clips[0][:300]]))
    # no repeat: a short final batch, then the end of the data
    r = nsynth.NsynthDataReader(path, 2, num_samples=300, shuffle=False, repeat=False, audio_max_length=512)
    sizes = [r.next()[0].shape[0] for _ in range(3)]
    assert sizes == [2, 2, 1]
    with pytest.raises(StopIteration):
        r.next()
    # shuffle: every pass is a permutation of the file
    r = nsynth.NsynthDataReader(path, 5, num_samples=4, shuffle=True, repeat=True, audio_max_length=512, seed=3)
    for _ in range(3):
        _, y = r.next()
        assert sorted(y.argmax(1).tolist()) == [40, 41, 42, 43, 44]
    # FixedLenFeature([audio_max_length]) rejects other lengths
    with pytest.raises(ValueError):
        nsynth.NsynthDataReader(path, 2, audio_max_length=64000).next()
    # unreduced records come back whole
    r = nsynth.NsynthDataReader(path, 1, reduced=False, shuffle=False, audio_max_length=512)
    assert set(r.next()[0]) == {"audio", "pitch"}


def test_corrupt_records_are_detected(tmp_path):
    path, _ = _write(tmp_path, 1, length=64)
    raw = bytearray(open(path, "rb").read())
    n, = struct.unpack("<Q", raw[:8])
    assert len(raw) == 16 + n
    assert len(list(nsynth.read_records(path, verify_payload_crc=True))) == 1
    bad = bytearray(raw); bad[20] ^= 0xFF
    open(path, "wb").write(bad)
    with pytest.raises(IOError):
        list(nsynth.read_records(path, verify_payload_crc=True))
    bad = bytearray(raw); bad[0] ^= 0x01
    open(path, "wb").write(bad)
    with pytest.raises(IOError):
        list(nsynth.read_records(path))
    open(path, "wb").write(raw[:-3])
    with pytest.raises(IOError):
        list(nsynth.read_records(path))
