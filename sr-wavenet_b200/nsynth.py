"""NSynth TFRecord reader with the contract of the reference's nsynth.py (nsynth.py:5-52), without TensorFlow.

The reference builds ``tf.data.TFRecordDataset(filepath).map(parse).shuffle(10000).repeat().batch(batch_size)`` and
serves batches from a private tf.Session; ``next()`` returns ``(audio[:, :num_samples], one_hot(pitch, 128))`` when
``reduced`` (the only mode the drivers use, teacher.py:53 / student.py:55).  TensorFlow is not part of this build, so the
two file formats involved are read directly:

* TFRecord framing: ``uint64 length | uint32 masked_crc32c(length) | data | uint32 masked_crc32c(data)``, little endian,
  ``masked = ((crc >> 15) | (crc << 17)) + 0xa282ead8``;
* ``tf.train.Example`` protobuf: ``Example{features=1}``, ``Features{map<string, Feature> feature=1}``,
  ``Feature{bytes_list=1 | float_list=2 | int64_list=3}``, each list ``{repeated value=1}`` (floats / ints packed or not).

``write_tfrecord`` is the inverse (create_tfrecord.py:5-58 writes ``pitch`` + ``audio`` only; filter_tfrecord.py keeps the
same two features), used by the tests and to prepare data without TensorFlow.  Pure host I/O, nothing here touches the GPU.
"""
import struct

import numpy as np

# ---- crc32c (Castagnoli), table driven ----------------------------------------------------------------------------
_CRC_TABLE = None


def _crc_table():
    global _CRC_TABLE
    if _CRC_TABLE is None:
        t = np.zeros(256, dtype=np.uint32)
        for i in range(256):
            c = i
            for _ in range(8):
                c = (c >> 1) ^ 0x82F63B78 if c & 1 else c >> 1
            t[i] = c
        _CRC_TABLE = [int(v) for v in t]
    return _CRC_TABLE


def crc32c(data):
    t = _crc_table()
    c = 0xFFFFFFFF
    for b in bytes(data):
        c = t[(c ^ b) & 0xFF] ^ (c >> 8)
    return c ^ 0xFFFFFFFF


def masked_crc32c(data):
    c = crc32c(data)
    return ((((c >> 15) | (c << 17)) & 0xFFFFFFFF) + 0xA282EAD8) & 0xFFFFFFFF


# ---- TFRecord framing ------------------------------------------------------------------------------------------------
def read_records(path, verify_payload_crc=False):
    """Yields the payload of every record.  The length CRC is always checked (it is 12 bytes); the payload CRC costs a
    Python loop over 256 KB of audio per record, so it is opt-in."""
    with open(path, "rb") as f:
        while True:
            head = f.read(12)
            if not head:
                return
            if len(head) < 12:
                raise IOError("truncated TFRecord header in %s" % path)
            n, = struct.unpack("<Q", head[:8])
            if struct.unpack("<I", head[8:])[0] != masked_crc32c(head[:8]):
                raise IOError("corrupt TFRecord length in %s" % path)
            data = f.read(n)
            tail = f.read(4)
            if len(data) < n or len(tail) < 4:
                raise IOError("truncated TFRecord in %s" % path)
            if verify_payload_crc and struct.unpack("<I", tail)[0] != masked_crc32c(data):
                raise IOError("corrupt TFRecord payload in %s" % path)
            yield data


def write_records(path, payloads):
    with open(path, "wb") as f:
        for data in payloads:
            head = struct.pack("<Q", len(data))
            f.write(head + struct.pack("<I", masked_crc32c(head)) + data + struct.pack("<I", masked_crc32c(data)))


# ---- protobuf wire format (the subset tf.train.Example uses) ---------------------------------------------------------
def _varint(buf, pos):
    out, shift = 0, 0
    while True:
        b = buf[pos]
        pos += 1
        out |= (b & 0x7F) << shift
        if not b & 0x80:
            return out, pos
        shift += 7


def _fields(buf):
    """Yields (field number, wire type, value) of one message; value is an int (varint, fixed) or a memoryview (bytes)."""
    pos, n = 0, len(buf)
    while pos < n:
        key, pos = _varint(buf, pos)
        num, wt = key >> 3, key & 7
        if wt == 0:
            val, pos = _varint(buf, pos)
        elif wt == 1:
            val = bytes(buf[pos:pos + 8]); pos += 8
        elif wt == 2:
            ln, pos = _varint(buf, pos)
            val = buf[pos:pos + ln]; pos += ln
        elif wt == 5:
            val = bytes(buf[pos:pos + 4]); pos += 4
        else:
            raise ValueError("unsupported protobuf wire type %d" % wt)
        yield num, wt, val


def _parse_feature(buf):
    for num, wt, val in _fields(buf):
        if wt != 2:
            continue
        if num == 1:                                           # BytesList
            return [bytes(v) for n, w, v in _fields(val) if n == 1 and w == 2]
        if num == 2:                                           # FloatList: packed (one bytes field) or one fixed32 per value
            parts = []
            for n, w, v in _fields(val):
                if n == 1 and w == 2:
                    parts.append(np.frombuffer(bytes(v), dtype="<f4"))
                elif n == 1 and w == 5:
                    parts.append(np.frombuffer(v, dtype="<f4"))
            return np.concatenate(parts) if parts else np.zeros(0, dtype=np.float32)
        if num == 3:                                           # Int64List: packed varints or one varint per value
            out = []
            for n, w, v in _fields(val):
                if n == 1 and w == 2:
                    p = 0
                    while p < len(v):
                        x, p = _varint(v, p)
                        out.append(x - (1 << 64) if x >= 1 << 63 else x)
                elif n == 1 and w == 0:
                    out.append(v - (1 << 64) if v >= 1 << 63 else v)
            return np.asarray(out, dtype=np.int64)
    return None


def parse_example(data):
    """Serialized tf.train.Example -> {feature name: float32 array | int64 array | list of bytes}."""
    out = {}
    buf = memoryview(data)
    for num, wt, features in _fields(buf):
        if num != 1 or wt != 2:
            continue
        for n2, w2, entry in _fields(features):                # map<string, Feature> entries
            if n2 != 1 or w2 != 2:
                continue
            key, feat = None, None
            for n3, w3, v in _fields(entry):
                if n3 == 1 and w3 == 2:
                    key = bytes(v).decode("utf-8")
                elif n3 == 2 and w3 == 2:
                    feat = _parse_feature(v)
            if key is not None:
                out[key] = feat
    return out


def _enc_varint(x):
    x &= (1 << 64) - 1
    out = bytearray()
    while True:
        b = x & 0x7F
        x >>= 7
        if x:
            out.append(b | 0x80)
        else:
            out.append(b)
            return bytes(out)


def _ld(num, payload):                                         # length-delimited field
    return _enc_varint((num << 3) | 2) + _enc_varint(len(payload)) + payload


def serialize_example(features):
    """{name: float array | int array | bytes | list of bytes} -> serialized tf.train.Example (packed lists, as
    tf.train.Example.SerializeToString writes them)."""
    entries = b""
    for key in sorted(features):
        v = features[key]
        if isinstance(v, (bytes, bytearray)):
            v = [bytes(v)]
        if isinstance(v, list) and (not v or isinstance(v[0], (bytes, bytearray))):
            feat = _ld(1, b"".join(_ld(1, bytes(b)) for b in v))
        else:
            a = np.asarray(v)
            if a.dtype.kind == "f":
                feat = _ld(2, _ld(1, a.astype("<f4").tobytes()))
            else:
                feat = _ld(3, _ld(1, b"".join(_enc_varint(int(x)) for x in a.reshape(-1))))
        entries += _ld(1, _ld(1, key.encode("utf-8")) + _ld(2, feat))
    return _ld(1, entries)


def write_tfrecord(path, examples):
    """examples: iterable of feature dicts (see serialize_example)."""
    write_records(path, (serialize_example(e) for e in examples))


# ---- the reader class of nsynth.py ----------------------------------------------------------------------------------
class NsynthDataReader(object):
    """nsynth.py:5-52.  ``next()`` -> ``(audio [B, num_samples] float32, pitch one-hot [B, 128] float32)`` when
    ``reduced``; otherwise a list of B feature dicts.  ``shuffle`` keeps tf.data's semantics (a 10000-element buffer, one
    uniformly random element leaves per draw); ``repeat`` restarts the file for ever; without it the final batch may be
    short and the next call raises StopIteration (tf.errors.OutOfRangeError in the reference).  ``seed`` is an addition
    (the reference is unseeded)."""

    SHUFFLE_BUFFER = 10000

    def __init__(self, filepath, batch_size, num_samples=16000, reduced=True, shuffle=True, repeat=True,
                 audio_max_length=64000, seed=None):
        self.filepath, self.batch_size, self.num_samples = filepath, batch_size, num_samples
        self.reduced, self.shuffle, self.repeat, self.audio_max_length = reduced, shuffle, repeat, audio_max_length
        self._rng = np.random.default_rng(seed)
        self._it = self._batches()

    def _parsed(self):
        while True:
            n = 0
            for rec in read_records(self.filepath):
                ex = parse_example(rec)
                audio = ex.get("audio")
                if audio is None or audio.shape[0] != self.audio_max_length:     # tf.FixedLenFeature([audio_max_length])
                    raise ValueError("feature 'audio' has %s values, expected %d" % (
                        None if audio is None else audio.shape[0], self.audio_max_length))
                if ex.get("pitch") is None or ex["pitch"].shape[0] != 1:
                    raise ValueError("feature 'pitch' must hold one int64")
                n += 1
                yield ex
            if not self.repeat or n == 0:
                return

    def _shuffled(self, src):
        buf = []
        for ex in src:
            buf.append(ex)
            if len(buf) > self.SHUFFLE_BUFFER:
                i = int(self._rng.integers(len(buf)))
                buf[i], buf[-1] = buf[-1], buf[i]
                yield buf.pop()
        while buf:
            i = int(self._rng.integers(len(buf)))
            buf[i], buf[-1] = buf[-1], buf[i]
            yield buf.pop()

    def _batches(self):
        # map -> shuffle(10000) -> repeat -> batch: every pass over the file is shuffled on its own (the buffer drains when
        # the upstream ends), and batches may span the boundary between two passes
        src = self._shuffled_passes() if self.shuffle else self._parsed()
        batch = []
        for ex in src:
            batch.append(ex)
            if len(batch) == self.batch_size:
                yield self._emit(batch)
                batch = []
        if batch:
            yield self._emit(batch)

    def _shuffled_passes(self):
        while True:
            one_pass = NsynthDataReader.__new__(NsynthDataReader)
            one_pass.__dict__.update(self.__dict__)
            one_pass.repeat = False
            n = 0
            for ex in self._shuffled(one_pass._parsed()):
                n += 1
                yield ex
            if not self.repeat or n == 0:
                return

    def _emit(self, batch):
        if not self.reduced:
            return batch
        audio = np.stack([ex["audio"][:self.num_samples] for ex in batch]).astype(np.float32)
        pitch = np.zeros((len(batch), 128), dtype=np.float32)
        for i, ex in enumerate(batch):
            p = int(ex["pitch"][0])
            if 0 <= p < 128:                                   # tf.one_hot: out-of-range indices give an all-zero row
                pitch[i, p] = 1.0
        return audio, pitch

    def next(self):
        return next(self._it)
