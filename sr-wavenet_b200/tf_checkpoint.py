"""TensorFlow checkpoint (tensor bundle, "V2") reader / writer without TensorFlow -- SURVEY.md 8(f)-3.

The reference saves and restores with ``tf.train.Saver`` (model.py:119, 217-239, 412, 540-567): a checkpoint
``<logdir>/model.ckpt-<step>`` is the pair ``model.ckpt-<step>.index`` + ``model.ckpt-<step>.data-00000-of-00001`` and
``<logdir>/checkpoint`` names the latest one.  This module reads that pair into ``{variable name: ndarray}`` (the names
are the TF variable names the rest of the package already uses) and writes one, so trained reference teachers /
students can be loaded and this build's weights handed back.

Format, from the published sources (tensorflow/core/util/tensor_bundle, tensorflow/core/lib/io/table = LevelDB's table):

* ``.index`` is a LevelDB table.  Footer = last 48 bytes: metaindex BlockHandle, index BlockHandle (each two varint64:
  offset, size), zero padding to 40 bytes, magic 0xdb4775248b80fb57 (little endian).  A block is a run of entries
  ``varint32 shared | varint32 non_shared | varint32 value_len | key suffix | value`` followed by the restart array
  (uint32 offsets) and its length (uint32); on disk every block is followed by a 5-byte trailer (compression type,
  masked CRC-32C of block + type).  The index block maps separator keys to the BlockHandles of the data blocks.
* keys are variable names in byte order; the empty key holds a ``BundleHeaderProto`` (num_shards = 1, endianness = 2,
  version = 3), every other value is a ``BundleEntryProto`` (dtype = 1, shape = 2, shard_id = 3, offset = 4, size = 5,
  crc32c = 6 fixed32 -- the masked CRC-32C of the tensor bytes).
* ``.data-<shard>-of-<shards>`` holds the raw little-endian tensor bytes at those offsets.

PARITY UNPINNED: no TensorFlow and no TensorFlow-written checkpoint exist in this environment, so the reader is checked
against this module's own writer, against hand-assembled blocks (prefix-compressed keys, several data blocks) and
against the CRCs the format carries -- not against a file produced by TF itself.  Snappy-compressed blocks (type 1) are
reported as unsupported (BundleWriter writes its index uncompressed).
"""
import os
import struct

import numpy as np

from .nsynth import _enc_varint, _fields, _ld, _varint, masked_crc32c

MAGIC = 0xdb4775248b80fb57
# DataType enum values of tensorflow/core/framework/types.proto that a Saver writes for this model family
DTYPES = {1: np.dtype("<f4"), 2: np.dtype("<f8"), 3: np.dtype("<i4"), 9: np.dtype("<i8"), 10: np.dtype(np.bool_)}
DTYPE_IDS = {v: k for k, v in DTYPES.items()}


# ---- LevelDB table ---------------------------------------------------------------------------------------------------
def _block_entries(block):
    """[(key, value)] of one block (restart array ignored: entries are decoded front to back)."""
    n_restarts, = struct.unpack("<I", block[-4:])
    end = len(block) - 4 - 4 * n_restarts
    pos, key, out = 0, b"", []
    while pos < end:
        shared, pos = _varint(block, pos)
        non_shared, pos = _varint(block, pos)
        vlen, pos = _varint(block, pos)
        key = key[:shared] + bytes(block[pos:pos + non_shared])
        pos += non_shared
        out.append((key, bytes(block[pos:pos + vlen])))
        pos += vlen
    return out


def _read_block(buf, offset, size, verify=True):
    block, trailer = buf[offset:offset + size], buf[offset + size:offset + size + 5]
    if len(block) < size or len(trailer) < 5:
        raise IOError("truncated table block")
    if verify and struct.unpack("<I", trailer[1:])[0] != masked_crc32c(block + trailer[:1]):
        raise IOError("corrupt table block (CRC mismatch)")
    if trailer[0] != 0:
        raise NotImplementedError("compressed table block (type %d): only uncompressed index files are supported" % trailer[0])
    return block


def _handle(buf, pos):
    off, pos = _varint(buf, pos)
    size, pos = _varint(buf, pos)
    return off, size, pos


def read_table(path, verify=True):
    """All (key, value) pairs of a LevelDB table file, in key order."""
    buf = open(path, "rb").read()
    if len(buf) < 48 or struct.unpack("<Q", buf[-8:])[0] != MAGIC:
        raise IOError("%s is not a LevelDB table (bad magic)" % path)
    footer = buf[-48:]
    _, _, pos = _handle(footer, 0)                      # metaindex (no filter blocks in a bundle index)
    ioff, isize, _ = _handle(footer, pos)
    out = []
    for _, handle in _block_entries(_read_block(buf, ioff, isize, verify)):
        off, size, _ = _handle(handle, 0)
        out.extend(_block_entries(_read_block(buf, off, size, verify)))
    return out


def _build_block(pairs, restart_interval=16):
    body, restarts, prev = bytearray(), [], b""
    for i, (k, v) in enumerate(pairs):
        shared = 0
        if i % restart_interval == 0:
            restarts.append(len(body))
        else:
            while shared < min(len(prev), len(k)) and prev[shared] == k[shared]:
                shared += 1
        body += _enc_varint(shared) + _enc_varint(len(k) - shared) + _enc_varint(len(v)) + k[shared:] + v
        prev = k
    if not restarts:
        restarts = [0]
    return bytes(body) + b"".join(struct.pack("<I", r) for r in restarts) + struct.pack("<I", len(restarts))


def write_table(path, pairs, block_size=4096):
    """Writes sorted (key, value) pairs as an uncompressed LevelDB table."""
    pairs = sorted(pairs)
    out, index = bytearray(), []

    def emit(block):
        off = len(out)
        out.extend(block + b"\x00" + struct.pack("<I", masked_crc32c(block + b"\x00")))
        return off, len(block)

    cur, cur_bytes = [], 0
    for k, v in pairs:
        cur.append((k, v))
        cur_bytes += len(k) + len(v) + 3
        if cur_bytes >= block_size:
            off, size = emit(_build_block(cur))
            index.append((cur[-1][0], _enc_varint(off) + _enc_varint(size)))
            cur, cur_bytes = [], 0
    if cur or not index:
        off, size = emit(_build_block(cur))
        index.append((cur[-1][0] if cur else b"", _enc_varint(off) + _enc_varint(size)))
    moff, msize = emit(_build_block([]))
    ioff, isize = emit(_build_block(index, restart_interval=1))
    footer = _enc_varint(moff) + _enc_varint(msize) + _enc_varint(ioff) + _enc_varint(isize)
    out.extend(footer + b"\x00" * (40 - len(footer)) + struct.pack("<Q", MAGIC))
    with open(path, "wb") as f:
        f.write(bytes(out))


# ---- bundle protos ---------------------------------------------------------------------------------------------------
def _parse_entry(buf):
    e = {"dtype": 0, "shape": [], "shard_id": 0, "offset": 0, "size": 0, "crc32c": None, "slices": 0}
    for num, wt, val in _fields(memoryview(buf)):
        if num == 1 and wt == 0:
            e["dtype"] = val
        elif num == 2 and wt == 2:                       # TensorShapeProto: repeated Dim dim = 2 { int64 size = 1 }
            for n2, w2, dim in _fields(val):
                if n2 == 2 and w2 == 2:
                    size = 0
                    for n3, w3, v3 in _fields(dim):
                        if n3 == 1 and w3 == 0:
                            size = v3
                    e["shape"].append(size)
        elif num == 3 and wt == 0:
            e["shard_id"] = val
        elif num == 4 and wt == 0:
            e["offset"] = val
        elif num == 5 and wt == 0:
            e["size"] = val
        elif num == 6 and wt == 5:
            e["crc32c"], = struct.unpack("<I", val)
        elif num == 7:
            e["slices"] += 1
    return e


def _enc_entry(dtype_id, shape, offset, size, crc):
    dims = b"".join(_ld(2, _enc_varint((1 << 3) | 0) + _enc_varint(int(d))) for d in shape)
    out = _enc_varint((1 << 3) | 0) + _enc_varint(dtype_id) + _ld(2, dims)
    if offset:
        out += _enc_varint((4 << 3) | 0) + _enc_varint(offset)
    out += _enc_varint((5 << 3) | 0) + _enc_varint(size)
    out += _enc_varint((6 << 3) | 5) + struct.pack("<I", crc)
    return out


def _header_shards(buf):
    shards = 1
    for num, wt, val in _fields(memoryview(buf)):
        if num == 1 and wt == 0:
            shards = val
        elif num == 2 and wt == 0 and val != 0:
            raise NotImplementedError("big-endian tensor bundle")
    return shards


def list_variables(prefix):
    """[(name, shape, dtype)] of the checkpoint ``prefix`` (the path without .index / .data-...)."""
    out = []
    for k, v in read_table(prefix + ".index"):
        if k:
            e = _parse_entry(v)
            out.append((k.decode("utf-8"), tuple(e["shape"]), DTYPES.get(e["dtype"])))
    return out


def read_checkpoint(prefix, verify=True):
    """{variable name: ndarray} of the checkpoint ``prefix``.  ``verify`` checks the per-tensor CRC-32C."""
    pairs = read_table(prefix + ".index", verify)
    shards = 1
    for k, v in pairs:
        if not k:
            shards = _header_shards(v)
    files, out = {}, {}
    for k, v in pairs:
        if not k:
            continue
        e = _parse_entry(v)
        name = k.decode("utf-8")
        if e["slices"]:
            raise NotImplementedError("partitioned variable %s (tensor slices)" % name)
        if e["dtype"] not in DTYPES:
            continue                                       # string / resource tensors: nothing the model needs
        sid = e["shard_id"]
        if sid not in files:
            files[sid] = open("%s.data-%05d-of-%05d" % (prefix, sid, shards), "rb")
        f = files[sid]
        f.seek(e["offset"])
        raw = f.read(e["size"])
        if len(raw) != e["size"]:
            raise IOError("truncated data shard for %s" % name)
        if verify and e["crc32c"] is not None and masked_crc32c(raw) != e["crc32c"]:
            raise IOError("CRC mismatch for %s" % name)
        out[name] = np.frombuffer(raw, dtype=DTYPES[e["dtype"]]).reshape(e["shape"]).copy()
    for f in files.values():
        f.close()
    return out


def write_checkpoint(prefix, tensors):
    """Writes ``{name: ndarray}`` as a single-shard tensor bundle ``prefix.index`` + ``prefix.data-00000-of-00001``."""
    pairs = [(b"", _enc_varint((1 << 3) | 0) + _enc_varint(1) + _ld(3, _enc_varint((1 << 3) | 0) + _enc_varint(1)))]
    offset = 0
    with open(prefix + ".data-00000-of-00001", "wb") as f:
        for name in sorted(tensors, key=lambda s: s.encode("utf-8")):
            a = np.asarray(tensors[name])                    # (ascontiguousarray would turn a scalar into shape (1,))
            a = a if a.flags.c_contiguous else np.ascontiguousarray(a)
            if a.dtype not in DTYPE_IDS:
                a = a.astype(np.float32)
            raw = a.astype(a.dtype.newbyteorder("<")).tobytes()
            f.write(raw)
            pairs.append((name.encode("utf-8"), _enc_entry(DTYPE_IDS[a.dtype], a.shape, offset, len(raw), masked_crc32c(raw))))
            offset += len(raw)
    write_table(prefix + ".index", pairs)


def latest_checkpoint(logdir):
    """tf.train.latest_checkpoint: the prefix named by ``<logdir>/checkpoint``, or None."""
    state = os.path.join(logdir, "checkpoint")
    if not os.path.exists(state):
        return None
    with open(state) as f:
        for line in f:
            if line.startswith("model_checkpoint_path:"):
                name = line.split('"')[1]
                return name if os.path.isabs(name) else os.path.join(logdir, name)
    return None
