"""TensorFlow V2 checkpoint ("tensor bundle") reader / writer without TensorFlow.

The reference saves and restores with `tf.train.Saver` (base_model.py:219,223-243): files
`<dir>/VSR-<step>.index`, `<dir>/VSR-<step>.data-00000-of-00001` and the text file `<dir>/checkpoint`
(`model_checkpoint_path: "VSR-<step>"`, read by `tf.train.get_checkpoint_state`, base_model.py:235).
TensorFlow 1.12 is not installable here, so the format is restated from its public definition:

* `.index` is an SSTable in the LevelDB table format (tensorflow/core/lib/io/table*, identical to
  leveldb/doc/table_format.md): data blocks of prefix-compressed key/value entries with a restart
  array, a 1-byte compression tag + masked CRC-32C trailer per block, a metaindex block, an index
  block and a 48-byte footer ending in the magic 0xdb4775248b80fb57.
* key ""        -> BundleHeaderProto  {1: num_shards, 2: endianness, 3: VersionDef{1: producer}}
* key <varname> -> BundleEntryProto   {1: dtype, 2: TensorShapeProto{2: dim{1: size}}, 3: shard_id,
                                       4: offset, 5: size, 6: fixed32 masked crc32c of the bytes}
  (tensorflow/core/protobuf/tensor_bundle.proto)
* `.data-SSSSS-of-NNNNN` holds the raw little-endian tensor bytes at [offset, offset+size).

PARITY UNPINNED: no checkpoint ships with the reference (checkpoint/README.md:3) and no TF-written
file is available offline; the reader is exercised against files produced by the writer below
(round trip), against hand-assembled blocks, and against the CRC-32C known-answer vectors.
Snappy-compressed blocks (TF's BundleWriter does not produce them) are rejected with an error.
"""
from __future__ import annotations

import os
import re
import struct

import numpy as np

TABLE_MAGIC = 0xDB4775248B80FB57
_MASK_DELTA = 0xA282EAD8

# tensorflow/core/framework/types.proto
_DT = {1: np.float32, 2: np.float64, 3: np.int32, 4: np.uint8, 5: np.int16, 6: np.int8, 9: np.int64,
       10: np.bool_, 17: np.uint16, 19: np.float16, 22: np.uint32, 23: np.uint64}
_DT_REV = {np.dtype(v): k for k, v in _DT.items()}


# ---- CRC-32C (Castagnoli), table driven -------------------------------------------------------
def _make_table():
    tab = np.zeros(256, np.uint32)
    for i in range(256):
        c = i
        for _ in range(8):
            c = (c >> 1) ^ (0x82F63B78 if c & 1 else 0)
        tab[i] = c
    return tab


_CRC_TAB = _make_table()
_CRC_LIST = [int(v) for v in _CRC_TAB]


def _native_crc():
    """pfnl_crc32c of libpfnl_b200.so (host code, ~1 GB/s) when the library is built; the pure-Python
    loop below computes the same function and is what the CPU tests compare it with."""
    try:
        from ._lib import lib
        return lib.pfnl_crc32c
    except Exception:  # library not built: stay in Python
        return None


_NATIVE = None


def crc32c(data: bytes, crc: int = 0) -> int:
    """CRC-32C of `data` (reflected polynomial 0x82F63B78, init/xorout 0xFFFFFFFF)."""
    global _NATIVE
    if len(data) >= 4096:
        if _NATIVE is None:
            _NATIVE = _native_crc() or False
        if _NATIVE:
            return int(_NATIVE(bytes(data), len(data), crc))
    return crc32c_py(data, crc)


def crc32c_py(data: bytes, crc: int = 0) -> int:
    c = crc ^ 0xFFFFFFFF
    tab = _CRC_LIST
    for b in data:
        c = tab[(c ^ b) & 0xFF] ^ (c >> 8)
    return c ^ 0xFFFFFFFF


def mask_crc(crc: int) -> int:
    """leveldb/TF crc32c::Mask: rotate right by 15 bits and add a constant."""
    return (((crc >> 15) | (crc << 17)) + _MASK_DELTA) & 0xFFFFFFFF


def unmask_crc(m: int) -> int:
    rot = (m - _MASK_DELTA) & 0xFFFFFFFF
    return ((rot >> 17) | (rot << 15)) & 0xFFFFFFFF


# ---- varints / minimal protobuf ---------------------------------------------------------------
def _get_varint(buf, pos):
    shift = 0
    val = 0
    while True:
        b = buf[pos]
        pos += 1
        val |= (b & 0x7F) << shift
        if not b & 0x80:
            return val, pos
        shift += 7
        if shift > 70:
            raise ValueError("malformed varint")


def _put_varint(v):
    out = bytearray()
    v &= (1 << 64) - 1
    while True:
        b = v & 0x7F
        v >>= 7
        if v:
            out.append(b | 0x80)
        else:
            out.append(b)
            return bytes(out)


def _pb_fields(buf):
    """Yield (field_number, wire_type, value) of one serialized message."""
    pos = 0
    n = len(buf)
    while pos < n:
        key, pos = _get_varint(buf, pos)
        f, wt = key >> 3, key & 7
        if wt == 0:
            v, pos = _get_varint(buf, pos)
        elif wt == 1:
            v = struct.unpack_from("<Q", buf, pos)[0]
            pos += 8
        elif wt == 2:
            ln, pos = _get_varint(buf, pos)
            v = bytes(buf[pos:pos + ln])
            pos += ln
        elif wt == 5:
            v = struct.unpack_from("<I", buf, pos)[0]
            pos += 4
        else:
            raise ValueError(f"unsupported protobuf wire type {wt}")
        yield f, wt, v


def _pb_varint(field, v):
    return _put_varint(field << 3) + _put_varint(v)


def _pb_bytes(field, b):
    return _put_varint((field << 3) | 2) + _put_varint(len(b)) + b


def _pb_fixed32(field, v):
    return _put_varint((field << 3) | 5) + struct.pack("<I", v)


def _parse_shape(buf):
    dims = []
    for f, _, v in _pb_fields(buf):
        if f == 2:  # repeated Dim
            size = 0
            for f2, _, v2 in _pb_fields(v):
                if f2 == 1:
                    size = v2 if v2 < (1 << 63) else v2 - (1 << 64)
            dims.append(size)
    return tuple(dims)


def _parse_entry(buf):
    e = {"dtype": 0, "shape": (), "shard_id": 0, "offset": 0, "size": 0, "crc32c": None, "slices": 0}
    for f, _, v in _pb_fields(buf):
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
            e["slices"] += 1
    return e


def _parse_header(buf):
    h = {"num_shards": 0, "endianness": 0, "producer": 0}
    for f, _, v in _pb_fields(buf):
        if f == 1:
            h["num_shards"] = v
        elif f == 2:
            h["endianness"] = v
        elif f == 3:
            for f2, _, v2 in _pb_fields(v):
                if f2 == 1:
                    h["producer"] = v2
    return h


# ---- LevelDB table ----------------------------------------------------------------------------
def _read_block(buf, offset, size, verify=True):
    """Block contents at a BlockHandle; checks the compression tag and the masked CRC trailer."""
    if offset + size + 5 > len(buf):
        raise ValueError("table block handle points outside the file")
    contents = bytes(buf[offset:offset + size])
    ctype = buf[offset + size]
    stored = struct.unpack_from("<I", buf, offset + size + 1)[0]
    if verify and unmask_crc(stored) != crc32c(contents + bytes([ctype])):
        raise ValueError("table block checksum mismatch (corrupt .index file)")
    if ctype == 1:
        raise ValueError("snappy-compressed table block: not supported (TF's BundleWriter writes uncompressed)")
    if ctype != 0:
        raise ValueError(f"unknown table block compression tag {ctype}")
    return contents


def _block_entries(block):
    """(key, value) pairs of one table block (prefix-compressed keys, restart array at the end)."""
    if len(block) < 4:
        raise ValueError("table block too small")
    num_restarts = struct.unpack_from("<I", block, len(block) - 4)[0]
    end = len(block) - 4 - 4 * num_restarts
    if end < 0:
        raise ValueError("bad restart array")
    pos = 0
    key = b""
    while pos < end:
        shared, pos = _get_varint(block, pos)
        non_shared, pos = _get_varint(block, pos)
        vlen, pos = _get_varint(block, pos)
        if shared > len(key):
            raise ValueError("bad key prefix length")
        key = key[:shared] + bytes(block[pos:pos + non_shared])
        pos += non_shared
        yield key, bytes(block[pos:pos + vlen])
        pos += vlen


def read_table(path, verify=True):
    """All (key, value) pairs of an SSTable file, in key order."""
    with open(path, "rb") as f:
        buf = f.read()
    if len(buf) < 48:
        raise ValueError(f"{path}: too small to be a table file")
    footer = buf[-48:]
    if struct.unpack("<Q", footer[40:])[0] != TABLE_MAGIC:
        raise ValueError(f"{path}: bad table magic (not a TF V2 checkpoint .index file)")
    pos = 0
    _, pos = _get_varint(footer, pos)          # metaindex handle (unused)
    _, pos = _get_varint(footer, pos)
    ioff, pos = _get_varint(footer, pos)
    isize, pos = _get_varint(footer, pos)
    out = []
    for _, handle in _block_entries(_read_block(buf, ioff, isize, verify)):
        boff, p = _get_varint(handle, 0)
        bsize, p = _get_varint(handle, p)
        out.extend(_block_entries(_read_block(buf, boff, bsize, verify)))
    return out


class _BlockBuilder:
    def __init__(self, restart_interval=16):
        self.buf = bytearray()
        self.restarts = [0]
        self.counter = 0
        self.last_key = b""
        self.interval = restart_interval

    def add(self, key, value):
        shared = 0
        if self.counter < self.interval:
            m = min(len(key), len(self.last_key))
            while shared < m and key[shared] == self.last_key[shared]:
                shared += 1
        else:
            self.restarts.append(len(self.buf))
            self.counter = 0
        self.buf += _put_varint(shared) + _put_varint(len(key) - shared) + _put_varint(len(value))
        self.buf += key[shared:] + value
        self.last_key = key
        self.counter += 1

    def finish(self):
        out = bytes(self.buf) + b"".join(struct.pack("<I", r) for r in self.restarts)
        return out + struct.pack("<I", len(self.restarts))

    def size(self):
        return len(self.buf) + 4 * len(self.restarts) + 4


def write_table(path, items, block_size=4096):
    """Write sorted (key, value) byte pairs as an uncompressed SSTable."""
    items = sorted(items)
    out = bytearray()

    def emit(contents):
        off = len(out)
        out.extend(contents)
        out.append(0)  # kNoCompression
        out.extend(struct.pack("<I", mask_crc(crc32c(contents + b"\x00"))))
        return _put_varint(off) + _put_varint(len(contents))

    index = _BlockBuilder(restart_interval=1)
    bb = _BlockBuilder()
    last = None
    for k, v in items:
        bb.add(k, v)
        last = k
        if bb.size() >= block_size:
            index.add(last, emit(bb.finish()))
            bb = _BlockBuilder()
            last = None
    if last is not None or not items:
        index.add(last if last is not None else b"", emit(bb.finish()))
    meta_handle = emit(_BlockBuilder().finish())
    index_handle = emit(index.finish())
    footer = meta_handle + index_handle
    footer += b"\x00" * (40 - len(footer)) + struct.pack("<Q", TABLE_MAGIC)
    out.extend(footer)
    with open(path, "wb") as f:
        f.write(bytes(out))


# ---- bundle -----------------------------------------------------------------------------------
def _data_path(prefix, shard, num_shards):
    return f"{prefix}.data-{shard:05d}-of-{num_shards:05d}"


def list_variables(prefix):
    """{name: (numpy dtype, shape)} of a V2 checkpoint `prefix` (path without .index)."""
    out = {}
    for k, v in read_table(prefix + ".index"):
        if k == b"":
            continue
        e = _parse_entry(v)
        out[k.decode("utf-8")] = (_DT.get(e["dtype"]), e["shape"])
    return out


def read_bundle(prefix, names=None, verify_crc=True):
    """Read tensors of a V2 checkpoint -> {name: ndarray}.  `names`: iterable to restrict to."""
    header = None
    entries = {}
    for k, v in read_table(prefix + ".index"):
        if k == b"":
            header = _parse_header(v)
        else:
            entries[k.decode("utf-8")] = _parse_entry(v)
    if header is None:
        raise ValueError(f"{prefix}.index: missing bundle header entry")
    if header["endianness"] != 0:
        raise ValueError("big-endian tensor bundle: not supported")
    want = list(entries) if names is None else list(names)
    shards = {}
    out = {}
    try:
        for name in want:
            if name not in entries:
                raise KeyError(f"variable '{name}' not found in checkpoint {prefix}")
            e = entries[name]
            if e["slices"]:
                raise ValueError(f"variable '{name}' is stored as partitioned slices: not supported")
            dt = _DT.get(e["dtype"])
            if dt is None:
                raise ValueError(f"variable '{name}': unsupported dtype enum {e['dtype']}")
            sid = e["shard_id"]
            if sid not in shards:
                shards[sid] = open(_data_path(prefix, sid, header["num_shards"]), "rb")
            count = int(np.prod(e["shape"], dtype=np.int64)) if e["shape"] else 1
            if count * np.dtype(dt).itemsize != e["size"]:
                raise ValueError(f"variable '{name}': size {e['size']} does not match shape {e['shape']}")
            f = shards[sid]
            f.seek(e["offset"])
            b = f.read(e["size"])
            if len(b) != e["size"]:
                raise ValueError(f"variable '{name}': data file truncated")
            if verify_crc and e["crc32c"] is not None and unmask_crc(e["crc32c"]) != crc32c(b):
                raise ValueError(f"variable '{name}': data checksum mismatch")
            out[name] = np.frombuffer(b, dtype=dt).reshape(e["shape"]).copy()
    finally:
        for f in shards.values():
            f.close()
    return out


def write_bundle(prefix, tensors):
    """Write {name: ndarray} as a single-shard V2 checkpoint (<prefix>.index + .data-00000-of-00001)."""
    items = []
    header = _pb_varint(1, 1) + _pb_varint(2, 0) + _pb_bytes(3, _pb_varint(1, 1))
    items.append((b"", header))
    offset = 0
    with open(_data_path(prefix, 0, 1), "wb") as f:
        for name in sorted(tensors):
            a = np.asarray(tensors[name], order="C")  # (ascontiguousarray would turn a scalar into shape (1,))
            if a.dtype not in _DT_REV:
                raise ValueError(f"'{name}': dtype {a.dtype} has no TF DataType mapping here")
            b = a.astype(a.dtype.newbyteorder("<"), copy=False).tobytes()
            shape = b"".join(_pb_bytes(2, _pb_varint(1, int(d))) for d in a.shape)
            e = _pb_varint(1, _DT_REV[a.dtype]) + _pb_bytes(2, shape)
            if offset:
                e += _pb_varint(4, offset)
            e += _pb_varint(5, len(b)) + _pb_fixed32(6, mask_crc(crc32c(b)))
            items.append((name.encode("utf-8"), e))
            f.write(b)
            offset += len(b)
    write_table(prefix + ".index", items)


# ---- checkpoint state file ('checkpoint', CheckpointState text proto) ----------------------------
def read_checkpoint_state(checkpoint_dir):
    """-> model_checkpoint_path (as written, may be relative to the directory) or None
    (tf.train.get_checkpoint_state, base_model.py:235)."""
    p = os.path.join(checkpoint_dir, "checkpoint")
    if not os.path.isfile(p):
        return None
    with open(p, "rt") as f:
        m = re.search(r'^\s*model_checkpoint_path:\s*"(.*)"\s*$', f.read(), re.M)
    return m.group(1) if m else None


def write_checkpoint_state(checkpoint_dir, name, keep=()):
    with open(os.path.join(checkpoint_dir, "checkpoint"), "wt") as f:
        f.write(f'model_checkpoint_path: "{name}"\n')
        for k in list(keep) + [name]:
            f.write(f'all_model_checkpoint_paths: "{k}"\n')


def latest_checkpoint(checkpoint_dir):
    """Prefix of the checkpoint the reference's `load` would restore: the state file's entry, by
    basename inside `checkpoint_dir` (base_model.py:236-238), else None."""
    mp = read_checkpoint_state(checkpoint_dir)
    if not mp:
        return None
    prefix = os.path.join(checkpoint_dir, os.path.basename(mp))
    return prefix if os.path.isfile(prefix + ".index") else None
