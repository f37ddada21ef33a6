"""TensorFlow "bundle V2" checkpoint reader / writer (``<prefix>.index`` + ``<prefix>.data-00000-of-00001``).

The reference saves and restores its variables with ``tf.train.Saver`` (training.py:141, 203-207; demo.py:33;
embeddings.py:168) and ships ``model/air-model.{index,meta}`` (the data shard is missing from the copy at hand).
This module reads and writes that format without TensorFlow so that published weights can drive the CUDA model and
its checkpoints load back into the reference:

* ``.index`` is a LevelDB-style table: one (or more) data blocks of prefix-compressed (key, value) entries with
  restart points every 16 entries, an empty metaindex block, an index block, and a 48-byte footer ending in the
  magic ``0xdb4775248b80fb57``; every block is followed by a 1-byte compression type (0) and a masked CRC32C.
  Key ``""`` holds a ``BundleHeaderProto``; every other key is a variable name holding a ``BundleEntryProto``
  {dtype, shape, shard_id, offset, size, crc32c}.
* ``.data-00000-of-00001`` is the raw little-endian tensor bytes, concatenated in key order.

The ``.meta`` MetaGraphDef is a TensorFlow graph description and is neither needed nor written here.
"""
from __future__ import annotations

import struct
from collections import OrderedDict, namedtuple

import numpy as np

MAGIC = 0xDB4775248B80FB57
DT_FLOAT, DT_INT32 = 1, 3
_DTYPES = {DT_FLOAT: np.dtype("<f4"), DT_INT32: np.dtype("<i4")}
_RESTART_INTERVAL = 16

BundleEntry = namedtuple("BundleEntry", "dtype shape shard_id offset size crc32c")

# ---- CRC32C (Castagnoli), as used by LevelDB / TensorFlow ------------------------------------------
_CRC_TABLE = None


def _crc_table():
    global _CRC_TABLE
    if _CRC_TABLE is None:
        t = []
        for i in range(256):
            c = i
            for _ in range(8):
                c = (c >> 1) ^ 0x82F63B78 if c & 1 else c >> 1
            t.append(c)
        _CRC_TABLE = t
    return _CRC_TABLE


def crc32c(data: bytes, crc: int = 0) -> int:
    """CRC-32C; uses the SSE4.2 implementation in libair_b200.so when the library is built, else pure Python."""
    try:
        from . import _cabi
        return int(_cabi.lib().air_crc32c(bytes(data), len(data), crc))
    except Exception:  # noqa: BLE001  (library not built: checkpoint I/O still works, just slowly)
        pass
    t = _crc_table()
    c = crc ^ 0xFFFFFFFF
    for b in data:
        c = t[(c ^ b) & 0xFF] ^ (c >> 8)
    return c ^ 0xFFFFFFFF


def mask_crc(crc: int) -> int:
    return (((crc >> 15) | (crc << 17)) + 0xA282EAD8) & 0xFFFFFFFF


# ---- varints / protos ---------------------------------------------------------------------------------
def _get_varint(b, i):
    r = s = 0
    while True:
        c = b[i]
        i += 1
        r |= (c & 0x7F) << s
        s += 7
        if not c & 0x80:
            return r, i


def _put_varint(v: int) -> bytes:
    out = bytearray()
    while True:
        c = v & 0x7F
        v >>= 7
        if v:
            out.append(c | 0x80)
        else:
            out.append(c)
            return bytes(out)


def _decode_entry(v: bytes) -> BundleEntry:
    dtype = shard = offset = size = crc = 0
    shape = []
    i = 0
    while i < len(v):
        tag, i = _get_varint(v, i)
        field, wt = tag >> 3, tag & 7
        if wt == 0:
            val, i = _get_varint(v, i)
            if field == 1:
                dtype = val
            elif field == 3:
                shard = val
            elif field == 4:
                offset = val
            elif field == 5:
                size = val
        elif wt == 5:
            (val,) = struct.unpack_from("<I", v, i)
            i += 4
            if field == 6:
                crc = val
        elif wt == 2:
            ln, i = _get_varint(v, i)
            sub = v[i:i + ln]
            i += ln
            if field == 2:  # TensorShapeProto { repeated Dim dim = 2 { int64 size = 1 } }
                j = 0
                while j < len(sub):
                    t2, j = _get_varint(sub, j)
                    l2, j = _get_varint(sub, j)
                    dim = sub[j:j + l2]
                    j += l2
                    if t2 >> 3 == 2:
                        k = 0
                        sz = 0
                        while k < len(dim):
                            t3, k = _get_varint(dim, k)
                            val3, k = _get_varint(dim, k)
                            if t3 >> 3 == 1:
                                sz = val3
                        shape.append(sz)
        else:
            raise ValueError(f"unexpected wire type {wt} in BundleEntryProto")
    return BundleEntry(dtype, tuple(shape), shard, offset, size, crc)


def _encode_entry(e: BundleEntry) -> bytes:
    shape = b"".join(b"\x12" + _put_varint(len(d)) + d for d in (b"\x08" + _put_varint(s) for s in e.shape))
    out = b"\x08" + _put_varint(e.dtype) + b"\x12" + _put_varint(len(shape)) + shape
    if e.shard_id:
        out += b"\x18" + _put_varint(e.shard_id)
    if e.offset:
        out += b"\x20" + _put_varint(e.offset)
    out += b"\x28" + _put_varint(e.size) + b"\x35" + struct.pack("<I", e.crc32c)
    return out


_HEADER = b"\x08\x01\x1a\x02\x08\x01"  # BundleHeaderProto{num_shards: 1, version{producer: 1}}, little endian


# ---- table reading ---------------------------------------------------------------------------------------
def _read_block(buf: bytes, offset: int, size: int, verify: bool):
    raw = buf[offset:offset + size]
    if verify:
        ctype = buf[offset + size]
        (stored,) = struct.unpack_from("<I", buf, offset + size + 1)
        if ctype != 0:
            raise ValueError("compressed table blocks are not supported")
        if mask_crc(crc32c(raw + bytes([ctype]))) != stored:
            raise ValueError("table block checksum mismatch")
    (n_restarts,) = struct.unpack_from("<I", raw, len(raw) - 4)
    end = len(raw) - 4 - 4 * n_restarts
    i, key, out = 0, b"", []
    while i < end:
        shared, i = _get_varint(raw, i)
        unshared, i = _get_varint(raw, i)
        vlen, i = _get_varint(raw, i)
        key = key[:shared] + raw[i:i + unshared]
        i += unshared
        out.append((key, raw[i:i + vlen]))
        i += vlen
    return out


def read_index(path: str, verify: bool = True) -> "OrderedDict[str, BundleEntry]":
    """Variable name -> BundleEntry for every tensor listed in ``<prefix>.index``."""
    buf = open(path, "rb").read()
    if len(buf) < 48 or struct.unpack_from("<Q", buf, len(buf) - 8)[0] != MAGIC:
        raise ValueError(f"{path}: not a TensorFlow bundle index (bad magic)")
    foot = buf[-48:]
    _, i = _get_varint(foot, 0)
    _, i = _get_varint(foot, i)
    idx_off, i = _get_varint(foot, i)
    idx_size, i = _get_varint(foot, i)
    entries = OrderedDict()
    for _, handle in _read_block(buf, idx_off, idx_size, verify):
        off, j = _get_varint(handle, 0)
        size, j = _get_varint(handle, j)
        for key, val in _read_block(buf, off, size, verify):
            if key == b"":
                if val[:2] != b"\x08\x01":
                    raise ValueError("multi-shard bundles are not supported")
                continue
            entries[key.decode()] = _decode_entry(val)
    return entries


def load_checkpoint(prefix: str, verify_crc: bool = True) -> "OrderedDict[str, np.ndarray]":
    """All tensors of the bundle ``prefix`` (``prefix.index`` + ``prefix.data-00000-of-00001``)."""
    entries = read_index(prefix + ".index")
    data = open(prefix + ".data-00000-of-00001", "rb").read()
    out = OrderedDict()
    for name, e in entries.items():
        raw = data[e.offset:e.offset + e.size]
        if len(raw) != e.size:
            raise ValueError(f"{name}: data shard is truncated")
        if verify_crc and mask_crc(crc32c(raw)) != e.crc32c:
            raise ValueError(f"{name}: tensor checksum mismatch")
        out[name] = np.frombuffer(raw, dtype=_DTYPES[e.dtype]).reshape(e.shape).copy()
    return out


# ---- table writing ---------------------------------------------------------------------------------------
def _build_block(items) -> bytes:
    out, restarts, last = bytearray(), [], b""
    for n, (key, val) in enumerate(items):
        shared = 0
        if n % _RESTART_INTERVAL == 0:
            restarts.append(len(out))
        else:
            m = min(len(key), len(last))
            while shared < m and key[shared] == last[shared]:
                shared += 1
        out += _put_varint(shared) + _put_varint(len(key) - shared) + _put_varint(len(val)) + key[shared:] + val
        last = key
    if not restarts:
        restarts = [0]
    for r in restarts:
        out += struct.pack("<I", r)
    out += struct.pack("<I", len(restarts))
    return bytes(out)


def _short_successor(key: bytes) -> bytes:
    for i, c in enumerate(key):
        if c != 0xFF:
            return key[:i] + bytes([c + 1])
    return key


def _short_separator(a: bytes, b: bytes) -> bytes:
    m = min(len(a), len(b))
    i = 0
    while i < m and a[i] == b[i]:
        i += 1
    if i < m and a[i] < 0xFF and a[i] + 1 < b[i]:
        return a[:i] + bytes([a[i] + 1])
    return a


def encode_index(entries: "OrderedDict[str, BundleEntry]", block_size: int = 262144) -> bytes:
    """Bytes of an ``.index`` file for the given entries (sorted by key, like tf.train.Saver writes them)."""
    items = [(b"", _HEADER)] + [(k.encode(), _encode_entry(e)) for k, e in sorted(entries.items())]
    out = bytearray()
    index_items, cur, cur_size = [], [], 0

    def flush(next_key):
        nonlocal cur, cur_size
        if not cur:
            return
        blk = _build_block(cur)
        handle = _put_varint(len(out)) + _put_varint(len(blk))
        sep = _short_separator(cur[-1][0], next_key) if next_key is not None else _short_successor(cur[-1][0])
        out.extend(blk + b"\x00" + struct.pack("<I", mask_crc(crc32c(blk + b"\x00"))))
        index_items.append((sep, handle))
        cur, cur_size = [], 0

    for n, (k, v) in enumerate(items):
        cur.append((k, v))
        cur_size += len(k) + len(v) + 3
        if cur_size >= block_size and n + 1 < len(items):
            flush(items[n + 1][0])
    flush(None)
    meta = _build_block([])
    meta_handle = _put_varint(len(out)) + _put_varint(len(meta))
    out.extend(meta + b"\x00" + struct.pack("<I", mask_crc(crc32c(meta + b"\x00"))))
    idx = _build_block(index_items)
    idx_handle = _put_varint(len(out)) + _put_varint(len(idx))
    out.extend(idx + b"\x00" + struct.pack("<I", mask_crc(crc32c(idx + b"\x00"))))
    foot = meta_handle + idx_handle
    out.extend(foot + b"\x00" * (40 - len(foot)) + struct.pack("<Q", MAGIC))
    return bytes(out)


def save_checkpoint(prefix: str, tensors) -> "OrderedDict[str, BundleEntry]":
    """Write ``tensors`` (name -> array-like, float32 or int32) as a single-shard bundle."""
    entries, blob, off = OrderedDict(), bytearray(), 0
    for name in sorted(tensors):
        a = np.asarray(tensors[name])
        if a.dtype.kind == "f":
            a, dt = a.astype("<f4"), DT_FLOAT
        elif a.dtype.kind in "iu":
            a, dt = a.astype("<i4"), DT_INT32
        else:
            raise TypeError(f"{name}: unsupported dtype {a.dtype}")
        raw = np.ascontiguousarray(a).tobytes()
        entries[name] = BundleEntry(dt, tuple(int(s) for s in a.shape), 0, off, len(raw), mask_crc(crc32c(raw)))
        blob += raw
        off += len(raw)
    with open(prefix + ".data-00000-of-00001", "wb") as f:
        f.write(bytes(blob))
    with open(prefix + ".index", "wb") as f:
        f.write(encode_index(entries))
    return entries


# ---- AIRModel <-> reference variable names ----------------------------------------------------------------
def _graph_name(scope, key):
    """ParamStore key -> TF variable name: everything lives under <scope>/rnn/ (variable_scope("rnn"),
    air_model.py:537) except the CNN front-end, created under <scope>/cnn/ (:511)."""
    return f"{scope}/{key}" if key.startswith("cnn/") else f"{scope}/rnn/{key}"


def model_tensors(store, scope="air", with_optimizer=True):
    """The reference's checkpoint contents (model/air-model.index) for a ParamStore: variables under
    ``<scope>/rnn/``, ``<scope>/global_step`` and, optionally, the Adam slots / beta powers under
    ``<scope>/training/`` (air_model.py:69-70, 537, 654, 692)."""
    out = OrderedDict()
    out[f"{scope}/global_step"] = np.asarray(store.global_step, dtype=np.int32)
    for k, v in store.named_views().items():
        out[_graph_name(scope, k)] = v.detach().cpu().numpy()
    if with_optimizer:
        m, vv = store.named_adam()
        for k in m:
            out[f"{scope}/training/{_graph_name(scope, k)}/Adam"] = m[k].detach().cpu().numpy()
            out[f"{scope}/training/{_graph_name(scope, k)}/Adam_1"] = vv[k].detach().cpu().numpy()
        st = store.state.detach().cpu().numpy()
        out[f"{scope}/training/beta1_power"] = np.float32(st[0])
        out[f"{scope}/training/beta2_power"] = np.float32(st[1])
    return out


def save_model(store, prefix, scope="air", with_optimizer=True):
    return save_checkpoint(prefix, model_tensors(store, scope, with_optimizer))


def restore_model(store, prefix, scope="air", verify_crc=True):
    """Load a reference (or own) checkpoint into a ParamStore; Adam slots are optional."""
    import torch
    t = load_checkpoint(prefix, verify_crc)
    pre, cpre = f"{scope}/rnn/", f"{scope}/cnn/"
    named = {k[len(pre):]: torch.from_numpy(v) for k, v in t.items() if k.startswith(pre)}
    named.update({k[len(scope) + 1:]: torch.from_numpy(v) for k, v in t.items() if k.startswith(cpre)})
    store.load_named(named)
    if f"{scope}/global_step" in t:
        store.global_step = int(t[f"{scope}/global_step"])
    m, vv = store.named_adam()
    for k in m:
        tk = f"{scope}/training/{_graph_name(scope, k)}"
        if tk + "/Adam" in t:
            m[k].copy_(torch.from_numpy(t[tk + "/Adam"]))
            vv[k].copy_(torch.from_numpy(t[tk + "/Adam_1"]))
    if f"{scope}/training/beta1_power" in t:
        store.state[0] = float(t[f"{scope}/training/beta1_power"])
        store.state[1] = float(t[f"{scope}/training/beta2_power"])
    return t
