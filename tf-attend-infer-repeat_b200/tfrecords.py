"""Reader / writer for the reference's multi-MNIST TFRecord files, without TensorFlow.

The reference stores its datasets as TFRecords of ``tf.train.Example`` (multi_mnist.py:186-212) and feeds the
model from them (``read_and_decode`` :228-251 for training batches, ``read_test_data`` :254-296 for the test
set).  Schema per example: ``height``, ``width``, ``digits`` (int64 lists), ``indices``, ``positions``, ``boxes``,
``labels`` (int32 arrays as bytes) and ``image`` (canvas*canvas float32 as bytes).

File framing (TFRecordWriter): for every record  uint64 length | uint32 masked_crc32c(length) | data |
uint32 masked_crc32c(data), little endian.  CRC-32C comes from checkpoint.py (C implementation when built).
"""
from __future__ import annotations

import struct

import numpy as np

from .checkpoint import _get_varint, _put_varint, crc32c, mask_crc


# ---- tf.train.Example encoding -----------------------------------------------------------------------
def _ld(field: int, payload: bytes) -> bytes:  # length-delimited field
    return _put_varint((field << 3) | 2) + _put_varint(len(payload)) + payload


def _feature_int64(values) -> bytes:
    packed = b"".join(_put_varint(int(v) & 0xFFFFFFFFFFFFFFFF) for v in values)
    return _ld(3, _ld(1, packed))            # Feature.int64_list{ value (packed) }


def _feature_bytes(value: bytes) -> bytes:
    return _ld(1, _ld(1, value))             # Feature.bytes_list{ value }


def encode_example(features: dict) -> bytes:
    """features: name -> bytes (bytes feature) or list/array of ints (int64 feature)."""
    entries = b""
    for name in sorted(features):
        v = features[name]
        feat = _feature_bytes(v) if isinstance(v, (bytes, bytearray)) else _feature_int64(v)
        entries += _ld(1, _ld(1, name.encode()) + _ld(2, feat))   # Features.feature map entry {key, value}
    return _ld(1, entries)                                          # Example.features


def _fields(buf: bytes):
    i = 0
    while i < len(buf):
        tag, i = _get_varint(buf, i)
        field, wt = tag >> 3, tag & 7
        if wt == 0:
            val, i = _get_varint(buf, i)
        elif wt == 2:
            ln, i = _get_varint(buf, i)
            val = buf[i:i + ln]
            i += ln
        elif wt == 5:
            val = buf[i:i + 4]
            i += 4
        elif wt == 1:
            val = buf[i:i + 8]
            i += 8
        else:
            raise ValueError(f"unsupported wire type {wt}")
        yield field, wt, val


def decode_example(data: bytes) -> dict:
    """-> name -> bytes (bytes_list[0]) | list[int] (int64_list) | list[float] (float_list)."""
    out = {}
    for f, _, features in _fields(data):
        if f != 1:
            continue
        for f2, _, entry in _fields(features):
            if f2 != 1:
                continue
            key, feat = None, None
            for f3, _, v in _fields(entry):
                if f3 == 1:
                    key = v.decode()
                elif f3 == 2:
                    feat = v
            for kind, _, lst in _fields(feat or b""):
                if kind == 1:      # bytes_list
                    vals = [v for f4, _, v in _fields(lst) if f4 == 1]
                    out[key] = vals[0] if len(vals) == 1 else vals
                elif kind == 3:    # int64_list (packed or not)
                    ints = []
                    for f4, wt, v in _fields(lst):
                        if f4 != 1:
                            continue
                        if wt == 0:
                            ints.append(v)
                        else:
                            j = 0
                            while j < len(v):
                                x, j = _get_varint(v, j)
                                ints.append(x)
                    out[key] = [x - (1 << 64) if x >> 63 else x for x in ints]
                elif kind == 2:    # float_list (packed)
                    fl = []
                    for f4, wt, v in _fields(lst):
                        if f4 == 1:
                            fl.extend(struct.unpack(f"<{len(v) // 4}f", v))
                    out[key] = fl
    return out


# ---- record framing -------------------------------------------------------------------------------------
def write_record(f, data: bytes) -> None:
    head = struct.pack("<Q", len(data))
    f.write(head + struct.pack("<I", mask_crc(crc32c(head))) + data + struct.pack("<I", mask_crc(crc32c(data))))


def iter_records(path: str, verify_crc: bool = True):
    with open(path, "rb") as f:
        while True:
            head = f.read(12)
            if not head:
                return
            if len(head) < 12:
                raise ValueError(f"{path}: truncated record header")
            (n,), (hcrc,) = struct.unpack("<Q", head[:8]), struct.unpack("<I", head[8:])
            if verify_crc and mask_crc(crc32c(head[:8])) != hcrc:
                raise ValueError(f"{path}: corrupt record length")
            data = f.read(n)
            tail = f.read(4)
            if len(data) < n or len(tail) < 4:
                raise ValueError(f"{path}: truncated record")
            if verify_crc and mask_crc(crc32c(data)) != struct.unpack("<I", tail)[0]:
                raise ValueError(f"{path}: corrupt record data")
            yield data


# ---- the reference's dataset functions ------------------------------------------------------------------
def write_to_records(filename, images, indices, positions, boxes, labels, digits):
    """multi_mnist.py:186-212 (writes ``filename + '.tfrecords'``)."""
    rows, cols = np.asarray(images[0]).shape
    with open(filename + ".tfrecords", "wb") as f:
        for i in range(len(images)):
            a32 = lambda v: np.asarray(v, dtype=np.int32).tobytes()
            write_record(f, encode_example({
                "height": [rows], "width": [cols], "digits": [int(digits[i])],
                "indices": a32(indices[i]), "positions": a32(positions[i]), "boxes": a32(boxes[i]),
                "labels": a32(labels[i]), "image": np.ravel(np.asarray(images[i], dtype=np.float32)).tobytes()}))


def read_test_data(filename, shift_zero_digits_images=False):
    """multi_mnist.py:254-296: (images, digits, indices, positions, boxes, labels) lists; with
    ``shift_zero_digits_images`` the first empty image stays first and the other empty ones move to the end."""
    images_list, digits_list, indices_list, positions_list, boxes_list, labels_list = [], [], [], [], [], []
    for rec in iter_records(filename):
        ex = decode_example(rec)
        d = int(ex["digits"][0])
        images_list.append(np.frombuffer(ex["image"], dtype=np.float32).copy())
        digits_list.append(d)
        indices_list.append(np.frombuffer(ex["indices"], dtype=np.int32)[:d].copy())
        positions_list.append(np.frombuffer(ex["positions"], dtype=np.int32)[:d * 2].copy())
        boxes_list.append(np.frombuffer(ex["boxes"], dtype=np.int32)[:d * 2].copy())
        labels_list.append(np.frombuffer(ex["labels"], dtype=np.int32)[:d].copy())
    if shift_zero_digits_images:
        # one stable permutation does the reference's three-way concatenation (multi_mnist.py:284-294): rank 0 for the
        # first empty canvas, rank 1 for every canvas with digits, rank 2 for the remaining empty ones
        digits_arr = np.asarray(digits_list)
        rank = np.where(digits_arr > 0, 1, 2)
        rank[np.flatnonzero(digits_arr == 0)[0]] = 0    # (IndexError without an empty canvas, as in the reference)
        order = np.argsort(rank, kind="stable")
        images_list, digits_list = np.asarray(images_list)[order], digits_arr[order]
    return images_list, digits_list, indices_list, positions_list, boxes_list, labels_list


class TFRecordFile:
    """A memory-mapped multi-MNIST TFRecord file with an index built ONCE by the native scanner
    (air_tfrecord_index: framing + CRCs + tf.train.Example parse): ``digits`` [n] int32 and the byte offset of every
    record's ``image`` payload.  Batches are then plain multi-threaded gathers out of the page cache
    (air_gather_rows) -- no per-record Python work, unlike the reference's 4 reader threads + shuffle queue
    (multi_mnist.py:228-251, training.py:76-81)."""

    def __init__(self, path, canvas_size=50, verify_crc=True):
        import ctypes
        import mmap
        from . import _cabi as C
        self._C, self._ct = C, ctypes
        self.path, self.row_bytes = path, canvas_size * canvas_size * 4
        self._f = open(path, "rb")
        size = self._f.seek(0, 2)
        self._mm = mmap.mmap(self._f.fileno(), 0, access=mmap.ACCESS_READ) if size else None
        self._buf = np.frombuffer(self._mm, dtype=np.uint8) if size else np.zeros(0, np.uint8)
        lib, base = C.lib(), self._buf.ctypes.data
        n = int(lib.air_tfrecord_index(base, size, int(verify_crc), 0, None, None, None))
        if n < 0:
            raise ValueError(f"{path}: {lib.air_last_error().decode()}")
        self.offsets, lens = np.zeros(n, np.uint64), np.zeros(n, np.uint32)
        self.digits = np.zeros(n, np.int32)
        if n:
            got = int(lib.air_tfrecord_index(base, size, 0, n, self.offsets.ctypes.data, lens.ctypes.data, self.digits.ctypes.data))
            if got != n:
                raise ValueError(f"{path}: {lib.air_last_error().decode()}")
            if not (lens == self.row_bytes).all():
                raise ValueError(f"{path}: record image has {int(lens[lens != self.row_bytes][0]) // 4} pixels, "
                                 f"expected {canvas_size}^2")

    def __len__(self):
        return len(self.digits)

    def gather(self, idx, out=None, threads=None):
        """images [len(idx), canvas^2] float32 of the records ``idx`` (int array) into ``out`` (a CPU tensor, ideally
        pinned) -> (images, digits int32 tensor)."""
        import os
        import torch
        idx = np.ascontiguousarray(idx, dtype=np.int64)
        n = len(idx)
        if out is None:
            out = torch.empty(n, self.row_bytes // 4)
        if out.shape[0] < n or out.shape[1] * 4 != self.row_bytes or out.dtype != torch.float32 or not out.is_contiguous():
            raise ValueError("gather: bad output buffer")
        offs = np.ascontiguousarray(self.offsets[idx])
        self._C.check(self._C.lib().air_gather_rows(self._buf.ctypes.data, offs.ctypes.data, n, self.row_bytes, out.data_ptr(),
                                                    threads or os.cpu_count() or 1), "air_gather_rows")
        return out[:n], torch.from_numpy(self.digits[idx])

    def shuffle_order(self, shuffle_buffer=10000, seed=0, epochs=1):
        """Record indices in the order a shuffle queue with min_after_dequeue = shuffle_buffer emits them."""
        order = np.zeros(len(self) * epochs, np.int64)
        self._C.check(self._C.lib().air_shuffle_order(len(self), int(shuffle_buffer), int(epochs), int(seed), order.ctypes.data),
                      "air_shuffle_order")
        return order

    def batches(self, batch_size, shuffle_buffer=10000, seed=0, epochs=1, pin_memory=True, ring=3):
        """(images [batch, canvas^2] float32, digits [batch] int32) batches in shuffle-queue order; incomplete
        trailing batches are dropped, as shuffle_batch does.  Pinned buffers are reused round-robin (``ring`` of them):
        a batch stays valid until ``ring - 1`` further batches have been drawn."""
        import torch
        order = self.shuffle_order(shuffle_buffer, seed, epochs)
        pin = pin_memory and torch.cuda.is_available()
        bufs = [torch.empty(batch_size, self.row_bytes // 4, pin_memory=pin) for _ in range(max(ring, 1))]
        for k in range(len(order) // batch_size):
            imgs, digs = self.gather(order[k * batch_size:(k + 1) * batch_size], out=bufs[k % len(bufs)])
            yield imgs, (digs.pin_memory() if pin else digs)


def read_and_decode(filename, batch_size, canvas_size, shuffle_buffer=10000, seed=0, epochs=1, pin_memory=True):
    """Batches of (images [batch, canvas*canvas] float32, digits [batch] int32) like multi_mnist.py:228-251
    (TFRecordReader + shuffle_batch with min_after_dequeue=10000), ready for ``AIRModel.feed``.  Each batch owns its
    memory (use TFRecordFile.batches for the zero-allocation ring of pinned buffers)."""
    f = TFRecordFile(filename, canvas_size)
    for imgs, digs in f.batches(batch_size, shuffle_buffer, seed, epochs, pin_memory=False, ring=1):
        imgs, digs = imgs.clone(), digs.clone()
        if pin_memory:
            import torch
            if torch.cuda.is_available():
                imgs, digs = imgs.pin_memory(), digs.pin_memory()
        yield imgs, digs
