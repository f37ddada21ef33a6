"""Minimal protobuf wire-format decoder for TensorFlow's MetaGraphDef / GraphDef / NodeDef / AttrValue / TensorProto
(test infrastructure; no TensorFlow, no generated proto classes).  Field numbers follow the public TF 1.x .proto files
(tensorflow/core/framework/{graph,node_def,attr_value,tensor,tensor_shape,types}.proto, protobuf/meta_graph.proto)."""
import struct

import numpy as np


def _varint(buf, pos):
    r = 0
    shift = 0
    while True:
        b = buf[pos]
        pos += 1
        r |= (b & 0x7F) << shift
        if not b & 0x80:
            return r, pos
        shift += 7


def fields(buf):
    """Yield (field_number, wire_type, value) of one message; value = int (varint/fixed) or memoryview (len-delimited)."""
    pos, n = 0, len(buf)
    while pos < n:
        key, pos = _varint(buf, pos)
        fn, wt = key >> 3, key & 7
        if wt == 0:
            v, pos = _varint(buf, pos)
        elif wt == 1:
            v = bytes(buf[pos:pos + 8])
            pos += 8
        elif wt == 2:
            ln, pos = _varint(buf, pos)
            v = buf[pos:pos + ln]
            pos += ln
        elif wt == 5:
            v = bytes(buf[pos:pos + 4])
            pos += 4
        else:
            raise ValueError(f"wire type {wt}")
        yield fn, wt, v


def _sint(v):
    return v - (1 << 64) if v >= (1 << 63) else v


# DataType enum (types.proto)
DT = {1: np.float32, 2: np.float64, 3: np.int32, 4: np.uint8, 5: np.int16, 6: np.int8, 7: object, 9: np.int64,
      10: np.bool_, 20: "resource", 8: np.complex64}


def shape(buf):
    """TensorShapeProto -> list of ints (None if unknown rank)."""
    dims = []
    for fn, wt, v in fields(buf):
        if fn == 2:
            size = 0
            for f2, _, v2 in fields(v):
                if f2 == 1:
                    size = _sint(v2)
            dims.append(size)
        elif fn == 3 and v:
            return None
    return dims


def _packed(v, wt, fmt, size):
    if wt == 2:
        return list(struct.unpack(f"<{len(v) // size}{fmt}", bytes(v)))
    return [struct.unpack("<" + fmt, v)[0]]


def _packed_varint(v, wt):
    if wt != 2:
        return [_sint(v)]
    out, pos = [], 0
    while pos < len(v):
        x, pos = _varint(v, pos)
        out.append(_sint(x))
    return out


def tensor(buf):
    """TensorProto -> numpy array."""
    dtype, shp, content = None, [], None
    fv, iv, sv, bv, i64v, dv = [], [], [], [], [], []
    for fn, wt, v in fields(buf):
        if fn == 1:
            dtype = v
        elif fn == 2:
            shp = shape(v)
        elif fn == 4:
            content = bytes(v)
        elif fn == 5:
            fv += _packed(v, wt, "f", 4)
        elif fn == 6:
            dv += _packed(v, wt, "d", 8)
        elif fn == 7:
            iv += _packed_varint(v, wt)
        elif fn == 8:
            sv.append(bytes(v))
        elif fn == 10:
            i64v += _packed_varint(v, wt)
        elif fn == 11:
            bv += _packed_varint(v, wt)
    npdt = DT[dtype]
    n = int(np.prod(shp)) if shp else 1
    if npdt is object:
        a = np.empty(n, object)
        for i in range(n):
            a[i] = sv[i] if i < len(sv) else (sv[-1] if sv else b"")
        return a.reshape(shp)
    if content is not None and len(content):
        return np.frombuffer(content, npdt).reshape(shp).copy()
    vals = {np.float32: fv, np.float64: dv, np.int32: iv, np.int64: i64v, np.bool_: bv, np.uint8: iv, np.int16: iv,
            np.int8: iv}[npdt]
    if not vals:
        return np.zeros(shp, npdt)
    a = np.array(vals, npdt)
    if a.size < n:                                        # "last value repeats" rule of TensorProto
        a = np.concatenate([a, np.full(n - a.size, a[-1], npdt)])
    return a.reshape(shp)


def attr(buf):
    """AttrValue -> python value."""
    for fn, wt, v in fields(buf):
        if fn == 2:
            return bytes(v)
        if fn == 3:
            return _sint(v)
        if fn == 4:
            return struct.unpack("<f", v)[0]
        if fn == 5:
            return bool(v)
        if fn == 6:
            return ("dtype", v)
        if fn == 7:
            return ("shape", shape(v))
        if fn == 8:
            return tensor(v)
        if fn == 10:
            return ("func", bytes(v))
        if fn == 1:
            out = []
            for f2, w2, v2 in fields(v):
                if f2 == 2:
                    out.append(bytes(v2))
                elif f2 == 3:
                    out += _packed_varint(v2, w2)
                elif f2 == 4:
                    out += _packed(v2, w2, "f", 4)
                elif f2 == 5:
                    out += [bool(x) for x in _packed_varint(v2, w2)]
                elif f2 == 6:
                    out += [("dtype", x) for x in _packed_varint(v2, w2)]
                elif f2 == 7:
                    out.append(("shape", shape(v2)))
                elif f2 == 8:
                    out.append(tensor(v2))
            return out
    return None


class Node:
    __slots__ = ("name", "op", "inputs", "ctrl", "attr")

    def __init__(self, buf):
        self.inputs, self.ctrl, self.attr = [], [], {}
        for fn, _, v in fields(buf):
            if fn == 1:
                self.name = bytes(v).decode()
            elif fn == 2:
                self.op = bytes(v).decode()
            elif fn == 3:
                s = bytes(v).decode()
                if s.startswith("^"):
                    self.ctrl.append(s[1:])
                else:
                    name, _, idx = s.partition(":")
                    self.inputs.append((name, int(idx) if idx else 0))
            elif fn == 5:
                k, val = None, None
                for f2, _, v2 in fields(v):
                    if f2 == 1:
                        k = bytes(v2).decode()
                    elif f2 == 2:
                        val = attr(v2)
                self.attr[k] = val

    def __repr__(self):
        return f"<{self.op} {self.name}>"


def load_metagraph(path):
    """MetaGraphDef file -> (dict name -> Node in file order, tensorflow_version string)."""
    with open(path, "rb") as f:
        buf = memoryview(f.read())
    nodes, version = {}, None
    for fn, _, v in fields(buf):
        if fn == 1:                                       # MetaInfoDef
            for f2, _, v2 in fields(v):
                if f2 == 5:
                    version = bytes(v2).decode()
        elif fn == 2:                                     # GraphDef
            for f2, _, v2 in fields(v):
                if f2 == 1:
                    nd = Node(v2)
                    nodes[nd.name] = nd
    return nodes, version
