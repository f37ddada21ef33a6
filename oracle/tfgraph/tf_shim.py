"""A numpy stand-in for the handful of ``tf.*`` calls made by the reference's ``air/transformer.py`` and
``air/concrete.py``, so that THEIR SOURCE runs unmodified without TensorFlow (test infrastructure).

    with tf_shim.installed(uniform=u):                    # sys.modules["tensorflow"] = this module, temporarily
        ref = tf_shim.load_reference_module("/root/reference/air/transformer.py")
    out = ref.transformer(U, theta, (28, 28))

Tensors are plain float32 / int32 numpy arrays, every function is eager, one rounding per op like TF's executor
(matmul with a float inner dimension sums separately rounded products in k order: no FMA, as oracle/tfgraph/interp.py).
``tf.random_uniform`` returns the array given to ``installed(uniform=...)`` (noise is always injected here).
Complements the graph interpreter: the saved graph pins what training.py built; this pins the functions of the checked-in
source that the graph does not contain (batch_transformer, concrete_binary_sample with hard=True)."""
import builtins
import contextlib
import importlib.util
import sys
import types

import numpy as np

f32 = np.float32
_uniform = [None]


class _Arr(np.ndarray):
    """ndarray with the two TF tensor methods the reference calls (get_shape().as_list())."""

    def get_shape(self):
        return types.SimpleNamespace(as_list=lambda: list(self.shape))


def tensor(a):
    return np.asarray(a).view(_Arr)


def _dt(d):
    return {"float32": np.float32, "int32": np.int32}.get(d, d)


def shape(x):
    return np.array(np.shape(x), np.int32)


def cast(x, dtype):
    return np.asarray(x).astype(_dt(dtype))


def zeros(shape, dtype="float32"):
    return np.zeros(tuple(int(s) for s in np.atleast_1d(shape)) if np.size(shape) else (), _dt(dtype))


def ones(shape, dtype="float32"):
    return np.ones(tuple(int(s) for s in np.atleast_1d(shape)), _dt(dtype))


def ones_like(x):
    return np.ones_like(x)


def stack(values, axis=0):
    return np.stack([np.asarray(v) for v in values], axis)


def transpose(x, perm):
    return np.transpose(x, perm)


def expand_dims(x, axis):
    return np.expand_dims(x, axis)


def reshape(x, shp):
    return np.reshape(x, tuple(int(s) for s in np.atleast_1d(shp)))


def matmul(a, b):
    a, b = np.asarray(a), np.asarray(b)
    if a.dtype.kind != "f":
        return np.matmul(a, b)
    prod = a[..., :, :, None] * b[..., None, :, :]
    acc = prod[..., :, 0, :]
    for k in builtins.range(1, prod.shape[-2]):
        acc = acc + prod[..., :, k, :]
    return acc


def floor(x):
    return np.floor(x)


def clip_by_value(x, lo, hi):
    return np.minimum(np.maximum(x, lo), hi)


def range(n):          # noqa: A001 - mirrors tf.range
    return np.arange(int(n), dtype=np.int32)


def gather(params, indices):
    return np.take(params, np.asarray(indices), axis=0)


def add_n(xs):
    acc = xs[0]
    for x in xs[1:]:
        acc = acc + x
    return acc


def linspace(start, stop, num):
    n = int(num)
    start, stop = f32(start), f32(stop)
    if n == 1:
        return np.array([start], f32)
    step = f32((stop - start) / f32(n - 1))
    return (start + step * np.arange(n, dtype=f32)).astype(f32)


def concat(axis, values):
    return np.concatenate(values, axis)


def tile(x, multiples):
    return np.tile(x, tuple(int(m) for m in multiples))


def slice(x, begin, size):   # noqa: A001 - mirrors tf.slice
    return x[tuple(builtins.slice(int(b), None if int(s) == -1 else int(b) + int(s)) for b, s in zip(begin, size))]


def log(x):
    return np.log(x)


def exp(x):
    return np.exp(x)


def round(x):          # noqa: A001 - tf.round: half to even
    return np.rint(x)


def stop_gradient(x):
    return x


def random_uniform(shp, minval=0, maxval=1):
    u = np.asarray(_uniform[0], f32)
    assert tuple(int(s) for s in shp) == u.shape
    return u


_scopes = []
_params = [None]
_normal = [None]


@contextlib.contextmanager
def variable_scope(name, *a, **k):
    _scopes.append(name)
    try:
        yield types.SimpleNamespace(name="/".join(_scopes))
    finally:
        _scopes.pop()


def random_normal(shp):
    """Next tensor of the injected normal-noise sequence (``installed(normal=[...])``)."""
    n = np.asarray(_normal[0].pop(0), f32)
    assert tuple(int(s) for s in shp) == n.shape, (tuple(shp), n.shape)
    return n


def sqrt(x):
    return np.sqrt(x)


def _softplus(x):
    thr = f32(np.log(np.finfo(np.float32).eps) + 2.0)     # softplus_op.h: x > -thr -> x, x < thr -> exp(x)
    e = np.exp(x)
    return np.where(x > -thr, x, np.where(x < thr, e, np.log(e + f32(1.0)))).astype(np.float32)


def _fully_connected(inputs, num_outputs, activation_fn=None, scope=None):
    """tf.contrib.layers.fully_connected: activation(inputs @ weights + biases); the variables come from the
    parameter dict given to ``installed(params=...)`` under ``<scope name>/weights|biases``."""
    w, b = (np.asarray(_params[0][f"{scope.name}/{n}"], f32) for n in ("weights", "biases"))
    assert w.shape == (inputs.shape[1], num_outputs)
    y = np.matmul(inputs, w) + b
    return activation_fn(y) if activation_fn is not None else y


layers = types.ModuleType("tensorflow.contrib.layers")
layers.fully_connected = _fully_connected
contrib = types.ModuleType("tensorflow.contrib")
contrib.layers = layers


nn = types.SimpleNamespace(sigmoid=lambda x: (f32(1.0) / (f32(1.0) + np.exp(-x))).astype(np.asarray(x).dtype),
                           softplus=_softplus)


@contextlib.contextmanager
def installed(uniform=None, normal=None, params=None):
    """Make ``import tensorflow`` (and ``tensorflow.contrib.layers``) resolve to this module inside the block, with
    the injected uniform noise, normal-noise sequence and variable values."""
    names = {"tensorflow": sys.modules[__name__], "tensorflow.contrib": contrib, "tensorflow.contrib.layers": layers}
    saved = {k: sys.modules.get(k) for k in names}
    sys.modules.update(names)
    _uniform[0], _normal[0], _params[0] = uniform, list(normal) if normal is not None else None, params
    try:
        yield
    finally:
        _uniform[0] = _normal[0] = _params[0] = None
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v


def load_reference_module(path):
    spec = importlib.util.spec_from_file_location("ref_" + path.replace("/", "_").replace(".", "_"), path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


# ---- tf.train.Example / tf.python_io.TFRecordWriter (for the reference's multi_mnist.py:186-212) -------------------
def example_classes():
    """tf.train.{Example, Features, Feature, BytesList, FloatList, Int64List} built at run time with the protobuf
    library from the published schema (tensorflow/core/example/{example,feature}.proto)."""
    from google.protobuf import descriptor_pb2, descriptor_pool, message_factory
    fd = descriptor_pb2.FileDescriptorProto(name="air_shim_example.proto", package="airshim", syntax="proto3")
    F = descriptor_pb2.FieldDescriptorProto

    def msg(name, *fields):
        m = fd.message_type.add(name=name)
        for fname, num, typ, label, tname in fields:
            f = m.field.add(name=fname, number=num, type=typ, label=label)
            if tname:
                f.type_name = ".airshim." + tname
        return m
    msg("BytesList", ("value", 1, F.TYPE_BYTES, F.LABEL_REPEATED, None))
    msg("FloatList", ("value", 1, F.TYPE_FLOAT, F.LABEL_REPEATED, None))
    msg("Int64List", ("value", 1, F.TYPE_INT64, F.LABEL_REPEATED, None))
    feat = msg("Feature", ("bytes_list", 1, F.TYPE_MESSAGE, F.LABEL_OPTIONAL, "BytesList"),
               ("float_list", 2, F.TYPE_MESSAGE, F.LABEL_OPTIONAL, "FloatList"),
               ("int64_list", 3, F.TYPE_MESSAGE, F.LABEL_OPTIONAL, "Int64List"))
    feat.oneof_decl.add(name="kind")
    for f in feat.field:
        f.oneof_index = 0
    feats = msg("Features", ("feature", 1, F.TYPE_MESSAGE, F.LABEL_REPEATED, "Features.FeatureEntry"))
    entry = feats.nested_type.add(name="FeatureEntry")
    entry.options.map_entry = True
    entry.field.add(name="key", number=1, type=F.TYPE_STRING, label=F.LABEL_OPTIONAL)
    entry.field.add(name="value", number=2, type=F.TYPE_MESSAGE, label=F.LABEL_OPTIONAL, type_name=".airshim.Feature")
    msg("Example", ("features", 1, F.TYPE_MESSAGE, F.LABEL_OPTIONAL, "Features"))
    pool = descriptor_pool.DescriptorPool()
    pool.Add(fd)
    return types.SimpleNamespace(**{n: message_factory.GetMessageClass(pool.FindMessageTypeByName("airshim." + n))
                                    for n in ("Example", "Features", "Feature", "BytesList", "FloatList", "Int64List")})


def _crc32c_table():
    tab = []
    for i in builtins.range(256):
        c = i
        for _ in builtins.range(8):
            c = (c >> 1) ^ 0x82F63B78 if c & 1 else c >> 1
        tab.append(c)
    return tab


_CRC = _crc32c_table()


def masked_crc32c(data):
    """CRC-32C (Castagnoli), then TFRecord's mask: rotate right by 15 and add 0xa282ead8 (lib/hash/crc32c.h)."""
    c = 0xFFFFFFFF
    for b in data:
        c = _CRC[(c ^ b) & 0xFF] ^ (c >> 8)
    c ^= 0xFFFFFFFF
    return (((c >> 15) | (c << 17)) + 0xA282EAD8) & 0xFFFFFFFF


class TFRecordWriter:
    """tf.python_io.TFRecordWriter: u64 length, masked crc of the length, payload, masked crc of the payload."""

    def __init__(self, path):
        self.f = open(path, "wb")

    def write(self, record):
        import struct
        head = struct.pack("<Q", len(record))
        self.f.write(head + struct.pack("<I", masked_crc32c(head)) + record + struct.pack("<I", masked_crc32c(record)))

    def close(self):
        self.f.close()


def tf_record_iterator(path):
    """tf.python_io.tf_record_iterator: yields the payloads, checking both masked CRCs (independent of tfrecords.py)."""
    import struct
    with open(path, "rb") as f:
        while True:
            head = f.read(8)
            if not head:
                return
            (n,) = struct.unpack("<Q", head)
            assert struct.unpack("<I", f.read(4))[0] == masked_crc32c(head), "length crc"
            data = f.read(n)
            assert struct.unpack("<I", f.read(4))[0] == masked_crc32c(data), "payload crc"
            yield data


python_io = types.SimpleNamespace(TFRecordWriter=TFRecordWriter, tf_record_iterator=tf_record_iterator)


class _Train:
    def __getattr__(self, name):
        return getattr(example_classes_cached(), name)


_example_cache = []


def example_classes_cached():
    if not _example_cache:
        _example_cache.append(example_classes())
    return _example_cache[0]


train = _Train()


class NumpyCompat:
    """``np`` for reference modules written against NumPy 1.x: ndarray.tostring() (removed in NumPy 2.3) is provided
    on the arrays returned by asarray / ravel."""

    class _A(np.ndarray):
        def tostring(self):
            return self.tobytes()

    def __getattr__(self, name):
        return getattr(np, name)

    def asarray(self, *a, **k):
        return np.asarray(*a, **k).view(self._A)

    def ravel(self, *a, **k):
        return np.ravel(*a, **k).view(self._A)

    @staticmethod
    def fromstring(data, dtype=float):                   # NumPy 1.x binary mode == frombuffer + copy
        return np.frombuffer(data, dtype=dtype).copy()
