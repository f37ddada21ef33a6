"""A numpy stand-in for the ``tf.*`` calls made by the reference's ``air/transformer.py``, ``air/concrete.py``,
``air/vae.py``, ``air/air_model.py`` (forward model, train=False) and ``multi_mnist.py`` (TFRecord writer / reader), so
that THEIR SOURCE runs unmodified without TensorFlow (test infrastructure).

    with tf_shim.installed(uniform=u):                    # sys.modules["tensorflow"] = this module, temporarily
        ref = tf_shim.load_reference_module("/root/reference/air/transformer.py")
    out = ref.transformer(U, theta, (28, 28))

Tensors are float32 / int32 numpy arrays, every function is eager, one rounding per op like TF's executor (matmul with
a float inner dimension of the Spatial Transformer sums separately rounded products in k order: no FMA, as
oracle/tfgraph/interp.py).  ``tf.while_loop`` is a Python loop, TensorArrays are lists, variables are looked up in the
parameter dict given to ``installed(params=...)``, ``tf.random_uniform`` / ``tf.random_normal`` serve the injected noise
in call order, summaries are collected in ``summaries``.  Not supported: anything that needs autodiff (train=True).
Complements the graph interpreter: the saved graph pins what training.py built (including the gradient graph); this
pins the checked-in revision of the source, and the parts no saved graph contains (batch_transformer,
concrete_binary_sample with hard=True, the cnn=True front-end's wiring, the TFRecord functions)."""
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


def transpose(x, perm=None):
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


def concat(values, axis):
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
    src = _uniform[0]
    u = np.asarray(src.pop(0) if isinstance(src, list) else src, f32)
    assert tuple(int(s) for s in shp) == u.shape
    return u


_scopes = []
_params = [None]
_normal = [None]


@contextlib.contextmanager
def variable_scope(name, *a, **k):
    _scopes.append(name)
    try:
        yield types.SimpleNamespace(name="/".join(_scopes), reuse=bool(k.get("reuse")))
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


def _param(name):
    """Variable value by its full scope path, or by the path with the model's ``<scope>/rnn/`` prefix removed (the
    key convention of oracle.init_params / the checkpoint names minus ``air/rnn/``)."""
    p = _params[0]
    if name in p:
        return np.asarray(p[name], f32)
    parts = name.split("/")
    return np.asarray(p["/".join(parts[2:]) if parts[1] == "rnn" else "/".join(parts[1:])], f32)


_RELU = object()


def _fully_connected(inputs, num_outputs, activation_fn=_RELU, scope=None):
    """tf.contrib.layers.fully_connected: activation(inputs @ weights + biases), ReLU unless told otherwise; the
    variables come from the parameter dict given to ``installed(params=...)``."""
    if activation_fn is _RELU:
        activation_fn = nn.relu
    w, b = (_param(f"{scope.name}/{n}") for n in ("weights", "biases"))
    assert w.shape == (inputs.shape[1], num_outputs)
    y = np.matmul(inputs, w) + b
    return activation_fn(y) if activation_fn is not None else y


contrib_layers = types.ModuleType("tensorflow.contrib.layers")
contrib_layers.fully_connected = _fully_connected
contrib = types.ModuleType("tensorflow.contrib")
contrib.layers = contrib_layers


nn = types.SimpleNamespace(sigmoid=lambda x: (f32(1.0) / (f32(1.0) + np.exp(-x))).astype(np.asarray(x).dtype),
                           softplus=_softplus, relu=lambda x: np.maximum(x, f32(0.0)), tanh=lambda x: np.tanh(x))


@contextlib.contextmanager
def installed(uniform=None, normal=None, params=None, global_step=0):
    """Make ``import tensorflow`` (and ``tensorflow.contrib.layers``) resolve to this module inside the block, with
    the injected uniform noise, normal-noise sequence and variable values."""
    names = {"tensorflow": sys.modules[__name__], "tensorflow.contrib": contrib, "tensorflow.contrib.layers": contrib_layers,
             "tensorflow.contrib.rnn": rnn}
    _global_step[0] = global_step
    summaries.clear()
    saved = {k: sys.modules.get(k) for k in names}
    sys.modules.update(names)
    _uniform[0] = list(uniform) if isinstance(uniform, (list, tuple)) else uniform
    _normal[0], _params[0] = list(normal) if normal is not None else None, params
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


_example_cache = []


def example_classes_cached():
    if not _example_cache:
        _example_cache.append(example_classes())
    return _example_cache[0]


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


# ---- the rest of what air/air_model.py calls (whole-model forward, train=False) -------------------------------------
float32, int32 = np.float32, np.int32
_global_step = [0]
summaries = {}                                            # tag -> value of every tf.summary.scalar / image
sigmoid = nn.sigmoid


def constant(v, dtype=None):
    return np.asarray(v, dtype) if dtype else (np.int32(v) if isinstance(v, int) else np.asarray(v, f32))


def zeros_like(x):
    return np.zeros_like(x)


def less(a, b):
    return np.less(a, b)


def greater(a, b):
    return np.greater(a, b)


def equal(a, b):
    return np.equal(a, b)


def logical_and(a, b):
    return np.logical_and(a, b)


def reduce_any(x):
    return np.any(x)


def reduce_sum(x, axis=None):
    return np.sum(x, axis=axis)


def reduce_mean(x, axis=None):
    import warnings
    with np.errstate(all="ignore"), warnings.catch_warnings():
        warnings.simplefilter("ignore")
        return np.mean(x, axis=axis, dtype=np.asarray(x).dtype)


def square(x):
    return x * x


def minimum(a, b):
    return np.minimum(a, b)


def maximum(a, b):
    return np.maximum(a, b)


def where(c, t, e):
    c = np.asarray(c)
    if c.ndim == 1 and np.ndim(t) > 1:                     # vector condition selects rows
        c = c.reshape((-1,) + (1,) * (np.ndim(t) - 1))
    return np.where(c, t, e).astype(np.asarray(t).dtype)


def boolean_mask(x, mask):
    return np.asarray(x)[np.asarray(mask)]


def pad(x, paddings):
    return np.pad(x, [(int(a), int(b)) for a, b in paddings])


def get_variable(shape=None, dtype=None, initializer=None, trainable=True):
    """Only ``global_step`` is created this way (air_model.py:69); every other variable comes from a layer."""
    return np.int32(_global_step[0])


def constant_initializer(v):
    return v


def get_variable_scope():
    return types.SimpleNamespace(name="/".join(_scopes), reuse=False)


def trainable_variables():
    return []


def while_loop(cond, body, loop_vars):
    v = list(loop_vars)
    while bool(cond(*v)):
        v = list(body(*v))
    return v


class TensorArray:
    def __init__(self, dtype=None, size=0, dynamic_size=True):
        self.items = []

    def size(self):
        return len(self.items)

    def write(self, index, value):
        assert index == len(self.items)
        self.items.append(value)
        return self

    def stack(self):
        return np.stack(self.items)


class _LSTMCell:
    """tf.contrib.rnn.BasicLSTMCell (TF 1.3): [x, h] @ kernel + bias -> i, j, f, o; c' = c*sigmoid(f + 1) +
    sigmoid(i)*tanh(j); h' = tanh(c')*sigmoid(o); state = (c, h)."""

    def __init__(self, num_units, reuse=None):
        self.n = num_units

    def zero_state(self, batch, dtype):
        z = np.zeros((int(batch), self.n), dtype)
        return (z, z.copy())

    def __call__(self, inputs, state, scope=None):
        c, h = state
        gates = np.matmul(np.concatenate([inputs, h], 1), _param(scope.name + "/kernel")) + _param(scope.name + "/bias")
        i, j, f, o = np.split(gates, 4, axis=1)
        new_c = c * nn.sigmoid(f + f32(1.0)) + nn.sigmoid(i) * np.tanh(j)
        new_h = np.tanh(new_c) * nn.sigmoid(o)
        return new_h, (new_c, new_h)


rnn = types.ModuleType("tensorflow.contrib.rnn")
rnn.BasicLSTMCell = _LSTMCell
contrib.rnn = rnn


def _exponential_decay(learning_rate, global_step, decay_steps, decay_rate, staircase=False, name=None):
    p = f32(global_step) / f32(decay_steps)                # Cast, Cast, RealDiv (the op sequence of the saved graph)
    if staircase:
        p = np.floor(p)
    return f32(learning_rate) * np.power(f32(decay_rate), p)


def _scalar(name, value):
    summaries[name] = float(value)


def _image(name, tensor_, max_outputs=3):
    summaries["image:" + name] = np.asarray(tensor_)


summary = types.SimpleNamespace(scalar=_scalar, image=_image, histogram=lambda *a, **k: None)


def _resize_images(images, size):
    """tf.image.resize_images (bilinear, align_corners=False), as oracle/tfgraph/interp.py::op_ResizeBilinear."""
    oh, ow = int(size[0]), int(size[1])
    _, H, W, _ = images.shape

    def axis(n_in, n_out):
        src = np.arange(n_out, dtype=f32) * (f32(n_in) / f32(n_out))
        lo = np.floor(src).astype(np.int64)
        return lo, np.minimum(lo + 1, n_in - 1), (src - lo.astype(f32)).astype(f32)
    y0, y1, ly = axis(H, oh)
    x0, x1, lx = axis(W, ow)
    lx, ly = lx[None, None, :, None], ly[None, :, None, None]
    top = images[:, y0][:, :, x0] + (images[:, y0][:, :, x1] - images[:, y0][:, :, x0]) * lx
    bottom = images[:, y1][:, :, x0] + (images[:, y1][:, :, x1] - images[:, y1][:, :, x0]) * lx
    return top + (bottom - top) * ly


def _draw_bounding_boxes(images, boxes):
    out = np.array(images, copy=True)
    B, H, W, _ = out.shape
    for b in builtins.range(B):
        for ymin, xmin, ymax, xmax in np.asarray(boxes[b]):
            r0, r1, c0, c1 = int(ymin * (H - 1)), int(ymax * (H - 1)), int(xmin * (W - 1)), int(xmax * (W - 1))
            out[b, r0, c0:c1 + 1] = out[b, r1, c0:c1 + 1] = 1.0          # first colour of the table, first channel
            out[b, r0:r1 + 1, c0] = out[b, r0:r1 + 1, c1] = 1.0
    return out


image = types.SimpleNamespace(resize_images=_resize_images, draw_bounding_boxes=_draw_bounding_boxes)


def _conv2d(inputs, filters, kernel_size, strides, padding, activation, reuse=None, name=None):
    """tf.layers.conv2d, 'same', stride 1, NHWC x HWIO; variables ``<scope>/<name>/kernel|bias``."""
    assert padding == "same" and tuple(strides) == (1, 1)
    scope_name = "/".join(_scopes + [name])
    w, b = _param(scope_name + "/kernel"), _param(scope_name + "/bias")
    kh, kw = kernel_size
    B, H, W, _ = inputs.shape
    x = np.pad(inputs, [(0, 0), (kh // 2, kh // 2), (kw // 2, kw // 2), (0, 0)])
    out = np.zeros((B, H, W, filters), f32)
    for ky in builtins.range(kh):
        for kx in builtins.range(kw):
            out += np.matmul(x[:, ky:ky + H, kx:kx + W, :], w[ky, kx])
    return activation(out + b)


def _max_pooling2d(inputs, pool_size, strides, name=None):
    assert tuple(pool_size) == (2, 2) and strides == 2
    B, H, W, C = inputs.shape
    x = inputs[:, :H // 2 * 2, :W // 2 * 2]
    return x.reshape(B, H // 2, 2, W // 2, 2, C).max(axis=(2, 4))


layers = types.SimpleNamespace(conv2d=_conv2d, max_pooling2d=_max_pooling2d)        # tf.layers


class _NoOptimizer:
    """tf.train.AdamOptimizer stand-in for train=True FORWARD runs: there is no autodiff here, so compute_gradients
    reports one variable without a gradient (the source skips ``None`` gradients) and apply_gradients does nothing.
    The gradient graph, clipping and ApplyAdam are pinned by the saved graph instead (oracle/tfgraph/interp.py)."""

    def __init__(self, learning_rate):
        self.learning_rate = learning_rate

    def compute_gradients(self, loss):
        return [(None, types.SimpleNamespace(name="no_autodiff_in_the_shim"))]

    def apply_gradients(self, grads_and_vars, global_step=None):
        return None


def clip_by_global_norm(grads, clip_norm):
    assert all(g is None for g in grads)
    return list(grads), None


class _TrainNS:
    """tf.train: exponential_decay, and the Example / Features / Feature / *List message classes."""
    exponential_decay = staticmethod(_exponential_decay)
    AdamOptimizer = _NoOptimizer

    def __getattr__(self, name):
        return getattr(example_classes_cached(), name)


train = _TrainNS()


def _wrap_module():
    """Every function above returns plain ndarrays; the reference also calls ``.set_shape`` / ``.get_shape`` on
    tensors and applies functions to Python floats, so: results become _Arr views, float arguments become float32."""
    import functools

    def conv(v):
        if isinstance(v, np.ndarray) and not isinstance(v, _Arr):
            return v.view(_Arr)
        if isinstance(v, tuple):
            return tuple(conv(x) for x in v)
        return v

    def wrap(fn):
        @functools.wraps(fn)
        def inner(*a, **k):
            a = [f32(x) if isinstance(x, float) else x for x in a]
            k.pop("name", None)                          # graph-node names mean nothing here
            return conv(fn(*a, **k))
        return inner
    g = globals()
    skip = {"installed", "variable_scope", "load_reference_module", "example_classes", "example_classes_cached",
            "masked_crc32c", "tf_record_iterator", "tensor", "get_variable_scope", "while_loop", "trainable_variables",
            "constant_initializer", "clip_by_global_norm"}
    for name, fn in list(g.items()):
        if isinstance(fn, types.FunctionType) and fn.__module__ == __name__ and not name.startswith("_") and name not in skip:
            g[name] = wrap(fn)
    for ns in (nn, image):
        for name, fn in list(vars(ns).items()):
            setattr(ns, name, wrap(fn))
    contrib_layers.fully_connected = wrap(_fully_connected)
    _LSTMCell.__call__ = (lambda f: lambda self, *a, **k: conv(f(self, *a, **k)))(_LSTMCell.__call__)
    _LSTMCell.zero_state = (lambda f: lambda self, *a, **k: conv(f(self, *a, **k)))(_LSTMCell.zero_state)
    TensorArray.stack = (lambda f: lambda self: conv(f(self)))(TensorArray.stack)


_Arr.set_shape = lambda self, shape: None
_wrap_module()
