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
