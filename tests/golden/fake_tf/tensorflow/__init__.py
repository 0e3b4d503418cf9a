"""A NumPy stand-in for the handful of TensorFlow 2 operations the reference's hot-path files use, so that those
files can be imported UNMODIFIED from /root/reference and executed in this container (TensorFlow is not installable
here).  Test infrastructure only (tests/golden/make_golden_emulated.py).

"Tensors" are NumPy arrays.  Conventions that matter for parity:
  * floating-point data is float32 throughout; Python scalars are weak (NumPy >= 2 semantics = TF's conversion of
    constants to the tensor dtype);
  * tf.argsort(direction='DESCENDING') is stable on the negated values (ties: lower index first), tf.argmax returns
    the first maximal index, tf.boolean_mask keeps row-major order, tf.scatter_nd sums duplicates,
    divide_no_nan(x, 0) = 0, tf.one_hot(out of range) = 0, reduce_max of an empty axis = lowest float;
  * tf.range(start, limit, delta) in float32 accumulates start + delta + delta ... as TF's CPU kernel does;
  * transcendental functions are NumPy's float32 ones (libm), so values agree with TF / the oracle to a few ulp, not
    bit for bit.
What this pins is the reference's own control flow, operation order and broadcasting — not TensorFlow's kernels."""
import math as _math
import sys as _sys
import types as _types

import numpy as _np

float32 = _np.float32
float64 = _np.float64
int32 = _np.int32
int64 = _np.int64
bool = _np.bool_
string = str
_F = _np.float32


class _T(_np.ndarray):
    """ndarray with the one EagerTensor method the reference's unit tests call"""
    def numpy(self):
        return _np.asarray(self)


def _wrap(a):
    return _np.asarray(a).view(_T)


def _t(x, dtype=None):
    """convert_to_tensor: Python floats -> float32, Python ints -> int32, float64 arrays stay as given unless dtype."""
    if dtype is not None:
        return _np.asarray(x, dtype=dtype)
    if isinstance(x, _np.ndarray) or isinstance(x, _np.generic):
        return x
    a = _np.asarray(x)
    if a.dtype == _np.float64:
        return a.astype(_F)
    if a.dtype == _np.int64:
        return a.astype(_np.int32)
    return a


class _StubMeta(type):
    def __getattr__(cls, name):
        if name.startswith("__"):
            raise AttributeError(name)
        return cls


class _Stub(object, metaclass=_StubMeta):
    """Anything the hot path never executes (Keras layers, initialisers ...): callable, subclassable, attribute-proof."""
    def __init__(self, *a, **k):
        pass

    def __call__(self, *a, **k):
        return _Stub()

    def __getattr__(self, name):
        return _Stub()


class _StubModule(_types.ModuleType):
    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        return _Stub


def function(fn=None, **kwargs):
    if fn is None:
        return lambda f: f
    return fn


def convert_to_tensor(x, dtype=None):
    return _t(x, dtype)


def constant(x, dtype=None, shape=None):
    a = _t(x, dtype)
    return a if shape is None else _np.broadcast_to(a, shape).copy()


def cast(x, dtype):
    return _np.asarray(x).astype(dtype)


def shape(x):
    return _np.asarray(_np.shape(x), dtype=_np.int32)


def reshape(x, s):
    return _np.reshape(x, tuple(int(v) for v in _np.asarray(s).reshape(-1)))


def range(start, limit=None, delta=1, dtype=None):  # noqa: A001
    if limit is None:
        start, limit = 0, start
    is_float = dtype in (_np.float32, _np.float64) or any(isinstance(v, float) or (hasattr(v, "dtype") and _np.asarray(v).dtype.kind == "f")
                                                          for v in (start, limit, delta))
    if not is_float:
        return _np.arange(int(start), int(limit), int(delta), dtype=dtype or _np.int32)
    dt = dtype or _np.float32
    s, l, d = dt(start), dt(limit), dt(delta)
    n = int(_math.ceil(abs((float(l) - float(s)) / float(d))))   # RangeSize in tensorflow/core/kernels/sequence_ops.cc
    out = _np.empty((max(n, 0),), dtype=dt)
    v = s
    for i in _np.arange(max(n, 0)):
        out[i] = v
        v = dt(v + d)
    return out


def meshgrid(*xs, **kw):
    return _np.meshgrid(*xs, indexing=kw.get("indexing", "xy"))


def concat(values, axis):
    return _wrap(_concat(values, axis))


def _concat(values, axis):
    vs = [_t(v) for v in values]
    first = next((v.dtype for v, raw in zip(vs, values) if isinstance(raw, _np.ndarray)), vs[0].dtype)
    # TF converts Python lists / scalars among the inputs to the dtype of the tensors they are concatenated with
    vs = [v if isinstance(raw, _np.ndarray) else v.astype(first) for v, raw in zip(vs, values)]
    return _np.concatenate(vs, axis=axis)


def stack(values, axis=0):
    return _np.stack([_t(v) for v in values], axis=axis)


def split(x, sizes, axis=0):
    idx = _np.cumsum([int(s) for s in sizes])[:-1]
    return _np.split(x, idx, axis=axis)


def where(c, a=None, b=None):
    return _np.where(c, _t(a), _t(b))


def zeros_like(x, dtype=None):
    return _np.zeros_like(x, dtype=dtype)


def zeros(s, dtype=float32):
    return _np.zeros(tuple(int(v) for v in _np.asarray(s).reshape(-1)), dtype=dtype)


def expand_dims(x, axis):
    return _np.expand_dims(x, axis)


def boolean_mask(x, m):
    m = _np.asarray(m).astype(_np.bool_)
    x = _np.asarray(x)
    return x.reshape((-1,) + x.shape[m.ndim:])[m.reshape(-1)]


def gather(x, idx, axis=0):
    return _np.take(_np.asarray(x), _np.asarray(idx).astype(_np.int64), axis=axis)


def argsort(x, direction="ASCENDING", axis=-1):
    x = _np.asarray(x)
    return (_np.argsort(-x if direction == "DESCENDING" else x, axis=axis, kind="stable")).astype(_np.int32)


def one_hot(idx, depth, dtype=float32):
    idx = _np.asarray(idx).astype(_np.int64)
    out = _np.zeros(idx.shape + (int(depth),), dtype=dtype)
    ok = (idx >= 0) & (idx < int(depth))
    out[ok, idx[ok]] = 1
    return out


def scatter_nd(indices, updates, shape):  # noqa: A002
    out = _np.zeros(tuple(int(v) for v in _np.asarray(shape).reshape(-1)), dtype=_np.asarray(updates).dtype)
    ind = _np.asarray(indices).astype(_np.int64)
    _np.add.at(out, tuple(ind[..., k] for k in _np.arange(ind.shape[-1])), updates)
    return out


def clip_by_value(x, lo, hi):
    return _np.minimum(_np.maximum(x, _t(lo)), _t(hi))


def equal(a, b):
    return _np.equal(a, b)


def argmax(x, axis=-1, output_type=_np.int64):
    return _np.argmax(x, axis=axis).astype(output_type)


def maximum(a, b):
    return _np.maximum(_t(a), _t(b))


def minimum(a, b):
    return _np.minimum(_t(a), _t(b))


def reduce_sum(x, axis=None, keepdims=False):
    return _wrap(_np.sum(x, axis=axis, keepdims=keepdims, dtype=_np.asarray(x).dtype))


# ---- transcendental back end --------------------------------------------------------------------------------
# FAKE_TF_MATH=libm (default): NumPy's float32 ufuncs (glibc / NumPy SIMD loops).
# FAKE_TF_MATH=torch: the same functions through PyTorch's CPU kernels (SLEEF-vectorised; torch.sigmoid is a fused
# kernel like Eigen's logistic) — an independent second implementation, used by tests/golden/tf_numerics_risk.py to
# measure how often a few-ulp difference in exp / sigmoid / atan / log flips a DISCRETE output (NMS order, ignore bit,
# target cell), since TensorFlow's own Eigen kernels cannot be run here.
import os as _os
_MATH = _os.environ.get("FAKE_TF_MATH", "libm")
if _MATH == "torch":
    import torch as _torch

    def _via_torch(fn):
        def run(x, *rest):
            a = _np.ascontiguousarray(_np.asarray(x, dtype=_F))
            with _np.errstate(all="ignore"):
                return fn(_torch.from_numpy(a.reshape(-1)), *rest).numpy().reshape(a.shape).astype(_F)
        return run
    _exp, _log, _log1p, _atan_f, _sigmoid_f = (_via_torch(_torch.exp), _via_torch(_torch.log), _via_torch(_torch.log1p),
                                               _via_torch(_torch.atan), _via_torch(_torch.sigmoid))

    def _pow(a, b):
        a, b = _np.broadcast_arrays(_np.asarray(a, dtype=_F), _np.asarray(b, dtype=_F))
        return _torch.pow(_torch.from_numpy(_np.ascontiguousarray(a)), _torch.from_numpy(_np.ascontiguousarray(b))).numpy().astype(_F)
else:
    def _exp(x):
        with _np.errstate(all="ignore"):
            return _np.exp(_np.asarray(x, dtype=_F)).astype(_F)

    def _log(x):
        with _np.errstate(all="ignore"):
            return _np.log(_np.asarray(x, dtype=_F)).astype(_F)

    def _log1p(x):
        with _np.errstate(all="ignore"):
            return _np.log1p(_np.asarray(x, dtype=_F)).astype(_F)

    def _atan_f(x):
        return _np.arctan(_np.asarray(x, dtype=_F)).astype(_F)

    def _sigmoid_f(x):
        x = _np.asarray(x, dtype=_F)
        with _np.errstate(all="ignore"):
            return (_F(1) / (_F(1) + _np.exp(-x))).astype(_F)

    def _pow(a, b):
        return _np.power(_t(a), _t(b))


def sigmoid(x):
    return _sigmoid_f(x)


def atan(x):
    return _atan_f(x)


def print(*a, **k):  # noqa: A001
    pass


def custom_gradient(fn):
    def wrapped(*args):
        out, grad = fn(*args)
        custom_gradient.last_grad = grad   # kept so that the golden generator can evaluate the hand-written gradient too
        return out
    return wrapped


def gradients(*a, **k):
    return None


def while_loop(cond, body, loop_vars, **kw):
    vs = list(loop_vars)
    while cond(*vs):
        vs = list(body(*vs))
    return vs


class TensorArray(object):
    def __init__(self, dtype, size=0, dynamic_size=False, **kw):
        self._dtype = dtype
        self._items = {}

    def write(self, idx, value):
        self._items[int(idx)] = _np.asarray(value, dtype=self._dtype)
        return self

    def stack(self):
        if not self._items:
            return _np.zeros((0,), dtype=self._dtype)
        return _np.stack([self._items[k] for k in sorted(self._items)], axis=0)


class TensorShape(tuple):
    pass


def _reduce_max(x, axis=None, keepdims=False):
    x = _np.asarray(x)
    if x.shape[axis if axis is not None else 0] == 0 and axis is not None:
        s = list(x.shape)
        del s[axis]
        return _np.full(s, _np.finfo(_F).min, dtype=x.dtype)
    return _np.max(x, axis=axis, keepdims=keepdims)


def _divide_no_nan(x, y):
    x, y = _np.broadcast_arrays(_t(x), _t(y))
    with _np.errstate(all="ignore"):
        return _np.where(y == 0, _np.zeros_like(x), x / _np.where(y == 0, _np.ones_like(y), y)).astype(x.dtype)


def _bce_logits(z, x):
    """max(x,0) - x z + log1p(exp(-|x|)), sigmoid_cross_entropy_with_logits"""
    z, x = _np.asarray(z, dtype=_F), _np.asarray(x, dtype=_F)
    with _np.errstate(all="ignore"):
        return (_np.maximum(x, _F(0)) - x * z + _log1p(_exp(-_np.abs(x)))).astype(_F)


def _mk(name, **attrs):
    m = _StubModule(name)
    for k, v in attrs.items():
        setattr(m, k, v)
    _sys.modules[name] = m
    return m


with _np.errstate(all="ignore"):
    pass

math = _mk(
    "tensorflow.math",
    maximum=maximum, minimum=minimum, square=lambda x: _np.square(_t(x)), reduce_sum=reduce_sum, reduce_max=_reduce_max,
    atan=atan, log=lambda x: _log(x), exp=lambda x: _exp(x),
    sigmoid=sigmoid, tanh=lambda x: _np.tanh(x), is_inf=_np.isinf, is_nan=_np.isnan, logical_and=_np.logical_and,
    logical_not=_np.logical_not, logical_or=_np.logical_or, not_equal=_np.not_equal, equal=_np.equal, greater=_np.greater,
    greater_equal=_np.greater_equal, less=_np.less, less_equal=_np.less_equal, floor=_np.floor,
    argmax=lambda x, axis=-1, output_type=_np.int64: _np.argmax(x, axis=axis).astype(output_type), divide_no_nan=_divide_no_nan,
    abs=_np.abs, sqrt=_np.sqrt, pow=lambda a, b: _pow(a, b))
reduce_max = _reduce_max
linalg = _mk("tensorflow.linalg", norm=lambda x, axis=None: _np.sqrt(_np.sum(_np.square(x), axis=axis, dtype=_np.asarray(x).dtype)))
nn = _mk("tensorflow.nn", sigmoid_cross_entropy_with_logits=lambda labels=None, logits=None: _bce_logits(labels, logits))
autograph = _mk("tensorflow.autograph", experimental=_mk("tensorflow.autograph.experimental", do_not_convert=lambda f=None, **k: f if f else (lambda g: g)))


class _Loss(object):
    """tf.keras.losses.Loss with the default SUM_OVER_BATCH_SIZE reduction: __call__ = mean over all elements of call()."""
    def __init__(self, *a, **k):
        pass

    def __call__(self, y_true, y_pred):
        v = _np.asarray(self.call(y_true, y_pred))
        return _wrap(_np.sum(v, dtype=_F) / _F(v.size) if v.ndim else v)


class _Huber(object):
    def __init__(self, delta=1.0, reduction=None, **k):
        self.delta = _F(delta)

    def __call__(self, y_true, y_pred):  # Reduction.NONE: mean over the last axis only
        e = _np.asarray(y_pred, dtype=_F) - _np.asarray(y_true, dtype=_F)
        a = _np.abs(e)
        v = _np.where(a <= self.delta, _F(0.5) * e * e, self.delta * a - _F(0.5) * self.delta * self.delta).astype(_F)
        return _np.mean(v, axis=-1, dtype=_F)


def _keras_bce(y_true, y_pred, from_logits=False):
    assert from_logits
    return _np.mean(_bce_logits(y_true, y_pred), axis=-1, dtype=_F)


_backend = _mk(
    "tensorflow.keras.backend",
    dtype=lambda x: _np.asarray(x).dtype, cast=lambda x, d: _wrap(_np.asarray(x).astype(d)), shape=shape, reshape=reshape,
    arange=lambda start, stop=None, step=1, dtype="int32": (_np.arange(int(start), dtype=dtype) if stop is None
                                                            else _np.arange(int(start), int(stop), int(step), dtype=dtype)),
    tile=lambda x, n: _np.tile(x, [int(v) for v in n]),
    concatenate=lambda xs, axis=-1: _np.concatenate(xs, axis=axis), sigmoid=sigmoid, exp=math.exp, log=math.log,
    maximum=maximum, minimum=minimum, expand_dims=lambda x, axis=-1: _np.expand_dims(x, axis),
    max=lambda x, axis=None, keepdims=False: _reduce_max(x, axis=axis, keepdims=keepdims), sum=lambda x, axis=None: reduce_sum(x, axis=axis),
    binary_crossentropy=lambda t, o, from_logits=False: _bce_logits(t, o) if from_logits else None,
    switch=lambda c, a, b: _np.where(c, a, b), zeros_like=lambda x: _np.zeros_like(x), constant=lambda x, dtype=None: _t(x, dtype),
    square=lambda x: _np.square(x))
keras = _mk(
    "tensorflow.keras",
    backend=_backend,
    losses=_mk("tensorflow.keras.losses", Loss=_Loss, Huber=_Huber, binary_crossentropy=_keras_bce,
               Reduction=_types.SimpleNamespace(NONE="none", SUM="sum", SUM_OVER_BATCH_SIZE="sum_over_batch_size")),
    layers=_mk("tensorflow.keras.layers", Layer=_Stub),
    activations=_mk("tensorflow.keras.activations"), regularizers=_mk("tensorflow.keras.regularizers"),
    initializers=_mk("tensorflow.keras.initializers"), Model=_Stub, Sequential=_Stub)
image = _mk("tensorflow.image")
_rng = _np.random.default_rng(12345)
random = _mk("tensorflow.random", uniform=lambda shape, minval=0, maxval=1, dtype=float32, **k: _wrap(
    (_rng.random(tuple(int(v) for v in shape), dtype=_np.float32) * _F(maxval - minval) + _F(minval)).astype(dtype)))
data = _mk("tensorflow.data")


def __getattr__(name):
    if name.startswith("__"):
        raise AttributeError(name)
    return _Stub
