"""Minimal `tensorflow` stand-in on torch-CPU: just the ops the reference's decomp stage calls (see ../README.md).

TEST INFRASTRUCTURE ONLY.  Semantics follow the TF 2.4 documentation of each op; `tf.Tensor` is `torch.Tensor`.
"""
import builtins as _b
import types as _types

import numpy as _np
import torch as _torch

Tensor = _torch.Tensor
_FLOAT = _torch.float32


def set_float(dtype):
    """dtype that `tf.float32` resolves to (float64 = high-precision run of the same reference code)."""
    global _FLOAT
    _FLOAT = dtype
    _torch.set_default_dtype(dtype)


class _FloatType:
    """`tf.float32`: resolved late so that set_float() affects modules that were imported earlier."""

    def resolve(self):
        return _FLOAT


float32 = _FloatType()
DType = object          # annotations only
float64 = _torch.float64
int32 = _torch.int32
int64 = _torch.int64
bool = _torch.bool   # noqa: A001


def _dt(dtype):
    if dtype is None:
        return None
    return dtype.resolve() if isinstance(dtype, _FloatType) else dtype


def _t(x, dtype=None):
    as_f32 = isinstance(dtype, _FloatType)
    dtype = _dt(dtype)
    if isinstance(x, _torch.Tensor):
        return x if dtype is None or x.dtype == dtype else x.to(dtype)
    a = _np.asarray(x)
    if as_f32 and a.dtype.kind == 'f':
        a = a.astype(_np.float32)           # dtype=tf.float32: the VALUES are float32-representable, whatever set_float() says
    if dtype is None and a.dtype.kind == 'f':
        dtype = _FLOAT                      # python floats / float64 numpy arrays become the working float type
    return _torch.as_tensor(a.copy() if not a.flags.writeable else a, dtype=dtype)


class Variable(_torch.Tensor):
    """tf.Variable: a leaf tensor with .assign(); trainable ones require grad (torch autograd = GradientTape)."""

    @staticmethod
    def __new__(cls, initial_value, trainable=True, dtype=None, name=None):
        data = _t(initial_value, dtype).detach().clone()
        return _torch.Tensor._make_subclass(cls, data, _b.bool(trainable) and data.is_floating_point())

    def assign(self, value):
        with _torch.no_grad():
            self.copy_(_t(value, self.dtype))
        return self

    def assign_add(self, value):
        with _torch.no_grad():
            self.add_(_t(value, self.dtype))
        return self

    def assign_sub(self, value):
        with _torch.no_grad():
            self.sub_(_t(value, self.dtype))
        return self

    def read_value(self):
        return self.detach().clone()

    @classmethod
    def __torch_function__(cls, func, types, args=(), kwargs=None):
        # results of ops on a Variable are plain tensors
        with _torch._C.DisableTorchFunctionSubclass():
            out = func(*args, **(kwargs or {}))
        return out


def convert_to_tensor(value, dtype=None, name=None):
    return _t(value, dtype)


def constant(value, dtype=None, shape=None, name=None):
    return _t(value, dtype)


def identity(x, name=None):
    return x


def stop_gradient(x):
    return x.detach()


def cast(x, dtype):
    return _t(x).to(_dt(dtype))


def shape(x):
    return tuple(_t(x).shape)


def zeros(shape, dtype=float32):
    return _torch.zeros(tuple(int(s) for s in shape), dtype=_dt(dtype))


def ones(shape, dtype=float32):
    return _torch.ones(tuple(int(s) for s in shape), dtype=_dt(dtype))


def zeros_like(x, dtype=None):
    return _torch.zeros_like(_t(x), dtype=_dt(dtype))


def ones_like(x, dtype=None):
    return _torch.ones_like(_t(x), dtype=_dt(dtype))


def eye(n, dtype=float32):
    return _torch.eye(int(n), dtype=_dt(dtype))


def range(start, limit=None, delta=1, dtype=None):   # noqa: A001
    if limit is None:
        start, limit = 0, start
    return _torch.arange(int(start), int(limit), int(delta), dtype=_dt(dtype) or _torch.int64)


def linspace(start, stop, num):
    # tf.linspace(0., stop, n) on float32: start + i * (stop - start) / (n - 1)
    n = int(num)
    if n == 1:
        return _torch.tensor([float(start)], dtype=_FLOAT)
    i = _torch.arange(n, dtype=_FLOAT)
    return float(start) + i * ((float(stop) - float(start)) / (n - 1))


def reshape(x, shape):
    return _torch.reshape(_t(x), tuple(int(s) for s in shape))


def transpose(x, perm=None):
    x = _t(x)
    if perm is None:
        perm = list(reversed(_b.range(x.dim())))
    return x.permute(*perm)


def concat(values, axis):
    return _torch.cat([_t(v) for v in values], dim=axis)


def stack(values, axis=0):
    return _torch.stack([_t(v) for v in values], dim=axis)


def tile(x, multiples):
    return _t(x).repeat(*[int(m) for m in multiples])


def repeat(x, repeats, axis=None):
    return _torch.repeat_interleave(_t(x), int(repeats), dim=axis)


def broadcast_to(x, shape):
    return _torch.broadcast_to(_t(x), tuple(int(s) for s in shape))


def where(condition, x=None, y=None):
    if x is None and y is None:
        return _torch.nonzero(condition)            # [M, rank] int64, row-major order (as tf.where)
    return _torch.where(condition, x, y)


def boolean_mask(tensor, mask):
    return _t(tensor)[mask]


def scatter_nd(indices, updates, shape):
    """Scatter `updates` into zeros(shape) at first-axis indices [M,1] (duplicates would sum, as in TF)."""
    assert indices.dim() == 2 and indices.shape[1] == 1
    out = _torch.zeros(tuple(int(s) for s in shape), dtype=updates.dtype)
    return out.index_add(0, indices[:, 0], updates)


def tensor_scatter_nd_update(tensor, indices, updates):
    assert indices.dim() == 2 and indices.shape[1] == 1
    out = tensor.clone()
    return out.index_copy(0, indices[:, 0], updates.to(out.dtype))


def gather(params, indices, axis=0, batch_dims=0):
    if batch_dims == 0:
        return _torch.index_select(params, axis, indices.reshape(-1)).reshape(
            params.shape[:axis] + indices.shape + params.shape[axis + 1:])
    raise NotImplementedError('tf.gather with batch_dims')


def einsum(eq, *ops):
    return _torch.einsum(eq, *ops)


def matmul(a, b, transpose_a=False, transpose_b=False):
    a = a.transpose(-1, -2) if transpose_a else a
    b = b.transpose(-1, -2) if transpose_b else b
    return a @ b


def _reduce(fn, x, axis, keepdims):
    x = _t(x)
    if axis is None:
        return fn(x)
    return fn(x, dim=axis, keepdim=keepdims)


def reduce_sum(x, axis=None, keepdims=False):
    return _reduce(_torch.sum, x, axis, keepdims)


def reduce_mean(x, axis=None, keepdims=False):
    return _reduce(_torch.mean, x, axis, keepdims)


def reduce_max(x, axis=None, keepdims=False):
    x = _t(x)
    return x.max() if axis is None else x.max(dim=axis, keepdim=keepdims).values


def reduce_min(x, axis=None, keepdims=False):
    x = _t(x)
    return x.min() if axis is None else x.min(dim=axis, keepdim=keepdims).values


def square(x):
    return x * x


def sqrt(x):
    return _torch.sqrt(x)


def abs(x):   # noqa: A001
    return _torch.abs(x)


def exp(x):
    return _torch.exp(x)


def argmax(x, axis=None):
    return _torch.argmax(x, dim=axis)               # first index among ties, int64 (tf.argmax default)


def one_hot(indices, depth, dtype=float32):
    return _torch.nn.functional.one_hot(indices, int(depth)).to(_dt(dtype))


def clip_by_value(t, clip_value_min, clip_value_max, name=None):
    return _torch.clamp(t, min=clip_value_min, max=clip_value_max)


def maximum(a, b):
    return _torch.maximum(_t(a), _t(b)) if isinstance(a, _torch.Tensor) or isinstance(b, _torch.Tensor) else _b.max(a, b)


def minimum(a, b):
    return _torch.minimum(_t(a), _t(b)) if isinstance(a, _torch.Tensor) or isinstance(b, _torch.Tensor) else _b.min(a, b)


def cumsum(x, axis=0):
    return _torch.cumsum(x, dim=axis)


def atan2(y, x):
    return _torch.atan2(y, x)


def acos(x):
    return _torch.acos(x)


def searchsorted(sorted_sequence, values, side='left'):
    return _torch.searchsorted(sorted_sequence, values, right=(side == 'right'))


def custom_gradient(f):
    def wrapped(*a, **k):
        return f(*a, **k)[0]
    return wrapped


def function(f=None, **_k):
    return f if f is not None else (lambda g: g)


# ---------------------------------------------------------------- tf.math
math = _types.ModuleType('tensorflow.math')
math.sin = _torch.sin
math.cos = _torch.cos
math.log = _torch.log
math.exp = _torch.exp
math.sqrt = _torch.sqrt
math.abs = _torch.abs
math.square = square
math.pow = lambda x, y: _torch.pow(x, y)
math.minimum = minimum
math.maximum = maximum
math.reduce_sum = reduce_sum
math.reduce_mean = reduce_mean
math.reduce_max = reduce_max
math.reduce_min = reduce_min


def _divide_no_nan(x, y):
    """tf.math.divide_no_nan: 0 wherever the denominator is 0 (value AND gradient)."""
    x, y = _torch.broadcast_tensors(_t(x), _t(y))
    zero = y == 0
    q = x / _torch.where(zero, _torch.ones_like(y), y)
    return _torch.where(zero, _torch.zeros_like(q), q)


def _cumprod(x, axis=0, exclusive=False):
    p = _torch.cumprod(x, dim=axis)
    if exclusive:
        p = _torch.cat([_torch.ones_like(p.narrow(axis, 0, 1)), p.narrow(axis, 0, p.shape[axis] - 1)], dim=axis)
    return p


math.divide_no_nan = _divide_no_nan
math.cumprod = _cumprod

# ---------------------------------------------------------------- tf.linalg / tf.nn
linalg = _types.ModuleType('tensorflow.linalg')


def _l2_normalize(x, axis=None, epsilon=1e-12):
    """tf.linalg.l2_normalize: x * rsqrt(max(sum(x**2, axis), epsilon)) -- epsilon bounds the SQUARED norm."""
    sq = _torch.sum(x * x, dim=axis, keepdim=True)
    return x * _torch.rsqrt(_torch.clamp(sq, min=epsilon))


linalg.l2_normalize = _l2_normalize
linalg.norm = lambda x, axis=None: _torch.linalg.vector_norm(x if isinstance(x, _torch.Tensor) else _torch.as_tensor(_np.asarray(x)), dim=axis)
nn = _types.ModuleType('tensorflow.nn')
nn.embedding_lookup = lambda params, ids: params[ids]
nn.l2_normalize = _l2_normalize

# ---------------------------------------------------------------- tf.random
random = _types.ModuleType('tensorflow.random')
random.queue = []        # tests push the exact `roll` tensors the reference code should draw


def _uniform(shape, minval=0., maxval=1., dtype=float32, seed=None):
    if random.queue:
        v = _t(random.queue.pop(0), dtype)
        assert tuple(v.shape) == tuple(int(s) for s in shape), (v.shape, shape)
        return v
    return _torch.rand(tuple(int(s) for s in shape), dtype=_dt(dtype)) * (maxval - minval) + minval


random.uniform = _uniform
random.normal = lambda shape, mean=0., stddev=1., dtype=float32: _torch.randn(tuple(shape), dtype=_dt(dtype)) * stddev + mean
random.set_seed = lambda seed: _torch.manual_seed(int(seed))

# ---------------------------------------------------------------- tf.debugging
debugging = _types.ModuleType('tensorflow.debugging')


class InvalidArgumentError(Exception):
    pass


errors = _types.ModuleType('tensorflow.errors')
errors.InvalidArgumentError = InvalidArgumentError


def _check_numerics(tensor, message):
    if not _torch.isfinite(tensor).all():
        raise InvalidArgumentError(message + ' : Tensor had NaN or Inf values')
    return tensor


def _assert_greater(x, y, message=None):
    if not (x > y).all():
        raise InvalidArgumentError(message or 'assert_greater failed')


debugging.check_numerics = _check_numerics
debugging.assert_greater = _assert_greater

# ---------------------------------------------------------------- tf.keras
keras = _types.ModuleType('tensorflow.keras')
keras.layers = _types.ModuleType('tensorflow.keras.layers')
keras.losses = _types.ModuleType('tensorflow.keras.losses')


class _KerasModel:
    """tf.keras.Model as the reference uses it ("only the parent's trackability"): __call__ -> call."""

    def __init__(self, *a, **k):
        pass

    def __call__(self, *a, **k):
        return self.call(*a, **k)


class _Activation:
    def __init__(self, name):
        self.name = name

    def __call__(self, x):
        if self.name == 'relu':
            return _torch.relu(x)
        if self.name == 'sigmoid':
            return _torch.sigmoid(x)
        raise NotImplementedError(self.name)


class _Dense:
    """tf.keras.layers.Dense: activation(x @ kernel + bias); kernel [in, units] glorot-uniform, bias zeros."""

    def __init__(self, units, activation=None):
        self.units = int(units)
        self.activation = activation
        self.kernel = None
        self.bias = None
        self.trainable = True
        self.built = False

    def build(self, in_dim):
        limit = (6.0 / (in_dim + self.units)) ** 0.5
        self.kernel = Variable((_torch.rand(in_dim, self.units, dtype=_FLOAT) * 2 - 1) * limit)
        self.bias = Variable(_torch.zeros(self.units, dtype=_FLOAT))
        self.built = True

    def set_weights(self, weights):
        self.kernel = Variable(_t(weights[0], float32))
        self.bias = Variable(_t(weights[1], float32))
        self.built = True

    @property
    def weights(self):
        return [self.kernel, self.bias]

    trainable_variables = weights

    def __call__(self, x):
        if not self.built:
            self.build(x.shape[-1])
        y = x @ self.kernel + self.bias
        return y if self.activation is None else self.activation(y)


keras.Model = _KerasModel
keras.layers.Dense = _Dense
keras.layers.Activation = _Activation
keras.losses.MSE = lambda y_true, y_pred: _torch.mean((y_pred - y_true) ** 2, dim=-1)


# ---------------------------------------------------------------- numpy interop
def _patch_numpy_interop():
    """`tf_tensor (op) ndarray` is a tensor op in TF (vq_layers.py:318-319 multiplies by `used.numpy()`); torch would
    route it through ndarray.__array_wrap__ and drop autograd.  Convert ndarray operands to tensors first."""
    if getattr(_torch.Tensor, '_vqn_np_patched', False):
        return

    def wrap(name):
        orig = getattr(_torch.Tensor, name)

        def op(self, other):
            if isinstance(other, _np.ndarray):
                other = _torch.as_tensor(other).to(self.dtype if other.dtype.kind == 'f' else None)
            return orig(self, other)
        setattr(_torch.Tensor, name, op)

    for nm in ('__mul__', '__rmul__', '__add__', '__radd__', '__sub__', '__rsub__', '__truediv__', '__rtruediv__'):
        wrap(nm)
    # TF tensors are immutable: `loss += term` REBINDS (vq_nfr.py:925-933 keeps loss_dict['rgb'] intact while the running
    # sum grows); torch would add in place and alias the two
    _torch.Tensor.__iadd__ = lambda self, other: self + other
    _torch.Tensor.__isub__ = lambda self, other: self - other
    _torch.Tensor.__imul__ = lambda self, other: self * other
    _torch.Tensor.__itruediv__ = lambda self, other: self / other
    _torch.Tensor._vqn_np_patched = True


_patch_numpy_interop()
