"""A numpy stand-in for the slice of TensorFlow 1.8 that the reference's hot-path files use.

Purpose: TensorFlow cannot be installed here, so the reference's Python sources cannot run as they
are. This module is registered as `tensorflow` (see `install()`), after which
/root/reference/deepbelief/layers/{rbm,bgrbm,gbrbm,ggrbm,gplvm,gprbm}.py, dist.py and preprocessing.py
import and execute UNMODIFIED, eagerly, on numpy arrays: every `tf.*` call computes its documented
result immediately and `Session.run(x)` returns x. Random ops pop arrays from `RANDOM_QUEUE` (pushed
by the golden generator in call order) so the draws are the injected ones.

What this is NOT: TensorFlow's kernels (Eigen GEMM rounding, Philox streams, autodiff). Gradients are
therefore pinned by finite differences of the reference's own loss functions, not by TF autodiff.
Used only by tests/golden/make_golden.py (run in the build container, where /root/reference exists).
"""
import contextlib
import sys
import types

import numpy as np
import scipy.linalg

RANDOM_QUEUE = []          # arrays consumed by random_uniform / random_normal, in call order
TRUNC_QUEUE = []           # arrays consumed by truncated_normal (initial W, rbm.py:59-62), in call order
ZEROS_INJECT = {}          # shape -> list of arrays handed out by tf.zeros(shape) (initial b / c, rbm.py:64-72)
_RNG = np.random.RandomState(0)


def _dt(dtype):
    if dtype is None:
        return None
    if isinstance(dtype, str):
        return np.dtype(dtype)
    return np.dtype(dtype)


def _arr(x, dtype=None):
    return np.asarray(x, dtype=_dt(dtype))


class _Var(np.ndarray):
    """ndarray carrying the TF variable name (layer.py:13-15 filters on it)."""
    name = ''

    def assign(self, value, name=None):        # returns the "op"; nothing is executed eagerly
        return value


def _make_var(value, name):
    v = np.array(value).view(_Var)
    v.name = name or 'Variable'
    return v


_SCOPE = []
_VARS = []


@contextlib.contextmanager
def variable_scope(name, *a, **k):
    _SCOPE.append(name)
    try:
        yield
    finally:
        _SCOPE.pop()


def get_variable(name=None, initializer=None, shape=None, dtype=None, trainable=True, **k):
    if initializer is None:
        initializer = np.zeros(shape, dtype=_dt(dtype) or np.float32)
    if callable(initializer):
        initializer = initializer()
    v = _make_var(_arr(initializer, dtype), '/'.join(_SCOPE + [name or 'var']) + ':0')
    v.trainable = trainable
    _VARS.append(v)
    return v


def Variable(initial_value, dtype=None, name=None, trainable=True, **k):
    return get_variable(name=name, initializer=initial_value, dtype=dtype, trainable=trainable)


def global_variables():
    return []          # nothing is session-owned here; Stateful then skips its Saver (layer.py:21-22)


def trainable_variables():
    return [v for v in _VARS if getattr(v, 'trainable', True)]


def constant(value, dtype=None, shape=None, name=None):
    return _arr(value, dtype)


def placeholder(dtype=None, shape=None, name=None):
    return None


def _take_random(shape, dtype, kind):
    shape = tuple(int(s) for s in np.atleast_1d(shape))
    if RANDOM_QUEUE:
        a = np.asarray(RANDOM_QUEUE.pop(0))
        assert a.shape == shape, 'injected %s draw has shape %s, op wants %s' % (kind, a.shape, shape)
        return a.astype(_dt(dtype) or np.float32)
    if kind == 'uniform':
        return _RNG.uniform(size=shape).astype(_dt(dtype) or np.float32)
    return _RNG.normal(size=shape).astype(_dt(dtype) or np.float32)


def random_uniform(shape, minval=0.0, maxval=1.0, dtype='float32', seed=None, name=None):
    return _take_random(shape, dtype, 'uniform') * (maxval - minval) + minval


def random_normal(shape, mean=0.0, stddev=1.0, dtype='float32', seed=None, name=None):
    return _take_random(shape, dtype, 'normal') * stddev + mean


def truncated_normal(shape, mean=0.0, stddev=1.0, dtype='float32', seed=None, name=None):
    shape = tuple(int(s) for s in shape)
    if TRUNC_QUEUE:            # injected initial weights (used as they are: parity runs choose W freely)
        a = np.asarray(TRUNC_QUEUE.pop(0))
        assert a.shape == shape, 'injected W has shape %s, op wants %s' % (a.shape, shape)
        return a.astype(_dt(dtype))
    x = _RNG.normal(size=shape)
    bad = np.abs(x) > 2.0
    while bad.any():
        x[bad] = _RNG.normal(size=int(bad.sum()))
        bad = np.abs(x) > 2.0
    return (x * stddev + mean).astype(_dt(dtype))


def zeros(shape, dtype='float32', name=None):
    shape = tuple(int(s) for s in np.atleast_1d(shape))
    pending = ZEROS_INJECT.get(shape)
    if pending:
        return np.asarray(pending.pop(0)).reshape(shape).astype(_dt(dtype))
    return np.zeros(shape, dtype=_dt(dtype))


def ones(shape, dtype='float32', name=None):
    return np.ones(tuple(int(s) for s in np.atleast_1d(shape)), dtype=_dt(dtype))


def matmul(a, b, transpose_a=False, transpose_b=False, name=None):
    a = np.asarray(a)
    b = np.asarray(b)
    return (a.T if transpose_a else a) @ (b.T if transpose_b else b)


def _stable_sigmoid(x):
    x = np.asarray(x)
    out = np.empty_like(x)
    pos = x >= 0
    out[pos] = 1.0 / (1.0 + np.exp(-x[pos]))
    ex = np.exp(x[~pos])
    out[~pos] = ex / (1.0 + ex)
    return out


def _softplus(x):
    x = np.asarray(x)
    return np.maximum(x, 0) + np.log1p(np.exp(-np.abs(x)))


def _sigmoid_ce(logits=None, labels=None, name=None):
    x, z = np.asarray(logits), np.asarray(labels)
    return np.maximum(x, 0) - x * z + np.log1p(np.exp(-np.abs(x)))


def _moments(x, axes, keep_dims=False, name=None):
    ax = tuple(axes)
    return np.mean(x, axis=ax, keepdims=keep_dims), np.var(x, axis=ax, keepdims=keep_dims)


def reduce_sum(x, axis=None, keepdims=False, keep_dims=None, name=None):
    return np.sum(x, axis=axis, keepdims=bool(keepdims or keep_dims))


def reduce_mean(x, axis=None, keepdims=False, keep_dims=None, name=None):
    return np.mean(x, axis=axis, keepdims=bool(keepdims or keep_dims))


def cast(x, dtype=None, name=None):
    return np.asarray(x).astype(_dt(dtype))


def cholesky(a, name=None):
    return np.linalg.cholesky(a)


def matrix_triangular_solve(matrix, rhs, lower=True, adjoint=False, name=None):
    return scipy.linalg.solve_triangular(matrix, rhs, lower=lower, trans='T' if adjoint else 'N')


def diag(x, name=None):
    return np.diag(x)


def diag_part(x, name=None):
    return np.diagonal(x).copy()


class InvalidArgumentError(Exception):
    pass


class FailedPreconditionError(Exception):
    pass


def assert_non_negative(x, name=None):
    if np.any(np.asarray(x) < 0):
        raise InvalidArgumentError('assert_non_negative failed')
    return None


@contextlib.contextmanager
def control_dependencies(deps):
    yield


def cond(pred, true_fn, false_fn, name=None):
    return true_fn() if bool(pred) else false_fn()


def assign(ref, value, name=None):
    return value


class Session:
    def __init__(self, *a, **k):
        self.graph = types.SimpleNamespace(finalize=lambda: None)

    def run(self, fetches, feed_dict=None):
        return fetches

    def close(self):
        pass


def install():
    """Register this module as `tensorflow` and return it."""
    tf = types.ModuleType('tensorflow')
    g = globals()
    for name in ('variable_scope', 'get_variable', 'Variable', 'global_variables', 'trainable_variables', 'constant',
                 'placeholder', 'random_uniform', 'random_normal', 'truncated_normal', 'zeros', 'ones', 'matmul',
                 'reduce_sum', 'reduce_mean', 'cast', 'cholesky', 'matrix_triangular_solve', 'diag', 'diag_part',
                 'assert_non_negative', 'control_dependencies', 'cond', 'assign', 'Session'):
        setattr(tf, name, g[name])
    tf.add = lambda a, b, name=None: np.add(a, b)
    tf.subtract = lambda a, b, name=None: np.subtract(a, b)
    tf.multiply = lambda a, b, name=None: np.multiply(a, b)
    tf.div = lambda a, b, name=None: np.divide(a, b)
    tf.negative = lambda a, name=None: np.negative(a)
    tf.square = lambda a, name=None: np.square(a)
    tf.sqrt = lambda a, name=None: np.sqrt(a)
    tf.log = lambda a, name=None: np.log(a)
    tf.exp = lambda a, name=None: np.exp(a)
    tf.tanh = lambda a, name=None: np.tanh(a)
    tf.sigmoid = lambda a, name=None: _stable_sigmoid(a)
    tf.transpose = lambda a, perm=None, name=None: np.transpose(a, perm)
    tf.greater = lambda a, b, name=None: np.greater(a, b)
    tf.logical_not = lambda a, name=None: np.logical_not(a)
    tf.identity = lambda a, name=None: a
    tf.stop_gradient = lambda a, name=None: a
    tf.shape = lambda a, name=None: np.array(np.shape(a))
    tf.reshape = lambda a, shape, name=None: np.reshape(a, shape)
    tf.expand_dims = lambda a, axis, name=None: np.expand_dims(a, axis)
    tf.tile = lambda a, m, name=None: np.tile(a, m)
    tf.concat = lambda v, axis, name=None: np.concatenate(v, axis=axis)
    tf.clip_by_value = lambda a, lo, hi, name=None: np.clip(a, lo, hi)

    def _slice(a, begin, size, name=None):          # size -1 = to the end of that axis
        a = np.asarray(a)
        idx = tuple(slice(int(b), None if int(sz) == -1 else int(b) + int(sz)) for b, sz in zip(begin, size))
        return a[idx]
    tf.slice = _slice
    tf.pad = lambda a, paddings, mode='CONSTANT', name=None: np.pad(a, np.asarray(paddings), mode='constant')
    tf.split = lambda a, n, axis=0, name=None: np.split(a, n, axis=axis)
    tf.reduce_min = lambda a, axis=None, name=None: np.min(a, axis=axis)
    tf.reduce_max = lambda a, axis=None, name=None: np.max(a, axis=axis)
    tf.assert_type = lambda a, t, name=None: None
    tf.Print = lambda a, data, name=None: a
    tf.is_variable_initialized = lambda v: True
    tf.assert_variables_initialized = lambda v=None: None
    tf.variables_initializer = lambda v, name=None: None
    tf.global_variables_initializer = lambda: None
    tf.get_default_session = lambda: None
    tf.int32 = np.int32
    tf.float32 = np.float32
    tf.float64 = np.float64
    tf.nn = types.SimpleNamespace(softplus=lambda a, name=None: _softplus(a), moments=_moments,
                                  sigmoid_cross_entropy_with_logits=_sigmoid_ce)
    tf.summary = types.SimpleNamespace(histogram=lambda *a, **k: None, scalar=lambda *a, **k: None,
                                       tensor=lambda *a, **k: None, merge_all=lambda: None,
                                       FileWriter=lambda *a, **k: None)
    tf.errors = types.SimpleNamespace(FailedPreconditionError=FailedPreconditionError,
                                      InvalidArgumentError=InvalidArgumentError)
    tf.train = types.SimpleNamespace(Optimizer=types.SimpleNamespace(GATE_GRAPH=2), Saver=lambda *a, **k: None,
                                     AdamOptimizer=lambda *a, **k: None, GradientDescentOptimizer=lambda *a, **k: None)
    sys.modules['tensorflow'] = tf
    return tf
