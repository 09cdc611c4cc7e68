"""Stand-ins for the handful of `tf.*` objects the reference's example scripts construct themselves
(SURVEY.md §8b "shim surface"): `tf.Session`, `tf.ConfigProto`, `tf.GPUOptions`,
`tf.train.{Adam,GradientDescent,Momentum}Optimizer`, `tf.summary.FileWriter`, `tf.get_variable`,
`tf.random_normal`. A script ports by replacing `import tensorflow as tf` with
`from deepbelief_b200 import compat as tf`.

Semantics shift: the reference's layer methods return symbolic tensors evaluated by `session.run`;
here they return CUDA tensors eagerly (or `Lazy` values for attributes such as `GPLVM.loss` that the
reference re-evaluates on every run), and `Session.run(x)` returns numpy arrays of x.
"""
import types

import numpy as np
import torch

from ._lib import db_opt_t, OPT_SGD, OPT_MOMENTUM, OPT_ADAM


class Lazy:
    """A value re-computed at every `Session.run` — the eager counterpart of a TF graph tensor."""

    def __init__(self, fn):
        self._fn = fn

    def eval(self):
        return self._fn()


def _to_numpy(x):
    if isinstance(x, Lazy):
        x = x.eval()
    if isinstance(x, torch.Tensor):
        return x.detach().cpu().numpy()
    return x


class Session:
    """`run` returns numpy values of eager tensors; `feed_dict` is accepted for Placeholder objects."""

    def __init__(self, config=None, graph=None):
        self.graph = types.SimpleNamespace(finalize=lambda: None)

    def run(self, fetches, feed_dict=None):
        if feed_dict:
            for ph, val in feed_dict.items():
                if isinstance(ph, Placeholder):
                    ph.value = val
        if isinstance(fetches, (list, tuple)):
            return [_to_numpy(f) for f in fetches]
        return _to_numpy(fetches)

    def close(self):
        pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        return False


class Placeholder:
    def __init__(self, dtype=None, shape=None, name=None):
        self.shape, self.name, self.value = shape, name, None


def placeholder(dtype=None, shape=None, name=None):
    return Placeholder(dtype, shape, name)


class ConfigProto:
    def __init__(self, **kw):
        self.__dict__.update(kw)


class GPUOptions:
    def __init__(self, **kw):
        self.__dict__.update(kw)


class Variable:
    """A named CUDA tensor (what `tf.get_variable` returns to the scripts, e.g. `x_test`)."""

    def __init__(self, value, name='Variable'):
        self.name = name
        self.value = torch.as_tensor(np.asarray(value, dtype=np.float32)).cuda().contiguous()

    def assign(self, value):
        self.value.copy_(torch.as_tensor(np.asarray(_to_numpy(value), dtype=np.float32)).reshape(self.value.shape))
        return self


def get_variable(name=None, initializer=None, shape=None, dtype=None, trainable=True):
    if initializer is None:
        initializer = np.zeros(shape, dtype=np.float32)
    return Variable(_to_numpy(initializer), name=name or 'Variable')


def random_normal(shape, mean=0.0, stddev=1.0, dtype=None, seed=None):
    return np.random.normal(mean, stddev, size=tuple(shape)).astype(np.float32)


def reset_default_graph():
    pass


# ---- optimizers: hyper-parameter holders; the update itself is the CUDA kernel db_optimizer_step --------------
class Optimizer:
    GATE_GRAPH = 2
    kind = OPT_SGD

    def __init__(self, learning_rate, weight_decay=0.0):
        self.learning_rate = float(learning_rate)
        self.weight_decay = float(weight_decay)
        self.step_count = 0

    def state_multiple(self):
        return 0

    def reset(self):
        """Slots and beta powers are re-initialised at the start of every train()/optimize()
        (rbm.py:107-117, gplvm.py:386-396)."""
        self.step_count = 0

    def descriptor(self, lr=None):
        self.step_count += 1
        o = db_opt_t()
        o.kind = self.kind
        o.lr = self.learning_rate if lr is None else float(lr)
        o.momentum = getattr(self, 'momentum', 0.0)
        o.beta1 = getattr(self, 'beta1', 0.9)
        o.beta2 = getattr(self, 'beta2', 0.999)
        o.eps = getattr(self, 'epsilon', 1e-8)
        o.weight_decay = self.weight_decay
        o.step = self.step_count
        clock = getattr(self, 'clock', None)            # device clock while the step is captured into a CUDA graph
        o.clock = clock.data_ptr() if clock is not None else None
        return o


class GradientDescentOptimizer(Optimizer):
    kind = OPT_SGD


class MomentumOptimizer(Optimizer):
    kind = OPT_MOMENTUM

    def __init__(self, learning_rate, momentum=0.9, weight_decay=0.0):
        super().__init__(learning_rate, weight_decay)
        self.momentum = float(momentum)

    def state_multiple(self):
        return 1


class AdamOptimizer(Optimizer):
    kind = OPT_ADAM

    def __init__(self, learning_rate=0.001, beta1=0.9, beta2=0.999, epsilon=1e-8, weight_decay=0.0):
        super().__init__(learning_rate, weight_decay)
        self.beta1, self.beta2, self.epsilon = float(beta1), float(beta2), float(epsilon)

    def state_multiple(self):
        return 2


class _FileWriter:
    """TensorBoard summaries are declared but only evaluated when a writer is passed;
    accepted and ignored here."""

    def __init__(self, logdir=None, graph=None):
        self.logdir = logdir

    def add_summary(self, summary, step=None):
        pass

    def close(self):
        pass


train = types.SimpleNamespace(Optimizer=Optimizer, GradientDescentOptimizer=GradientDescentOptimizer,
                              MomentumOptimizer=MomentumOptimizer, AdamOptimizer=AdamOptimizer)
summary = types.SimpleNamespace(FileWriter=_FileWriter, merge_all=lambda: None, scalar=lambda *a, **k: None,
                                histogram=lambda *a, **k: None)
float32 = np.float32
int32 = np.int32


# ---- running the reference's scripts unmodified --------------------------------------------------------------------
def install_as_tensorflow():
    """Register this module as `tensorflow` (sys.modules), so that a reference script's own first line
    `import tensorflow as tf` (examples/dbn_horses/train.py:1, gplvm_horses/train.py:1, gpdbn_horses/train.py:1)
    resolves to the shim without editing the script. Refuses to shadow a real TensorFlow."""
    import sys
    this = sys.modules[__name__]
    have = sys.modules.get('tensorflow')
    if have is not None and have is not this:
        raise RuntimeError('a real `tensorflow` module is already imported; not shadowing it')
    sys.modules['tensorflow'] = this
    return this


def _png_bytes(img):
    """Minimal 8-bit greyscale PNG encoder (zlib + CRC from the standard library)."""
    import struct
    import zlib
    a = np.asarray(img, dtype=np.float64)
    assert a.ndim == 2, 'greyscale images only'
    lo, hi = float(a.min()), float(a.max())
    a8 = np.zeros(a.shape, np.uint8) if hi <= lo else np.round((a - lo) * (255.0 / (hi - lo))).astype(np.uint8)
    raw = b''.join(b'\x00' + a8[r].tobytes() for r in range(a8.shape[0]))

    def chunk(tag, data):
        return struct.pack('>I', len(data)) + tag + data + struct.pack('>I', zlib.crc32(tag + data) & 0xffffffff)
    return (b'\x89PNG\r\n\x1a\n' + chunk(b'IHDR', struct.pack('>IIBBBBB', a8.shape[1], a8.shape[0], 8, 0, 0, 0, 0)) +
            chunk(b'IDAT', zlib.compress(raw, 6)) + chunk(b'IEND', b''))


def install_script_environment():
    """What `python -m deepbelief_b200.run <script>` sets up before executing a reference script verbatim: the
    `tensorflow` alias and `scipy.misc.imsave`, which the reference's test_generalisation.py scripts call
    (e.g. examples/gplvm_horses/test_generalisation.py:90) and which SciPy >= 1.2 no longer ships (min-max
    scaled 8-bit greyscale PNG, as the old function wrote)."""
    install_as_tensorflow()
    try:
        import warnings
        with warnings.catch_warnings():
            warnings.simplefilter('ignore')
            from scipy import misc
        if not hasattr(misc, 'imsave'):
            def imsave(name, arr, format=None):
                with open(name, 'wb') as f:
                    f.write(_png_bytes(arr))
            misc.imsave = imsave
    except ImportError:
        pass
