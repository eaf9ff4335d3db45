"""Minimal TensorFlow *shim* -- TEST INFRASTRUCTURE ONLY (used by oracle/make_golden.py).

TensorFlow is not installable in this image.  This package implements, on PyTorch CPU
tensors, exactly the TF ops that the reference's hot-path source files call
(``lib/models/egt_layers.py`` and the ``mha_block`` / ``edge_update_*`` closures in
``lib/models/graph_xformer_model_base.py``), following their documented semantics, so
that the UNMODIFIED reference source can be imported from /root/reference and executed
to produce golden vectors.  It is not TensorFlow and makes no numerical claims beyond
"same op sequence, fp32/fp64 IEEE arithmetic".
"""
import torch as _t

float32 = _t.float32
float64 = _t.float64
int64 = _t.int64
int32 = _t.int32
bool = _t.bool  # noqa: A001

# the reference calls ``x.set_shape([...])`` on tensors (static-shape hints; no-ops here)
if not hasattr(_t.Tensor, 'set_shape'):
    _t.Tensor.set_shape = lambda self, shape: None


def _dt(dtype):
    return dtype


def reshape(x, shape):
    return x.reshape(tuple(int(s) for s in shape))


def shape(x):
    return tuple(x.shape)


def unstack(x, num=None, axis=0):
    if isinstance(x, tuple):
        return list(x)
    return list(x.unbind(dim=axis))


def einsum(eq, *ops):
    return _t.einsum(eq, *ops)


def clip_by_value(x, lo, hi):
    return _t.clamp(x, lo, hi)


def cast(x, dtype):
    return x.to(dtype)


def where(cond, a, b):
    a_ = a if isinstance(a, _t.Tensor) else _t.tensor(a, dtype=_t.get_default_dtype())
    b_ = b if isinstance(b, _t.Tensor) else _t.tensor(b, dtype=_t.get_default_dtype())
    return _t.where(cond, a_, b_)


def sigmoid(x):
    return _t.sigmoid(x)


def reduce_sum(x, axis=None, keepdims=False):
    if axis is None:
        return x.sum()
    return x.sum(dim=axis, keepdim=keepdims)


def pad(x, paddings, mode='CONSTANT', constant_values=0):
    assert mode == 'CONSTANT'
    flat = []
    for lo, hi in reversed(list(paddings)):
        flat += [int(lo), int(hi)]
    return _t.nn.functional.pad(x, flat, mode='constant', value=constant_values)


def tile(x, multiples):
    return x.repeat(*[int(m) for m in multiples])


def ones(shape, dtype=float32):
    return _t.ones(tuple(int(s) for s in shape), dtype=dtype)


def concat(values, axis):
    return _t.cat(list(values), dim=axis)


def expand_dims(x, axis):
    return x.unsqueeze(axis)


def split(x, sizes, axis=0):
    return list(_t.split(x, list(sizes), dim=axis))


class _Math:
    @staticmethod
    def log(x):
        return _t.log(x)

    @staticmethod
    def divide_no_nan(a, b):
        return _t.where(b == 0, _t.zeros_like(a), a / b)


math = _Math()


class _NN:
    @staticmethod
    def softmax(x, axis=-1):
        return _t.softmax(x, dim=axis)

    @staticmethod
    def dropout(x, rate):
        # tf.nn.dropout: keep where uniform >= rate, scale kept values by 1/(1-rate)
        noise = random.uniform(tuple(x.shape), dtype=x.dtype, _tag='dropout')
        return x * (noise >= rate).to(x.dtype) / (1.0 - rate)


nn = _NN()


class _Random:
    """Seeded uniform draws; every draw is recorded so it can be injected into the
    restatement (oracle/egt_oracle.py ``uniform_noise`` / ``dropout_noise``)."""

    def __init__(self):
        self.generator = _t.Generator().manual_seed(0)
        self.draws = []

    def seed(self, s):
        self.generator.manual_seed(int(s))
        self.draws = []

    def uniform(self, shape, minval=0., maxval=1., dtype=float32, _tag='uniform'):
        u = _t.rand(tuple(int(s) for s in shape), generator=self.generator, dtype=_t.float64)
        u = (u * (maxval - minval) + minval).to(dtype)
        self.draws.append((_tag, u))
        return u


random = _Random()


class VariableAggregation:
    ONLY_FIRST_REPLICA = 'only_first_replica'


from . import keras  # noqa: E402,F401
