"""Minimal Keras shim on torch CPU tensors -- TEST INFRASTRUCTURE ONLY (see ../__init__.py).

Eager: a layer call computes immediately.  Implements the Keras behaviours the reference's
block relies on: lazily built ``Dense`` (kernel ``[in,out]``, glorot-uniform, zero bias),
``LayerNormalization`` (axis=-1, epsilon=1e-3), ``Add``, ``Dropout``, ``LeakyReLU``,
``Activation``, ``Lambda``; implicit mask propagation through ``_keras_mask`` (a layer
whose ``call`` accepts ``mask=`` receives the input masks; ``compute_mask`` sets output
masks; layers with ``supports_masking`` pass the mask through)."""
import inspect as _inspect
import math as _math
import types as _types

import torch as _t

_STATE = dict(learning_phase=0, seed=1234, counter=0)


class _Backend:
    @staticmethod
    def learning_phase():
        return _STATE['learning_phase']

    @staticmethod
    def set_learning_phase(v):
        _STATE['learning_phase'] = int(v)


backend = _Backend()


def _mask_of(x):
    if isinstance(x, (list, tuple)):
        return [_mask_of(v) for v in x]
    return getattr(x, '_keras_mask', None)


def _set_mask(x, m):
    if isinstance(x, _t.Tensor) and m is not None:
        x._keras_mask = m


class Layer:
    def __init__(self, name=None, trainable=True, dtype=None, **kwargs):
        self.name = name
        self.supports_masking = False
        self.built = False
        self._weights = {}
        self.last_output = None

    # -- weights ---------------------------------------------------------------------
    def add_weight(self, name=None, shape=(), dtype=None, initializer=None, trainable=True, **kw):
        w = _t.zeros(tuple(shape), dtype=dtype if isinstance(dtype, _t.dtype) else _t.get_default_dtype())
        self._weights[name] = w
        return w

    def get_config(self):
        return dict(name=self.name)

    def build(self, input_shape):
        self.built = True

    def compute_mask(self, inputs, mask=None):
        if self.supports_masking:
            return mask
        return None

    def __call__(self, inputs, *args, **kwargs):
        if not self.built:
            shp = [tuple(v.shape) for v in inputs] if isinstance(inputs, (list, tuple)) else tuple(inputs.shape)
            self.build(shp)
            self.built = True
        mask = _mask_of(inputs)
        call = self.call
        params = _inspect.signature(call).parameters
        if 'mask' in params and 'mask' not in kwargs:
            kwargs['mask'] = mask
        out = call(inputs, *args, **kwargs)
        # output masks
        try:
            om = self.compute_mask(inputs, mask)
        except TypeError:
            om = None
        if isinstance(out, (list, tuple)):
            if isinstance(om, (list, tuple)):
                for o, m in zip(out, om):
                    _set_mask(o, m)
        else:
            if isinstance(om, (list, tuple)):
                om = om[0] if len(om) else None
            _set_mask(out, om)
        self.last_output = out
        return out

    def call(self, inputs, **kwargs):
        return inputs


def _next_generator():
    _STATE['counter'] += 1
    return _t.Generator().manual_seed(_STATE['seed'] * 1000 + _STATE['counter'])


_ACTS = dict(elu=_t.nn.functional.elu, relu=_t.relu, tanh=_t.tanh, sigmoid=_t.sigmoid,
             linear=lambda v: v)


def _activation(a):
    if a is None:
        return lambda v: v
    if callable(a):
        return a
    return _ACTS[a]


class Dense(Layer):
    def __init__(self, units, activation=None, use_bias=True, kernel_regularizer=None, **kw):
        super().__init__(**kw)
        self.units = int(units)
        self.activation = _activation(activation)
        self.use_bias = use_bias
        self.supports_masking = True

    def build(self, input_shape):
        fan_in = int(input_shape[-1])
        lim = _math.sqrt(6.0 / (fan_in + self.units))            # glorot_uniform
        g = _next_generator()
        dt = _t.get_default_dtype()
        self.kernel = ((_t.rand(fan_in, self.units, generator=g, dtype=_t.float64) * 2 - 1) * lim).to(dt)
        # Keras initialises biases to zero; use small non-zero values so bias bugs show up
        self.bias = ((_t.rand(self.units, generator=g, dtype=_t.float64) * 2 - 1) * 0.1).to(dt)
        self._weights = dict(kernel=self.kernel, bias=self.bias)

    def call(self, inputs):
        y = inputs @ self.kernel
        if self.use_bias:
            y = y + self.bias
        return self.activation(y)


class LayerNormalization(Layer):
    def __init__(self, axis=-1, epsilon=1e-3, center=True, scale=True, **kw):
        super().__init__(**kw)
        assert axis == -1
        self.epsilon = epsilon
        self.supports_masking = True

    def build(self, input_shape):
        c = int(input_shape[-1])
        g = _next_generator()
        dt = _t.get_default_dtype()
        self.gamma = (_t.rand(c, generator=g, dtype=_t.float64) + 0.5).to(dt)
        self.beta = ((_t.rand(c, generator=g, dtype=_t.float64) * 2 - 1) * 0.1).to(dt)
        self._weights = dict(gamma=self.gamma, beta=self.beta)

    def call(self, inputs):
        mean = inputs.mean(dim=-1, keepdim=True)
        var = ((inputs - mean) ** 2).mean(dim=-1, keepdim=True)
        return (inputs - mean) * _t.rsqrt(var + self.epsilon) * self.gamma + self.beta


class BatchNormalization(Layer):
    def __init__(self, **kw):
        raise NotImplementedError('batch normalisation is not on the shimmed path')


class Add(Layer):
    def __init__(self, **kw):
        super().__init__(**kw)
        self.supports_masking = True

    def compute_mask(self, inputs, mask=None):
        if isinstance(mask, (list, tuple)):
            for m in mask:
                if m is not None:
                    return m
            return None
        return mask

    def call(self, inputs):
        out = inputs[0]
        for v in inputs[1:]:
            out = out + v
        return out


class Dropout(Layer):
    def __init__(self, rate, **kw):
        super().__init__(**kw)
        self.rate = rate
        self.supports_masking = True

    def call(self, inputs, training=None):
        if training is None:
            training = backend.learning_phase()
        if training and self.rate > 0:
            import tensorflow as tf
            return tf.nn.dropout(inputs, self.rate)
        return inputs


class LeakyReLU(Layer):
    def __init__(self, alpha=0.3, **kw):
        super().__init__(**kw)
        self.alpha = alpha
        self.supports_masking = True

    def call(self, inputs):
        return _t.where(inputs >= 0, inputs, self.alpha * inputs)


class Activation(Layer):
    def __init__(self, activation, **kw):
        super().__init__(**kw)
        self.fn = _activation(activation)
        self.supports_masking = True

    def call(self, inputs):
        return self.fn(inputs)


class Lambda(Layer):
    def __init__(self, function, mask=None, **kw):
        super().__init__(**kw)
        self.function = function
        self._mask_fn = mask

    def compute_mask(self, inputs, mask=None):
        if self._mask_fn is not None:
            return self._mask_fn(inputs, mask)
        return None

    def call(self, inputs, mask=None):
        if 'mask' in _inspect.signature(self.function).parameters:
            return self.function(inputs, mask=mask)
        return self.function(inputs)


layers = _types.SimpleNamespace(Layer=Layer, Dense=Dense, LayerNormalization=LayerNormalization,
                                BatchNormalization=BatchNormalization, Add=Add, Dropout=Dropout,
                                LeakyReLU=LeakyReLU, Activation=Activation, Lambda=Lambda)


class _Callback:
    def __init__(self, *a, **k):
        pass


callbacks = _types.SimpleNamespace(Callback=_Callback)
models = _types.SimpleNamespace(Model=object)
regularizers = _types.SimpleNamespace(l2=lambda v: None)
initializers = _types.SimpleNamespace(Constant=lambda value=0: None)
losses = _types.SimpleNamespace()
metrics = _types.SimpleNamespace()
