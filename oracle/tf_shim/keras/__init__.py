"""Minimal Keras 3 API surface for executing the reference's python/model.py at inference (see ../README.md).

Layer semantics follow the Keras documentation:
  Conv2D(filters, k, padding="same", use_bias)  NHWC input, HWIO kernel, cross-correlation, zero "same" padding (odd k)
  Dense(units)                                  y = x @ kernel [in, units] + bias, on the last axis
  BatchNormalization(momentum, epsilon)         inference: gamma * (x - moving_mean) / sqrt(moving_variance + eps) + beta
  Reshape(target_shape) / Flatten               batch axis kept
Weights are created with seeded normal values at build time and overwritten by the golden generator."""
import inspect
import types

import numpy as np
import torch

_rng = np.random.default_rng(12345)


def _randn(*shape, scale=0.05):
    return torch.from_numpy(_rng.standard_normal(shape) * scale)


class Variable:
    """Just enough of keras.Variable: .numpy(), .assign(), and use as a tensor through .value."""

    def __init__(self, value, name=None):
        self.value = torch.as_tensor(value, dtype=torch.float64)
        self.name = name

    def numpy(self):
        return self.value.numpy()

    def assign(self, v):
        v = torch.as_tensor(np.asarray(v), dtype=torch.float64)
        assert tuple(v.shape) == tuple(self.value.shape), (self.name, v.shape, self.value.shape)
        self.value = v

    @property
    def shape(self):
        return tuple(self.value.shape)


# ---------------------------------------------------------------- activations / ops
def _mish(x):
    return x * torch.tanh(torch.nn.functional.softplus(x))


def _softmax(x, axis=-1):
    return torch.softmax(x, dim=axis)


def _linear(x):
    return x


def _relu(x):
    return torch.relu(x)


def _tanh(x):
    return torch.tanh(x)


def _sigmoid(x):
    return torch.sigmoid(x)


_ACTS = {"mish": _mish, "relu": _relu, "tanh": _tanh, "sigmoid": _sigmoid, "softmax": _softmax, "linear": _linear}
for _n, _f in _ACTS.items():
    _f.__name__ = _n
activations = types.SimpleNamespace(mish=_mish, relu=_relu, tanh=_tanh, sigmoid=_sigmoid, softmax=_softmax,
                                    linear=_linear,
                                    serialize=lambda f: getattr(f, "__name__", str(f)),
                                    deserialize=lambda n: _ACTS[n],
                                    get=lambda n: _ACTS[n] if isinstance(n, str) else n)


def _cast(x, dtype=None):
    x = torch.as_tensor(x)
    if "float" in str(dtype) and (not x.dtype.is_floating_point or "64" in str(dtype)):
        return x.to(torch.float64)
    return x


def _broadcast_to(x, shape):
    return torch.broadcast_to(x, tuple(int(s) for s in shape))


ops = types.SimpleNamespace(
    cast=_cast,
    transpose=lambda x, axes=None: x.permute(*axes),
    concatenate=lambda xs, axis=0: torch.cat(list(xs), dim=axis),
    zeros=lambda shape, dtype=None: torch.zeros(tuple(shape), dtype=torch.float64),
    shape=lambda x: tuple(x.shape),
    expand_dims=lambda x, axis: x.unsqueeze(axis),
    broadcast_to=_broadcast_to,
    squeeze=lambda x, axis=None: x.squeeze() if axis is None else x.squeeze(axis),
    softplus=torch.nn.functional.softplus,
    minimum=lambda a, b: torch.minimum(torch.as_tensor(a, dtype=torch.float64), torch.as_tensor(b, dtype=torch.float64)),
    mean=lambda x, axis=None, keepdims=False: x.mean() if axis is None else x.mean(dim=axis, keepdim=keepdims),
    max=lambda x, axis=None, keepdims=False: x.max() if axis is None else torch.amax(x, dim=axis, keepdim=keepdims),
    abs=torch.abs,
)


# ---------------------------------------------------------------- layers
class Layer:
    def __init__(self, name=None, **kwargs):
        self.name = name or type(self).__name__.lower()
        self.built = False
        self._weights = []

    def add_weight(self, name=None, shape=None, initializer=None, trainable=True, regularizer=None, dtype=None):
        v = Variable(_randn(*shape), name=name)
        self._weights.append(v)
        return v

    def build(self, input_shape):
        pass

    def __call__(self, *args, **kwargs):
        if not self.built:
            if args and isinstance(args[0], torch.Tensor):
                self.build(tuple(args[0].shape))
            self.built = True
        params = inspect.signature(self.call).parameters
        if "training" in kwargs and "training" not in params and not any(p.kind == p.VAR_KEYWORD for p in params.values()):
            kwargs.pop("training")
        return self.call(*args, **kwargs)

    def call(self, *args, **kwargs):
        raise NotImplementedError

    def get_config(self):
        return {"name": self.name}


class Model(Layer):
    pass


class Conv2D(Layer):
    def __init__(self, filters, kernel_size, activation=None, kernel_regularizer=None, padding="valid", use_bias=True,
                 kernel_initializer=None, name=None, **kw):
        super().__init__(name=name)
        assert padding == "same" and activation is None
        self.filters, self.kernel_size, self.use_bias = filters, int(kernel_size), use_bias
        self.kernel = self.bias = None

    def build(self, input_shape):
        k = self.kernel_size
        self.kernel = Variable(_randn(k, k, input_shape[-1], self.filters), name="kernel")  # HWIO
        if self.use_bias:
            self.bias = Variable(torch.zeros(self.filters), name="bias")

    def call(self, x, training=False):
        assert self.kernel_size % 2 == 1
        y = torch.nn.functional.conv2d(x.permute(0, 3, 1, 2), self.kernel.value.permute(3, 2, 0, 1), padding=self.kernel_size // 2)
        y = y.permute(0, 2, 3, 1)
        return y + self.bias.value if self.use_bias else y


class Dense(Layer):
    def __init__(self, units, kernel_initializer=None, kernel_regularizer=None, use_bias=True, activation=None, name=None, **kw):
        super().__init__(name=name)
        assert activation is None
        self.units, self.use_bias = units, use_bias
        self.kernel = self.bias = None

    def build(self, input_shape):
        self.kernel = Variable(_randn(input_shape[-1], self.units), name="kernel")
        if self.use_bias:
            self.bias = Variable(torch.zeros(self.units), name="bias")

    def call(self, x):
        y = x @ self.kernel.value
        return y + self.bias.value if self.use_bias else y


class BatchNormalization(Layer):
    def __init__(self, momentum=0.99, epsilon=1e-3, axis=-1, name=None, **kw):
        super().__init__(name=name)
        assert axis == -1
        self.momentum, self.epsilon = momentum, epsilon
        self.gamma = self.beta = self.moving_mean = self.moving_variance = None

    def build(self, input_shape):
        c = input_shape[-1]
        self.gamma = Variable(torch.ones(c), name="gamma")
        self.beta = Variable(torch.zeros(c), name="beta")
        self.moving_mean = Variable(torch.zeros(c), name="moving_mean")
        self.moving_variance = Variable(torch.ones(c), name="moving_variance")

    def call(self, x, training=False):
        assert not training
        return self.gamma.value * (x - self.moving_mean.value) / torch.sqrt(self.moving_variance.value + self.epsilon) + self.beta.value


class Rescaling(Layer):
    def __init__(self, scale=1.0, offset=0.0, name=None, **kw):
        super().__init__(name=name)
        self.scale, self.offset = scale, offset

    def call(self, x):
        return x * self.scale + self.offset


class Identity(Layer):
    def call(self, x):
        return x


class Reshape(Layer):
    def __init__(self, target_shape, name=None, **kw):
        super().__init__(name=name)
        self.target_shape = tuple(target_shape)

    def call(self, x):
        return x.reshape((x.shape[0],) + self.target_shape)


class Flatten(Layer):
    def call(self, x):
        return x.reshape(x.shape[0], -1)


class Activation(Layer):
    def __init__(self, activation, name=None, **kw):
        super().__init__(name=name)
        self.fn = activations.get(activation)

    def call(self, x):
        return self.fn(x)


def Input(*a, **k):
    raise NotImplementedError("functional API is not needed for inference through P3achyGoModel.call")


layers = types.SimpleNamespace(Layer=Layer, Conv2D=Conv2D, Dense=Dense, BatchNormalization=BatchNormalization, Rescaling=Rescaling,
                               Identity=Identity, Reshape=Reshape, Flatten=Flatten, Activation=Activation, Input=Input)


# ---------------------------------------------------------------- things that only have to exist
class _Stub:
    def __init__(self, *a, **k):
        self.args, self.kwargs = a, k

    def __call__(self, *a, **k):
        raise NotImplementedError("training-only object")


initializers = types.SimpleNamespace(VarianceScaling=_Stub, deserialize=lambda c: c)
regularizers = types.SimpleNamespace(L2=_Stub)
losses = types.SimpleNamespace(SparseCategoricalCrossentropy=_Stub, CategoricalCrossentropy=_Stub, MeanSquaredError=_Stub, Huber=_Stub)
metrics = types.SimpleNamespace(kl_divergence=None)
saving = types.SimpleNamespace(register_keras_serializable=lambda package=None, name=None: (lambda cls: cls))
