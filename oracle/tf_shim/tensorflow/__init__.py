"""Minimal TensorFlow API surface for executing the reference's python/model.py at inference (see ../README.md).
Tensors are torch.float64 CPU tensors."""
import types

import torch

Tensor = torch.Tensor
float16, float32, float64, int32, int64, bool = "float16", "float32", "float64", "int32", "int64", "bool"
bfloat16 = "bfloat16"


def _is_float(dtype):
    return dtype is None or "float" in str(dtype)


def cast(x, dtype=None):
    """Casts never LOSE precision here (the golden vectors are float64): a float32 target keeps the tensor as it is, a
    float64 target (e.g. `v_pooled.dtype`) widens float32 constants."""
    x = torch.as_tensor(x)
    if _is_float(dtype):
        if not x.dtype.is_floating_point or "64" in str(dtype):
            return x.to(torch.float64)
        return x
    if "int" in str(dtype):
        return x.to(torch.int64)
    return x


def constant(v, dtype=None):
    return cast(torch.as_tensor(v), dtype)


def range(start, limit=None, delta=1, dtype=None):  # noqa: A001  (tf.range)
    """A float32 range stays float32, so that constants derived from it (the score-bin vector, python/model.py:1225-1228)
    carry TensorFlow's float32 rounding; they are widened when they meet float64 activations."""
    if limit is None:
        start, limit = 0, start
    t = torch.arange(start, limit, delta)
    if dtype is not None and "float32" in str(dtype):
        return t.to(torch.float32)
    return cast(t, dtype)


def reduce_mean(x, axis=None, keepdims=False):
    return x.mean() if axis is None else x.mean(dim=axis, keepdim=keepdims)


def reduce_sum(x, axis=None, keepdims=False):
    return x.sum() if axis is None else x.sum(dim=axis, keepdim=keepdims)


def square(x):
    return x * x


def sqrt(x):
    return torch.sqrt(x)


def stop_gradient(x):
    return x


def clip_by_value(x, lo, hi):
    return torch.clamp(x, lo, hi)


def squeeze(x, axis=None):
    return x.squeeze() if axis is None else x.squeeze(axis)


def zeros_like(x):
    return torch.zeros_like(x)


def maximum(a, b):
    return torch.maximum(torch.as_tensor(a, dtype=torch.float64), torch.as_tensor(b, dtype=torch.float64))


def pow(x, y):  # noqa: A001
    return torch.pow(x, y)


def _conv2d(x, w, strides=(1, 1, 1, 1), padding="SAME"):
    assert padding == "SAME" and tuple(strides) == (1, 1, 1, 1)
    k = w.shape[0]
    y = torch.nn.functional.conv2d(x.permute(0, 3, 1, 2), w.permute(3, 2, 0, 1), padding=k // 2)
    return y.permute(0, 2, 3, 1)


nn = types.SimpleNamespace(sigmoid=torch.sigmoid, conv2d=_conv2d)
math = types.SimpleNamespace(reduce_mean=reduce_mean, reduce_sum=reduce_sum, square=square,
                             cumsum=lambda x, axis=0: torch.cumsum(x, dim=axis))


class _Device:
    def __init__(self, *_):
        pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        return False


device = _Device
function = lambda f=None, **kw: (f if f is not None else (lambda g: g))  # noqa: E731
