"""jax.nn subset on torch tensors."""
import torch as _t


def silu(x):
    return x * _t.sigmoid(x)


swish = silu


def sigmoid(x):
    return _t.sigmoid(x)


def relu(x):
    return _t.relu(x)


def softmax(x, axis=-1):
    # jax.nn.softmax: exp(x - stop_gradient(max)) / sum
    return _t.softmax(x, dim=axis)


def celu(x, alpha=1.0):
    # jax.nn.celu: max(x,0) + alpha*expm1(min(x,0)/alpha)
    return _t.clamp(x, min=0) + alpha * _t.expm1(_t.clamp(x, max=0) / alpha)


def tanh(x):
    return _t.tanh(x)
