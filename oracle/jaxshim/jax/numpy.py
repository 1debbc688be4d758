"""jax.numpy subset on torch tensors."""
import math
import torch as _t

pi = math.pi


def _dt():
    import jax
    return jax.get_dtype()


def array(x, dtype=None):
    if isinstance(x, _t.Tensor):
        return x
    return _t.tensor(x, dtype=_dt())


asarray = array


def expand_dims(x, axis):
    return _t.unsqueeze(x, axis)


def concatenate(xs, axis=0):
    return _t.cat(list(xs), dim=axis)


def broadcast_to(x, shape):
    return _t.broadcast_to(x, tuple(shape))


def reshape(x, shape):
    return _t.reshape(x, tuple(shape))


def zeros_like(x):
    return _t.zeros_like(x)


def ones_like(x):
    return _t.ones_like(x)


def zeros(shape, dtype=None):
    return _t.zeros(shape, dtype=_dt())


def ones(shape, dtype=None):
    return _t.ones(shape, dtype=_dt())


def eye(n, m=None):
    return _t.eye(n, n if m is None else m, dtype=_dt())


def linspace(a, b, n):
    a = float(a)
    b = float(b)
    # computed in float64 then cast, like numpy/jnp linspace semantics
    return _t.linspace(a, b, int(n), dtype=_t.float64).to(_dt())


def _wrap(x):
    return x if isinstance(x, _t.Tensor) else _t.tensor(x, dtype=_dt())


def exp(x):
    return _t.exp(_wrap(x))


def log(x):
    return _t.log(_wrap(x))


def tanh(x):
    return _t.tanh(x)


def cos(x):
    return _t.cos(_wrap(x))


def sqrt(x):
    return _t.sqrt(_wrap(x))


def abs(x):  # noqa: A001
    return _t.abs(x)


def allclose(a, b, rtol=1e-5, atol=1e-8):
    return bool(_t.allclose(a, b, rtol=rtol, atol=atol))


ndarray = _t.Tensor
