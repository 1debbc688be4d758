"""Torch-backed shim of the jax API surface used by the SAKE reference (test infrastructure)."""
import torch as _torch
from . import numpy, nn, random  # noqa: F401
from . import experimental  # noqa: F401

_DTYPE = [_torch.float64]


def set_dtype(dt):
    _DTYPE[0] = dt


def get_dtype():
    return _DTYPE[0]


def jit(f=None, **kw):
    if f is None:
        return lambda g: g
    return f


def grad(f, argnums=0):
    def g(*args):
        args = list(args)
        a = args[argnums].detach().clone().requires_grad_(True)
        args[argnums] = a
        out = f(*args)
        (gr,) = _torch.autograd.grad(out, a)
        return gr
    return g


def vmap(f, *a, **k):
    raise NotImplementedError("jaxshim: vmap not provided")


class lax:  # noqa: N801
    @staticmethod
    def stop_gradient(x):
        return x.detach()
