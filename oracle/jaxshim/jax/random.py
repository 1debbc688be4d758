"""jax.random subset: keys are torch Generators (values differ from threefry; only used for init)."""
import torch as _t


def PRNGKey(seed):  # noqa: N802
    g = _t.Generator()
    g.manual_seed(int(seed))
    return g


def _dt():
    import jax
    return jax.get_dtype()


def normal(key, shape, dtype=None):
    return _t.randn(tuple(shape), generator=key, dtype=_t.float64).to(_dt())


def uniform(key, shape, dtype=None, minval=0.0, maxval=1.0):
    return (_t.rand(tuple(shape), generator=key, dtype=_t.float64) * (maxval - minval) + minval).to(_dt())
