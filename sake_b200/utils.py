"""Helpers with the reference's names (sake/utils.py).  ExpNormalSmearing itself runs inside the
fused edge kernel; `exp_normal_smearing_init` restates its initial parameters (utils.py:49-59)."""
import torch  # noqa: F401

from .init_params import exp_normal_smearing_init  # noqa: F401  (sake/utils.py:49-59)


def cosine_cutoff(x, lower=0.0, upper=5.0):
    """sake/utils.py:10-26 as written: the range masks computed there are discarded, so this is the whole function.
    Pass it (or functools.partial(cosine_cutoff, lower=..., upper=...)) as `cutoff=` of DenseSAKELayer / DenseSAKEModel:
    the layer evaluates it inside the attention kernels (SAKE_COSINE_CUTOFF), this torch version is for host-side use."""
    import math
    return 0.5 * (torch.cos(math.pi * (2 * (x - lower) / (upper - lower) + 1.0)) + 1.0)


def cutoff_params(cutoff):
    """None, or (lower, upper) when `cutoff` is cosine_cutoff / a functools.partial of it; anything else is not a
    closed form the kernels know and raises (the reference accepts an arbitrary callable, no script passes one)."""
    import functools
    import inspect
    if cutoff is None:
        return None
    fn, kw = cutoff, {}
    if isinstance(cutoff, functools.partial):
        fn, kw = cutoff.func, dict(cutoff.keywords or {})
        if cutoff.args:
            raise ValueError("cutoff partial must bind lower / upper by keyword")
    if getattr(fn, "__name__", "") != "cosine_cutoff":
        from ._lib import SakeError
        raise SakeError("cutoff must be cosine_cutoff or functools.partial(cosine_cutoff, lower=..., upper=...); "
                        "arbitrary callables cannot run inside the attention kernels")
    sig = inspect.signature(fn)
    lo = kw.get("lower", sig.parameters["lower"].default)
    hi = kw.get("upper", sig.parameters["upper"].default)
    return float(lo), float(hi)


def coloring(x, mean, std):
    # sake/utils.py:7-8
    return std * x + mean


def mae(x, y):
    # sake/utils.py:67-69
    return (x - y).abs().mean()
