"""DenseSAKELayer with the reference's flax-linen surface (sake/layers.py:42-52,107-235):
same constructor fields / defaults / positional order, `.init(key, h, x, v, mask)` returning
`{'params': tree}` with flax's parameter names, `.apply(variables, h, x, ...)`.
Tensors are torch CUDA tensors; the arithmetic is the CUDA library (ops.sake_layer)."""
import math

import torch

from . import ops
from .init_params import (_generator, dense_init, init_layer_params, lecun_normal,  # noqa: F401
                          init_model_params)


def flatten_tree(t, prefix=""):
    out = {}
    for k, v in t.items():
        if isinstance(v, dict):
            out.update(flatten_tree(v, prefix + k + "/"))
        else:
            out[prefix + k] = v
    return out


def unflatten_tree(flat):
    t = {}
    for k, v in flat.items():
        parts = k.split("/")
        d = t
        for q in parts[:-1]:
            d = d.setdefault(q, {})
        d[parts[-1]] = v
    return t


def tree_to(t, device):
    return {k: (tree_to(v, device) if isinstance(v, dict) else v.to(device)) for k, v in t.items()}


def _pad3(t):
    """2-D systems (scripts/dw4/run.py:17-19): zero-pad coordinates to 3 columns (exact)."""
    if t is None or t.shape[-1] == 3:
        return t
    if t.shape[-1] > 3:
        raise ops._lib.SakeError("coordinates must have <= 3 columns")
    return torch.cat([t, t.new_zeros(*t.shape[:-1], 3 - t.shape[-1])], dim=-1)


class DenseSAKELayer:
    """Drop-in for sake.layers.DenseSAKELayer (sake/layers.py:42-52,107-235)."""

    def __init__(self, out_features, hidden_features, activation=None, n_heads=4, update=True,
                 use_semantic_attention=True, use_euclidean_attention=True, use_spatial_attention=True,
                 cutoff=None, engine="auto"):
        self.out_features = out_features
        self.hidden_features = hidden_features
        self.activation = activation          # only silu (the default) is implemented in the kernels
        self.n_heads = n_heads
        self.update = update
        self.use_semantic_attention = use_semantic_attention
        self.use_euclidean_attention = use_euclidean_attention
        self.use_spatial_attention = use_spatial_attention
        self.cutoff = cutoff
        self.engine = engine
        if cutoff is not None:
            raise ops._lib.SakeError("cutoff is not supported by the CUDA layer (no script passes it)")
        if activation is not None and getattr(activation, "__name__", "") not in ("silu", "swish"):
            raise ops._lib.SakeError("only the default silu activation is implemented")
        if out_features != hidden_features:
            raise ops._lib.SakeError("out_features must equal hidden_features (every reference call site does)")

    # -- flax-style API ---------------------------------------------------------------------
    def init(self, key, h, x, v=None, mask=None, he=None):
        if he is not None:
            raise ops._lib.SakeError("edge features `he` are not supported")
        if h.shape[-1] != self.out_features:
            raise ops._lib.SakeError("residual needs in_features == out_features (layers.py:150)")
        gen = _generator(key)
        p = init_layer_params(gen, h.shape[-1], self.hidden_features, self.out_features, self.n_heads,
                              self.update, v is not None,
                              log_gamma=self.use_semantic_attention and self.use_euclidean_attention)
        return {"params": tree_to(p, h.device)}

    def apply(self, variables, h, x, v=None, mask=None, he=None, method=None):
        if method is not None:
            raise ops._lib.SakeError("sub-method application is not exposed by the fused layer")
        return self(variables["params"], h, x, v, mask, he)

    def __call__(self, params, h, x, v=None, mask=None, he=None):
        if he is not None:
            raise ops._lib.SakeError("edge features `he` are not supported")
        flat = flatten_tree(params)
        D = x.shape[-1]
        ho, xo, vo = ops.sake_layer(flat, h, _pad3(x), _pad3(v), mask, n_heads=self.n_heads,
                                    update=self.update, use_spatial_attention=self.use_spatial_attention,
                                    engine=self.engine)
        if D != 3:
            xo = xo[..., :D]
            vo = None if vo is None else vo[..., :D]
        return ho, xo, vo
