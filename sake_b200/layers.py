"""DenseSAKELayer with the reference's flax-linen surface (sake/layers.py:42-52,107-235):
same constructor fields / defaults / positional order, `.init(key, h, x, v, mask)` returning
`{'params': tree}` with flax's parameter names, `.apply(variables, h, x, ...)`.
Tensors are torch CUDA tensors; the arithmetic is the CUDA library (ops.sake_layer)."""
import math

import torch

from . import ops
from .utils import exp_normal_smearing_init


def _generator(key):
    if isinstance(key, torch.Generator):
        return key
    g = torch.Generator()
    g.manual_seed(int(key))
    return g


def lecun_normal(gen, shape):
    """flax default kernel init: truncated normal (+-2 sigma), std = sqrt(1/fan_in)/0.8796."""
    w = torch.empty(tuple(shape), dtype=torch.float64)
    torch.nn.init.trunc_normal_(w, mean=0.0, std=1.0, a=-2.0, b=2.0, generator=gen)
    return (w * (math.sqrt(1.0 / shape[0]) / 0.87962566103423978)).float()


def dense_init(gen, fan_in, fan_out, use_bias=True):
    p = {"kernel": lecun_normal(gen, (fan_in, fan_out))}
    if use_bias:
        p["bias"] = torch.zeros(fan_out)
    return p


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


def init_layer_params(gen, in_features, hidden_features, out_features, n_heads, update, has_v,
                      log_gamma=True, kernel_features=50):
    """Parameter tree of one layer, in flax creation semantics (velocity_mlp only when it is
    actually called at init: layers.py:226-229)."""
    F, H, A, K = in_features, hidden_features, n_heads, kernel_features
    C = A * H
    means, betas = exp_normal_smearing_init(K)
    p = {
        "edge_model": {
            "kernel": {"means": means, "betas": betas},
            "mlp_in": dense_init(gen, 2 * F, K),
            "mlp_out": {"layers_0": dense_init(gen, 2 * F + K + 1, H), "layers_2": dense_init(gen, H, H)},
        },
    }
    if log_gamma:
        p["log_gamma"] = -torch.log(torch.linspace(1.0, 5.0, A))
    p["semantic_attention_mlp"] = {"layers_0": dense_init(gen, H, A)}
    p["x_mixing"] = {"layers_0": dense_init(gen, C, C, use_bias=False)}
    p["post_norm_mlp"] = {"layers_0": dense_init(gen, C, H), "layers_2": dense_init(gen, H, H)}
    p["node_mlp"] = {"layers_0": dense_init(gen, F + C + H, H), "layers_2": dense_init(gen, H, out_features)}
    if update:
        p["v_mixing"] = dense_init(gen, C, 1, use_bias=False)
        if has_v:
            p["velocity_mlp"] = {"layers_0": dense_init(gen, out_features, H),
                                 "layers_2": dense_init(gen, H, 1, use_bias=False)}
    return p


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
