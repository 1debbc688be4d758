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
        from .utils import cutoff_params
        self._cutoff = cutoff_params(cutoff)     # None or (lower, upper) of cosine_cutoff (sake/utils.py:10-26)
        if activation is not None and getattr(activation, "__name__", "") not in ("silu", "swish"):
            raise ops._lib.SakeError("only the default silu activation is implemented")
        if out_features != hidden_features:
            raise ops._lib.SakeError("out_features must equal hidden_features (every reference call site does)")

    # -- flax-style API ---------------------------------------------------------------------
    def init(self, key, h, x, v=None, mask=None, he=None):
        if h.shape[-1] != self.out_features:
            raise ops._lib.SakeError("residual needs in_features == out_features (layers.py:150)")
        gen = _generator(key)
        p = init_layer_params(gen, h.shape[-1], self.hidden_features, self.out_features, self.n_heads,
                              self.update, v is not None,
                              log_gamma=self.use_semantic_attention and self.use_euclidean_attention,
                              edge_features=0 if he is None else he.shape[-1])
        return {"params": tree_to(p, h.device)}

    def apply(self, variables, *args, method=None, **kwargs):
        """flax's Module.apply: `apply(vars, h, x, v, mask, he)` runs the fused layer; `method=` (a bound method of
        this object or its name) runs one of the sub-methods below with the given positional / keyword arguments,
        as the reference's own tests do (sake/tests/test_mask.py:80,107,155,189)."""
        if method is None:
            return self(variables["params"], *args, **kwargs)
        name = method if isinstance(method, str) else method.__name__
        if name == "__call__":
            return self(variables["params"], *args, **kwargs)
        if name not in self.SUB_METHODS:
            raise ops._lib.SakeError(f"DenseSAKELayer has no sub-method `{name}` (the reference has none either: "
                                     "e.g. euclidean_attention only exists on SparseSAKELayer)")
        return getattr(self, name)(*args, params=variables["params"], **kwargs)

    # -- sub-methods of the reference module (sake/layers.py:108-186), for inspection and the reference's mask tests.
    # Plain torch on whatever device the tensors live on; the fused kernels compute the same quantities without
    # materialising them and never call these.
    SUB_METHODS = ("edge_model", "semantic_attention", "combined_attention", "spatial_attention", "aggregate",
                   "node_model", "velocity_model")

    @staticmethod
    def _dense(p, x):
        y = x @ p["kernel"]
        return y + p["bias"] if "bias" in p else y

    def edge_model(self, h_cat_ht, x_norm, *, params):
        # ContinuousFilterConvolutionWithConcatenation.__call__, sake/layers.py:28-40; RBF: sake/utils.py:61-65
        p = params["edge_model"]
        u = self._dense(p["mlp_in"], h_cat_ht)
        rbf = torch.exp(-p["kernel"]["betas"] * (torch.exp(-x_norm) - p["kernel"]["means"]) ** 2)
        z = torch.cat([h_cat_ht, rbf * u, x_norm], dim=-1)
        z = torch.nn.functional.silu(self._dense(p["mlp_out"]["layers_0"], z))
        return self._dense(p["mlp_out"]["layers_2"], z)

    def semantic_attention(self, h_e_mtx, mask=None, *, params):
        # sake/layers.py:153-168
        att = torch.nn.functional.celu(self._dense(params["semantic_attention_mlp"]["layers_0"], h_e_mtx), alpha=2.0)
        n = att.shape[-2]
        att = att - 1e5 * torch.eye(n, n, dtype=att.dtype, device=att.device).unsqueeze(-1)
        if mask is not None:
            att = att - 1e5 * (1 - mask.unsqueeze(-1))
        return torch.softmax(att, dim=-2)

    def combined_attention(self, x_minus_xt_norm, h_e_mtx, mask=None, *, params):
        # sake/layers.py:170-182, with the guarded normalisation of the kernels (a fully masked row gives 0, not 0/0)
        sem = self.semantic_attention(h_e_mtx, mask=mask, params=params)
        if self._cutoff is not None:
            from .utils import cosine_cutoff
            euc = cosine_cutoff(x_minus_xt_norm, *self._cutoff)
        else:
            euc = 1.0
        comb = euc * sem
        if mask is not None:
            comb = comb * mask.unsqueeze(-1)
        den = comb.sum(dim=-2, keepdim=True)
        comb = comb / torch.where(den > 0, den, torch.ones_like(den))
        return euc, sem, comb

    def spatial_attention(self, h_e_att, x_minus_xt, x_minus_xt_norm, mask=None, *, params):
        # sake/layers.py:108-133
        coef = torch.tanh(h_e_att @ params["x_mixing"]["layers_0"]["kernel"])
        direction = x_minus_xt / (x_minus_xt_norm + 1e-5)
        comb = direction.unsqueeze(-2) * coef.unsqueeze(-1)
        if mask is not None:
            m = mask.unsqueeze(-1).unsqueeze(-1)
            comb = comb * m
            csum = comb.sum(dim=-3) / (m.sum(dim=-3) + 1e-8)
        else:
            csum = comb.mean(dim=-3)
        nrm = (csum ** 2).sum(-1)
        p = params["post_norm_mlp"]
        hc = torch.nn.functional.silu(self._dense(p["layers_0"], nrm))
        hc = torch.nn.functional.silu(self._dense(p["layers_2"], hc))
        return hc, comb

    def aggregate(self, h_e_mtx, mask=None, *, params=None):
        # sake/layers.py:135-140
        if mask is not None:
            h_e_mtx = h_e_mtx * mask.unsqueeze(-1)
        return h_e_mtx.sum(dim=-2)

    def node_model(self, h, h_e, h_combinations, *, params):
        # sake/layers.py:142-151
        p = params["node_mlp"]
        out = torch.cat([h, h_e, h_combinations], dim=-1)
        out = torch.nn.functional.silu(self._dense(p["layers_0"], out))
        out = torch.nn.functional.silu(self._dense(p["layers_2"], out))
        return h + out

    def velocity_model(self, v, h, *, params):
        # sake/layers.py:69-76,184-186
        p = params["velocity_mlp"]
        g = torch.nn.functional.silu(self._dense(p["layers_0"], h))
        return 2.0 * torch.sigmoid(g @ p["layers_2"]["kernel"]) * v

    def __call__(self, params, h, x, v=None, mask=None, he=None):
        flat = flatten_tree(params)
        D = x.shape[-1]
        pair_u = pair_p = None
        if he is not None:
            # Edge features (sake/layers.py:201-202: h_cat_ht = concat(h_cat_ht, he)) only meet rows [2F, 2F+E) of
            # mlp_in and mlp_out[0]: two small GEMMs here give the per-pair terms the kernels add (SakePairTerms,
            # include/sake_b200.h), the kernels get the two kernels without those rows, and autograd carries the
            # returned cotangents back to `he` and to the two row blocks.
            F2, E = 2 * h.shape[-1], he.shape[-1]
            w_in, w_1 = flat["edge_model/mlp_in/kernel"], flat["edge_model/mlp_out/layers_0/kernel"]
            if w_in.shape[0] != F2 + E:
                raise ops._lib.SakeError(f"mlp_in kernel has {w_in.shape[0]} rows, expected 2F + E = {F2 + E}")
            K = w_in.shape[1]
            pair_u = torch.nn.functional.pad(he.float() @ w_in[F2:F2 + E], (0, (-K) % 4))
            pair_p = he.float() @ w_1[F2:F2 + E]
            flat = dict(flat)
            flat["edge_model/mlp_in/kernel"] = w_in[:F2]
            flat["edge_model/mlp_out/layers_0/kernel"] = torch.cat([w_1[:F2], w_1[F2 + E:]], dim=0)
        ho, xo, vo = ops.sake_layer(flat, h, _pad3(x), _pad3(v), mask, n_heads=self.n_heads,
                                    update=self.update, use_spatial_attention=self.use_spatial_attention,
                                    engine=self.engine, cutoff=self._cutoff, pair_u=pair_u, pair_p=pair_p)
        if D != 3:
            xo = xo[..., :D]
            vo = None if vo is None else vo[..., :D]
        return ho, xo, vo
