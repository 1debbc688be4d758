"""flax-default parameter initialisation of DenseSAKELayer / DenseSAKEModel (sake/layers.py:54-105,
sake/models.py:23-54), torch-only: this module does NOT load the CUDA library, so the
benchmark's CPU reference arm can build the same seeded weights without touching the product code path
(bench.py loads it by file path)."""
import math

import torch


def exp_normal_smearing_init(num_rbf=50, cutoff_lower=0.0, cutoff_upper=5.0):
    # sake/utils.py:49-59 (PhysNet defaults); computed in fp64 then rounded to fp32 like jnp
    start = math.exp(-cutoff_upper + cutoff_lower)
    means = torch.linspace(start, 1.0, num_rbf, dtype=torch.float64).float()
    betas = torch.full((num_rbf,), (2.0 / num_rbf * (1.0 - start)) ** -2, dtype=torch.float64).float()
    return means, betas


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


def init_layer_params(gen, in_features, hidden_features, out_features, n_heads, update, has_v,
                      log_gamma=True, kernel_features=50, edge_features=0):
    """Parameter tree of one layer, in flax creation semantics (velocity_mlp only when it is
    actually called at init: layers.py:226-229)."""
    F, H, A, K = in_features, hidden_features, n_heads, kernel_features
    C = A * H
    E = edge_features          # width of `he` (sake/layers.py:201-202): h_cat_ht grows to 2F + E
    means, betas = exp_normal_smearing_init(K)
    p = {
        "edge_model": {
            "kernel": {"means": means, "betas": betas},
            "mlp_in": dense_init(gen, 2 * F + E, K),
            "mlp_out": {"layers_0": dense_init(gen, 2 * F + E + K + 1, H), "layers_2": dense_init(gen, H, H)},
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




def init_model_params(gen, in_features, hidden_features, out_features, depth, n_heads=4, update=True, has_v=False,
                      log_gamma=True):
    """Parameter tree of DenseSAKEModel (sake/models.py:24-54): embedding_in, d0..d{L-1}, embedding_out."""
    H = hidden_features
    upd = [update] * depth if isinstance(update, bool) else list(update)
    p = {"embedding_in": dense_init(gen, in_features, H),
         "embedding_out": {"layers_0": dense_init(gen, H, H), "layers_2": dense_init(gen, H, out_features)}}
    for i in range(depth):
        p["d%d" % i] = init_layer_params(gen, H, H, H, n_heads, upd[i], has_v, log_gamma=log_gamma)
        has_v = has_v or upd[i]
    return p
