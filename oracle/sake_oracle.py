"""CPU oracle for the SAKE dense hot path — TEST INFRASTRUCTURE ONLY.

A torch (CPU, fp64 or fp32) restatement of the reference algorithm, written as the reference
writes it (naive form: every O(N^2) tensor is materialised).  Only `tests/`,
`__graft_entry__.smoke()` and `bench.py`'s cpu_baseline / `--impl reference` leg may import
this file; the product package `sake_b200/` never does.

Reference files restated (paths relative to /root/reference):
  sake/functional.py:7-44   get_x_minus_xt, get_x_minus_xt_norm, get_h_cat_ht
  sake/utils.py:28-65       ExpNormalSmearing
  sake/layers.py:9-40       double_sigmoid, ContinuousFilterConvolutionWithConcatenation
  sake/layers.py:54-105     SAKELayer.setup (parameter declarations)
  sake/layers.py:107-235    DenseSAKELayer
  sake/models.py:11-61      DenseSAKEModel
  sake/flows.py:12-27,97-188  CenteredGaussian, AugmentedFlowLayer/Model
  scripts/md17/run.py:46-58 energy / force closures

Pinning: the reference publishes no golden vectors (its tests are shape/invariance properties).
This oracle is pinned against the *unmodified reference source* executed under
`oracle/jaxshim` (a torch-backed stand-in for the jax/flax calls; JAX itself is not installable
here) — see `oracle/gen_golden.py`, `tests/golden/*.npz`, `tests/test_oracle_golden.py`.

Third-party semantics restated (jax / flax are un-vendored, unpinned deps of the reference:
requirements.txt:1-3): Dense y = x @ kernel + bias; silu = x*sigmoid(x);
celu(x,a) = max(x,0) + a*expm1(min(x,0)/a); softmax with max subtraction; relu'(0) = 0.

Mask semantics: `guarded=False` follows the reference exactly (fully masked rows give 0/0 = NaN,
sake/layers.py:178-180).  `guarded=True` (what the CUDA path implements) defines att = 0 for a
row whose normaliser is 0; everything else is identical.
"""
import math
import numpy as np
import torch

EPSILON = 1e-5  # sake/functional.py:4
INF = 1e5       # sake/functional.py:5


# ----------------------------------------------------------------------------------------------
# functional.py
# ----------------------------------------------------------------------------------------------
def get_x_minus_xt(x):
    # sake/functional.py:7-8 : r[i,j] = x[j] - x[i]
    return x.unsqueeze(-3) - x.unsqueeze(-2)


def get_x_minus_xt_norm(x_minus_xt, epsilon=EPSILON):
    # sake/functional.py:10-19
    return (torch.relu((x_minus_xt ** 2).sum(dim=-1, keepdim=True)) + epsilon) ** 0.5


def get_h_cat_ht(h):
    # sake/functional.py:33-44 : hc[i,j] = [h[j] | h[i]]
    n = h.shape[-2]
    shape = (*h.shape[:-2], n, n, h.shape[-1])
    return torch.cat([h.unsqueeze(-3).expand(shape), h.unsqueeze(-2).expand(shape)], dim=-1)


# ----------------------------------------------------------------------------------------------
# activations
# ----------------------------------------------------------------------------------------------
def silu(x):
    return x * torch.sigmoid(x)


def celu(x, alpha):
    return torch.clamp(x, min=0) + alpha * torch.expm1(torch.clamp(x, max=0) / alpha)


def dense(p, x):
    y = x @ p["kernel"]
    if "bias" in p:
        y = y + p["bias"]
    return y


# ----------------------------------------------------------------------------------------------
# utils.py / layers.py
# ----------------------------------------------------------------------------------------------
def exp_normal_smearing(p, dist, cutoff_lower=0.0, cutoff_upper=5.0):
    # sake/utils.py:34,61-65
    alpha = 5.0 / (cutoff_upper - cutoff_lower)
    return torch.exp(-p["betas"] * (torch.exp(alpha * (-dist + cutoff_lower)) - p["means"]) ** 2)


def edge_model(p, h_cat_ht, x_norm):
    # sake/layers.py:28-40
    h0 = h_cat_ht
    h = dense(p["mlp_in"], h_cat_ht)
    _x = exp_normal_smearing(p["kernel"], x_norm) * h
    z = torch.cat([h0, _x, x_norm], dim=-1)
    z = dense(p["mlp_out"]["layers_0"], z)
    z = silu(z)
    return dense(p["mlp_out"]["layers_2"], z)


def semantic_attention(p, h_e_mtx, mask=None):
    # sake/layers.py:153-168
    att = celu(dense(p["semantic_attention_mlp"]["layers_0"], h_e_mtx), 2.0)
    n = att.shape[-2]
    att = att - INF * torch.eye(n, n, dtype=att.dtype).unsqueeze(-1)
    if mask is not None:
        att = att - INF * (1 - mask.unsqueeze(-1))
    return torch.softmax(att, dim=-2)


def cosine_cutoff(x, lower=0.0, upper=5.0):
    # sake/utils.py:10-26 (the range masks computed at :24-25 are discarded there: the cosine is the whole function)
    return 0.5 * (torch.cos(math.pi * (2 * (x - lower) / (upper - lower) + 1.0)) + 1.0)


def combined_attention(p, x_norm, h_e_mtx, mask=None, guarded=True, cutoff=None):
    # sake/layers.py:170-182; cutoff: None (euclidean_attention = 1.0) or a callable of the pair distances
    sem = semantic_attention(p, h_e_mtx, mask=mask)
    comb = (1.0 if cutoff is None else cutoff(x_norm)) * sem
    if mask is not None:
        comb = comb * mask.unsqueeze(-1)
    den = comb.sum(dim=-2, keepdim=True)
    if guarded:
        den = torch.where(den > 0, den, torch.ones_like(den))
    return comb / den


def spatial_attention(p, h_e_att, x_minus_xt, x_norm, mask=None):
    # sake/layers.py:108-133
    coefficients = torch.tanh(h_e_att @ p["x_mixing"]["layers_0"]["kernel"])
    direction = x_minus_xt / (x_norm + 1e-5)
    combinations = direction.unsqueeze(-2) * coefficients.unsqueeze(-1)   # [..., i, j, C, D]
    if mask is not None:
        _mask = mask.unsqueeze(-1).unsqueeze(-1)
        combinations = combinations * _mask
        combinations_sum = combinations.sum(dim=-3) / (_mask.sum(dim=-3) + 1e-8)
    else:
        combinations_sum = combinations.mean(dim=-3)
    combinations_norm = (combinations_sum ** 2).sum(-1)
    hp = silu(dense(p["post_norm_mlp"]["layers_0"], combinations_norm))
    hp = silu(dense(p["post_norm_mlp"]["layers_2"], hp))
    return hp, combinations


def layer_forward(p, h, x, v=None, mask=None, *, update=True, use_spatial_attention=True,
                  guarded=True, cutoff=None, he=None):
    """DenseSAKELayer.__call__  (sake/layers.py:188-235)."""
    x_minus_xt = get_x_minus_xt(x)
    x_norm = get_x_minus_xt_norm(x_minus_xt)
    h_cat_ht = get_h_cat_ht(h)
    if he is not None:                                   # sake/layers.py:201-202
        h_cat_ht = torch.cat([h_cat_ht, he], dim=-1)
    h_e_mtx = edge_model(p["edge_model"], h_cat_ht, x_norm)
    att = combined_attention(p, x_norm, h_e_mtx, mask=mask, guarded=guarded, cutoff=cutoff)
    h_e_att = h_e_mtx.unsqueeze(-1) * att.unsqueeze(-2)                    # [..., i, j, H, A]
    h_e_att = h_e_att.reshape(*h_e_att.shape[:-2], -1)                    # c = f*A + a
    h_comb, delta_v = spatial_attention(p, h_e_att, x_minus_xt, x_norm, mask=mask)
    if not use_spatial_attention:
        h_comb = torch.zeros_like(h_comb)
        delta_v = torch.zeros_like(delta_v)
    # aggregate, sake/layers.py:135-140
    hm = h_e_att if mask is None else h_e_att * mask.unsqueeze(-1)
    h_e = hm.sum(dim=-2)
    # node_model, sake/layers.py:142-151
    out = torch.cat([h, h_e, h_comb], dim=-1)
    out = silu(dense(p["node_mlp"]["layers_0"], out))
    out = silu(dense(p["node_mlp"]["layers_2"], out))
    h = h + out
    if update:
        # sake/layers.py:218-232
        dv = (delta_v.transpose(-1, -2) @ p["v_mixing"]["kernel"]).transpose(-1, -2)  # [...,i,j,1,D]
        if mask is not None:
            dv = dv.sum(dim=(-2, -3)) / (mask.sum(-1, keepdim=True) + 1e-10)
        else:
            dv = dv.mean(dim=(-2, -3))
        if v is not None:
            g = silu(dense(p["velocity_mlp"]["layers_0"], h))
            g = 2.0 * torch.sigmoid(g @ p["velocity_mlp"]["layers_2"]["kernel"])
            v = g * v
        else:
            v = torch.zeros_like(x)
        v = dv + v
        x = x + v
    return h, x, v


def model_forward(params, h, x, v=None, mask=None, *, update=True, use_spatial_attention=True,
                  guarded=True, cutoff=None, he=None):
    """DenseSAKEModel.__call__ (sake/models.py:56-61).  depth = number of 'd<k>' entries."""
    depth = sum(1 for k in params if k.startswith("d") and k[1:].isdigit())
    upd = [update] * depth if isinstance(update, bool) else list(update)
    h = dense(params["embedding_in"], h)
    for k in range(depth):
        h, x, v = layer_forward(params["d%d" % k], h, x, v, mask, update=upd[k],
                                use_spatial_attention=use_spatial_attention, guarded=guarded, cutoff=cutoff, he=he)
    h = silu(dense(params["embedding_out"]["layers_0"], h))
    h = dense(params["embedding_out"]["layers_2"], h)
    return h, x, v


def energy(params, h, x, mask=None, atom_mask=None, **kw):
    """E[b] = sum_i out[b,i,:]  (scripts/md17/run.py:46-52; masked sum as scripts/qm9/run.py:58-60)."""
    y, _, _ = model_forward(params, h, x, mask=mask, **kw)
    if atom_mask is not None:
        y = y * atom_mask.unsqueeze(-1)
    return y.sum(dim=(-1, -2))


def energy_and_forces(params, h, x, mask=None, atom_mask=None, **kw):
    """F = -d(sum E)/dx (scripts/md17/run.py:54-58)."""
    x = x.detach().clone().requires_grad_(True)
    e = energy(params, h, x, mask=mask, atom_mask=atom_mask, **kw)
    (g,) = torch.autograd.grad(e.sum(), x)
    return e.detach(), -g


# ----------------------------------------------------------------------------------------------
# flows.py (config 5)
# ----------------------------------------------------------------------------------------------
def centered_gaussian_log_prob(value):
    # sake/flows.py:13-21
    n, d = value.shape[-2], value.shape[-1]
    r2 = (value ** 2).reshape(*value.shape[:-2], -1).sum(-1)
    return -0.5 * r2 - 0.5 * (n - 1) * d * math.log(2 * math.pi)


def flow_mp(p, h, x):
    # sake/flows.py:118-129
    x0 = x
    h = torch.cat([h, (x ** 2).sum(-1, keepdim=True)], dim=-1)
    h = torch.cat([h, torch.zeros_like(h[..., -1:, :])], dim=-2)
    x = torch.cat([x, torch.zeros_like(x[..., -1:, :])], dim=-2)
    h, x, _ = model_forward(p["sake_model"], h, x)
    x = x[..., :-1, :]
    h = h[..., :-1, :]
    translation = x - x0
    translation = translation - translation.mean(dim=-2, keepdim=True)
    s = silu(dense(p["scale_mlp"]["layers_0"], h))
    s = torch.tanh(s @ p["scale_mlp"]["layers_2"]["kernel"])
    scale = s.mean(dim=-2, keepdim=True)
    return scale, translation


def flow_layer_forward(p, h, x, v):
    # sake/flows.py:131-135
    scale, translation = flow_mp(p, h, x)
    v = torch.exp(scale) * v + translation
    log_det = scale.sum((-1, -2)) * v.shape[-1] * v.shape[-2]
    return x, v, log_det


def flow_layer_backward(p, h, x, v):
    # sake/flows.py:137-142
    scale, translation = flow_mp(p, h, x)
    v = torch.exp(-scale) * (v - translation)
    log_det = scale.sum((-1, -2)) * v.shape[-1] * v.shape[-2]
    return x, v, log_det


def flow_forward(params, h, x, v):
    # sake/flows.py:168-176
    depth = sum(1 for k in params if k.startswith("xv_"))
    s = 0.0
    for k in reversed(range(depth)):
        x, v, ld = flow_layer_forward(params["xv_%d" % k], h, x, v)
        s = s + ld
        v, x, ld = flow_layer_forward(params["vx_%d" % k], h, v, x)
        s = s + ld
    return x, v, s


def flow_backward(params, h, x, v):
    # sake/flows.py:178-186
    depth = sum(1 for k in params if k.startswith("xv_"))
    s = 0.0
    for k in range(depth):
        v, x, ld = flow_layer_backward(params["vx_%d" % k], h, v, x)
        s = s + ld
        x, v, ld = flow_layer_backward(params["xv_%d" % k], h, x, v)
        s = s + ld
    return x, v, s


# ----------------------------------------------------------------------------------------------
# parameter trees (flax naming)
# ----------------------------------------------------------------------------------------------
def tree_map(f, t):
    return {k: (tree_map(f, v) if isinstance(v, dict) else f(v)) for k, v in t.items()}


def tree_flatten(t, prefix=""):
    out = {}
    for k, v in t.items():
        if isinstance(v, dict):
            out.update(tree_flatten(v, prefix + k + "/"))
        else:
            out[prefix + k] = v
    return out


def tree_unflatten(flat):
    t = {}
    for k, v in flat.items():
        parts = k.split("/")
        d = t
        for q in parts[:-1]:
            d = d.setdefault(q, {})
        d[parts[-1]] = v
    return t


def params_to(t, dtype):
    return tree_map(lambda a: torch.as_tensor(np.asarray(a)).to(dtype), t)
