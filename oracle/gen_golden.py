"""Generate tests/golden/*.npz by executing the UNMODIFIED reference source.

Run in the build container (needs /root/reference):   python oracle/gen_golden.py

The reference (`/root/reference/sake/*.py`) is imported as-is; its `jax` / `flax` imports resolve
to `oracle/jaxshim` (torch-backed restatement of those third-party calls — JAX is not installable
in this image).  Every number stored comes out of the reference's own layers.py / models.py /
flows.py code paths, in fp64 and fp32, with forces / parameter gradients from torch autograd
running through that same code.  The fixtures travel to the GPU box; /root/reference does not.
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "jaxshim"))
sys.path.insert(0, "/root/reference")

import jax  # noqa: E402  (the shim)
import sake  # noqa: E402  (the reference)

OUT = os.path.join(HERE, "..", "tests", "golden")


def flatten(t, prefix=""):
    out = {}
    for k, v in t.items():
        if isinstance(v, dict):
            out.update(flatten(v, prefix + k + "/"))
        else:
            out[prefix + k] = v
    return out


def unflatten(flat):
    t = {}
    for k, v in flat.items():
        parts = k.split("/")
        d = t
        for q in parts[:-1]:
            d = d.setdefault(q, {})
        d[parts[-1]] = v
    return t


def perturb(params, rng):
    """Make biases / RBF params non-trivial (flax inits biases to 0); values are fp32-exact."""
    flat = flatten(params)
    out = {}
    for k, v in flat.items():
        a = v.detach().double().numpy().copy()
        if k.endswith("bias"):
            a = a + 0.1 * rng.standard_normal(a.shape)
        elif k.endswith("means"):
            a = a + 0.005 * rng.standard_normal(a.shape)
        elif k.endswith("betas"):
            a = a * (1.0 + 0.05 * rng.standard_normal(a.shape))
        out[k] = a.astype(np.float32)
    return out


def as_params(flat32, dtype):
    return {"params": unflatten({k: torch.tensor(v.astype(np.float64)).to(dtype) for k, v in flat32.items()})}


def T(a, dtype):
    return None if a is None else torch.tensor(np.asarray(a, dtype=np.float64)).to(dtype)


def run_both(fn):
    res = {}
    for tag, dt in (("f64", torch.float64), ("f32", torch.float32)):
        jax.set_dtype(dt)
        for k, v in fn(dt).items():
            res[f"{tag}/{k}"] = v.detach().double().numpy() if isinstance(v, torch.Tensor) else np.asarray(v)
    jax.set_dtype(torch.float64)
    return res


def save(name, flat_params, inputs, outputs, meta):
    d = {}
    for k, v in flat_params.items():
        d["param:" + k] = v
    for k, v in inputs.items():
        if v is not None:
            d["in:" + k] = np.asarray(v)
    for k, v in outputs.items():
        d["out:" + k] = v
    for k, v in meta.items():
        d["meta:" + k] = np.asarray(v)
    path = os.path.join(OUT, name + ".npz")
    np.savez_compressed(path, **d)
    print("wrote", path, "%.1f KB" % (os.path.getsize(path) / 1024))


def coords(rng, shape):
    n = shape[-2]
    return (rng.standard_normal(shape) * 0.62 * n ** (1.0 / 3.0)).astype(np.float32)


def case_layer(name, H, N, batch, with_v, with_mask, seed, update=True, cutoff=None, he_features=0):
    """cutoff: None or (lower, upper) -> DenseSAKELayer(cutoff=partial(sake.utils.cosine_cutoff, lower=, upper=))"""
    import functools
    rng = np.random.default_rng(seed)
    jax.set_dtype(torch.float64)
    cut = None if cutoff is None else functools.partial(sake.utils.cosine_cutoff, lower=cutoff[0], upper=cutoff[1])
    model = sake.layers.DenseSAKELayer(H, H, update=update, cutoff=cut)
    shp = tuple(batch) + (N,)
    h = rng.uniform(size=shp + (H,)).astype(np.float32)
    x = coords(rng, shp + (3,))
    v = rng.standard_normal(shp + (3,)).astype(np.float32) if with_v else None
    n_real = N - 1 if with_mask else N
    mask = None
    if with_mask:
        m = np.concatenate([np.ones(n_real), np.zeros(N - n_real)]).astype(np.float32)
        mask = np.broadcast_to(m[None, :] * m[:, None], shp + (N,)).copy()
    he = rng.standard_normal(shp + (N, he_features)).astype(np.float32) if he_features else None
    init = model.init(jax.random.PRNGKey(seed), T(h, torch.float64), T(x, torch.float64),
                      T(v, torch.float64), T(mask, torch.float64), T(he, torch.float64))
    flat = perturb(init["params"], rng)

    def fn(dt):
        p = as_params(flat, dt)
        xx = T(x, dt).requires_grad_(True)
        hh = T(h, dt).requires_grad_(True)
        hee = T(he, dt).requires_grad_(True) if he is not None else None
        ho, xo, vo = model.apply(p, hh, xx, T(v, dt), T(mask, dt), hee)
        ho, xo, vo = ho[..., :n_real, :], xo[..., :n_real, :], (vo[..., :n_real, :] if vo is not None else None)
        # a scalar that touches every output, for gradient pinning
        s = (ho ** 2).sum() + (xo * 0.3).sum() + ((vo * vo).sum() if (vo is not None and update) else 0.0)
        if hee is not None:
            gx, gh, ghe = torch.autograd.grad(s, [xx, hh, hee])
        else:
            gx, gh = torch.autograd.grad(s, [xx, hh])
        out = {"h": ho, "x": xo, "scalar": s, "grad_x": gx[..., :n_real, :], "grad_h": gh[..., :n_real, :]}
        if hee is not None:
            out["grad_he"] = ghe
        if vo is not None:
            out["v"] = vo
        return out

    meta = {"H": H, "N": N, "n_real": n_real, "update": int(update), "kind": "layer"}
    if cutoff is not None:
        meta["cutoff"] = np.asarray(cutoff, dtype=np.float64)
    save(name, flat, {"h": h, "x": x, "v": v, "mask": mask, "he": he}, run_both(fn), meta)


def case_model(name, H, F_in, depth, N, batch, with_v, seed, out_features=1, update=True, grad_keys=()):
    rng = np.random.default_rng(seed)
    jax.set_dtype(torch.float64)
    model = sake.models.DenseSAKEModel(hidden_features=H, out_features=out_features, depth=depth, update=update)
    shp = tuple(batch) + (N,)
    z = rng.integers(0, F_in, shp)
    h = np.eye(F_in, dtype=np.float32)[z]
    x = coords(rng, shp + (3,))
    v = rng.standard_normal(shp + (3,)).astype(np.float32) if with_v else None
    init = model.init(jax.random.PRNGKey(seed), T(h, torch.float64), T(x, torch.float64), T(v, torch.float64))
    flat = perturb(init["params"], rng)

    def fn(dt):
        p = as_params(flat, dt)
        leaves = flatten(p["params"])
        for k in grad_keys:
            leaves[k].requires_grad_(True)
        xx = T(x, dt).requires_grad_(True)
        ho, xo, vo = model.apply(p, T(h, dt), xx, T(v, dt))
        e = ho.sum(dim=(-1, -2))                       # scripts/md17/run.py:46-52
        grads = torch.autograd.grad(e.sum(), [xx] + [leaves[k] for k in grad_keys], retain_graph=True)
        out = {"h": ho, "x": xo, "energy": e, "forces": -grads[0]}
        if vo is not None:
            out["v"] = vo
        for k, g in zip(grad_keys, grads[1:]):
            out["grad:" + k] = g
        return out

    save(name, flat, {"h": h, "x": x, "v": v}, run_both(fn),
         {"H": H, "N": N, "depth": depth, "kind": "model",
          "update": np.asarray(update, dtype=np.int64)})


def case_flow(name, H, depth, mp_depth, N, D, B, seed, grad_keys=()):
    rng = np.random.default_rng(seed)
    jax.set_dtype(torch.float64)
    model = sake.flows.AugmentedFlowModel(depth=depth, mp_depth=mp_depth, hidden_features=H)
    h = np.zeros((B, N, 2), dtype=np.float32)           # scripts/lj13_aug/run.py:34
    x = rng.standard_normal((B, N, D)).astype(np.float32)
    x -= x.mean(-2, keepdims=True)
    v = rng.standard_normal((B, N, D)).astype(np.float32)
    v -= v.mean(-2, keepdims=True)
    init = model.init(jax.random.PRNGKey(seed), T(h, torch.float64), T(x, torch.float64), T(v, torch.float64))
    flat = perturb(init["params"], rng)

    def fn(dt):
        p = as_params(flat, dt)
        xf, vf, ldf = model.apply(p, T(h, dt), T(x, dt), T(v, dt))
        leaves = flatten(p["params"])
        for k in grad_keys:
            leaves[k].requires_grad_(True)
        xb, vb, ldb = model.apply(p, T(h, dt), T(x, dt), T(v, dt), method=model.f_backward)
        out = {"fwd_x": xf, "fwd_v": vf, "fwd_logdet": ldf, "bwd_x": xb, "bwd_v": vb, "bwd_logdet": ldb}
        if grad_keys:
            # the likelihood loss of scripts/lj13_aug/run.py:39-43 and its parameter gradient
            prior = sake.flows.CenteredGaussian
            loss = (-prior.log_prob(xb) - prior.log_prob(vb) + ldb).mean()
            grads = torch.autograd.grad(loss, [leaves[k] for k in grad_keys])
            out["loss"] = loss
            for k, g in zip(grad_keys, grads):
                out["grad:" + k] = g
        return out

    save(name, flat, {"h": h, "x": x, "v": v}, run_both(fn),
         {"H": H, "N": N, "D": D, "depth": depth, "mp_depth": mp_depth, "kind": "flow"})


def main_round2():
    """Fixtures added in round 2 (the round-1 files are left byte-identical): hidden_features = 64 flows, i.e. the
    shape the tcgen05 engine serves, LJ13-like (13 atoms, D = 3) and DW4-like (4 atoms, D = 2), with the
    likelihood loss and its gradient for one leaf of every kind of sub-module."""
    os.makedirs(OUT, exist_ok=True)
    torch.manual_seed(0)
    gk = ("xv_0/sake_model/d0/x_mixing/layers_0/kernel", "vx_0/sake_model/d1/edge_model/mlp_out/layers_2/kernel",
          "xv_0/scale_mlp/layers_0/kernel", "vx_0/sake_model/d1/v_mixing/kernel",
          "xv_0/sake_model/d1/node_mlp/layers_2/bias", "vx_0/sake_model/embedding_in/kernel")
    case_flow("flow_h64_n13_d3", 64, 1, 2, 13, 3, 4, 13, grad_keys=gk)
    case_flow("flow_h64_n4_d2", 64, 1, 2, 4, 2, 5, 14, grad_keys=gk)
    # DenseSAKELayer(cutoff=cosine_cutoff) (sake/layers.py:172-176, sake/utils.py:10-26): tcgen05 shape and generic shape
    case_layer("layer_h64_n9_b2_cutoff", 64, 9, (2,), True, False, 21, cutoff=(0.0, 5.0))
    case_layer("layer_h16_n6_cutoff", 16, 6, (), False, False, 22, cutoff=(0.5, 4.0))
    # edge features he (sake/layers.py:201-202), tcgen05 shape and generic shape
    case_layer("layer_h64_n9_b2_he5", 64, 9, (2,), True, False, 23, he_features=5)
    case_layer("layer_h16_n6_he3", 16, 6, (), False, False, 24, he_features=3)


def main():
    os.makedirs(OUT, exist_ok=True)
    torch.manual_seed(0)
    # shapes of sake/tests/test_layers.py:25-36 (unbatched, v=None)
    case_layer("layer_h16_n5", 16, 5, (), False, False, 2666)
    # masked + velocity, batch of 2; padded atom appended as in sake/tests/test_mask.py:202-220
    case_layer("layer_h16_n6_mask_v", 16, 6, (2,), True, True, 1984)
    # odd hidden size used by sake/tests/test_equivariance.py:8
    case_layer("layer_h7_n5_v", 7, 5, (), True, False, 2046)
    case_layer("layer_h64_n9_b2", 64, 9, (2,), True, False, 7)
    case_layer("layer_h16_n5_noupdate", 16, 5, (), False, False, 11, update=False)
    # sake/tests/test_model.py shapes; energy/forces as scripts/md17/run.py
    case_model("model_h16_d4_n5", 16, 16, 4, 5, (), False, 2666, out_features=16)
    case_model("model_h64_d2_n12_b2", 64, 8, 2, 12, (2,), False, 1234,
               grad_keys=("d0/x_mixing/layers_0/kernel", "d1/edge_model/mlp_out/layers_0/kernel",
                          "d0/edge_model/kernel/means", "embedding_in/kernel",
                          "d1/semantic_attention_mlp/layers_0/kernel"))
    case_model("model_h32_d3_n7_v_updlist", 32, 5, 3, 7, (3,), True, 99, update=[False, True, True],
               grad_keys=("d1/velocity_mlp/layers_0/kernel", "d1/v_mixing/kernel"))
    # flows: sake/tests/test_augmented_flow.py + LJ13 / DW4 shapes (small)
    case_flow("flow_h16_n4_d3", 16, 2, 2, 4, 3, 3, 5)
    case_flow("flow_h16_n4_d2", 16, 2, 2, 4, 2, 3, 6)


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "round2":
        main_round2()
    else:
        main()
        main_round2()
