"""GPU parity against the golden vectors (reference source run under oracle/jaxshim), through the
flax-style host API -> C ABI -> CUDA kernels.  Tolerances are the north-star ones: fp32 energies
1e-5 relative, forces 1e-4 absolute; raw layer outputs 1e-4 relative / 2e-5 absolute."""
import numpy as np
import pytest
import torch

from tests import golden_util as G

pytestmark = pytest.mark.gpu

ENGINES = ["fp32", "auto"]


def _dev(a):
    return None if a is None else torch.tensor(np.asarray(a), dtype=torch.float32, device="cuda")


def _params(g):
    import sake_b200.layers as L
    return L.unflatten_tree({k: _dev(v) for k, v in g["params"].items()})


def _close(a, b, rtol, atol, what):
    a = a.detach().double().cpu().numpy() if isinstance(a, torch.Tensor) else np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    assert a.shape == b.shape, (what, a.shape, b.shape)
    err = np.abs(a - b).max() if a.size else 0.0
    assert np.allclose(a, b, rtol=rtol, atol=atol), f"{what}: max abs err {err:.3e} (max |ref| {np.abs(b).max():.3e})"


@pytest.mark.parametrize("engine", ENGINES)
@pytest.mark.parametrize("name", G.names("layer"))
def test_layer_golden(name, engine):
    import sake_b200
    g = G.load(name)
    H = int(g["meta"]["H"])
    n_real = int(g["meta"]["n_real"])
    update = bool(int(g["meta"]["update"]))
    cutoff = None
    if "cutoff" in g["meta"]:                      # sake/layers.py:172-176 with sake.utils.cosine_cutoff
        import functools
        lo, hi = (float(c) for c in g["meta"]["cutoff"])
        cutoff = functools.partial(sake_b200.utils.cosine_cutoff, lower=lo, upper=hi)
    layer = sake_b200.DenseSAKELayer(H, H, update=update, engine=engine, cutoff=cutoff)
    p = _params(g)
    h = _dev(g["in"]["h"]).requires_grad_(True)
    x = _dev(g["in"]["x"]).requires_grad_(True)
    v = _dev(g["in"].get("v"))
    mask = _dev(g["in"].get("mask"))
    he = _dev(g["in"].get("he"))                       # edge features (sake/layers.py:201-202), round-2 fixtures
    if he is not None:
        he.requires_grad_(True)
    ho, xo, vo = layer.apply({"params": p}, h, x, v, mask, he)
    ho, xo = ho[..., :n_real, :], xo[..., :n_real, :]
    _close(ho, g["out"]["f32/h"], 1e-4, 2e-5, "h")
    _close(xo, g["out"]["f32/x"], 1e-4, 2e-5, "x")
    if vo is not None:
        vo = vo[..., :n_real, :]
        _close(vo, g["out"]["f32/v"], 1e-4, 2e-5, "v")
    # compare against the fp64 run of the reference too (truth)
    _close(ho, g["out"]["f64/h"], 1e-4, 2e-5, "h vs f64")
    if np.isnan(g["out"]["f32/grad_x"]).any():
        return     # the reference's own gradients are NaN for padded inputs (layers.py:178-180)
    s = (ho ** 2).sum() + (xo * 0.3).sum() + ((vo * vo).sum() if (vo is not None and update) else 0.0)
    if he is not None:
        gx, gh, ghe = torch.autograd.grad(s, [x, h, he])
        _close(ghe, g["out"]["f64/grad_he"], 1e-3, 1e-4, "grad_he")
    else:
        gx, gh = torch.autograd.grad(s, [x, h])
    _close(gx[..., :n_real, :], g["out"]["f64/grad_x"], 1e-3, 1e-4, "grad_x")
    _close(gh[..., :n_real, :], g["out"]["f64/grad_h"], 1e-3, 1e-4, "grad_h")


@pytest.mark.parametrize("engine", ENGINES)
@pytest.mark.parametrize("name", G.names("model"))
def test_model_golden(name, engine):
    import sake_b200
    import sake_b200.layers as L
    g = G.load(name)
    H, depth = int(g["meta"]["H"]), int(g["meta"]["depth"])
    upd = g["meta"]["update"]
    update = bool(upd) if upd.ndim == 0 else [bool(u) for u in upd]
    out_features = g["out"]["f32/h"].shape[-1]
    model = sake_b200.DenseSAKEModel(hidden_features=H, out_features=out_features, depth=depth, update=update,
                                     engine=engine)
    p = _params(g)
    flat = L.flatten_tree(p)
    gkeys = [k[len("f64/grad:"):] for k in g["out"] if k.startswith("f64/grad:")]
    for k in gkeys:
        flat[k].requires_grad_(True)
    x = _dev(g["in"]["x"]).requires_grad_(True)
    ho, xo, vo = model.apply({"params": p}, _dev(g["in"]["h"]), x, _dev(g["in"].get("v")))
    _close(ho, g["out"]["f32/h"], 1e-4, 2e-5, "h")
    _close(xo, g["out"]["f32/x"], 1e-4, 2e-5, "x")
    e = ho.sum(dim=(-1, -2))
    ref_e = g["out"]["f64/energy"]
    rel = np.abs(e.detach().double().cpu().numpy() - ref_e) / np.maximum(np.abs(ref_e), 1e-12)
    assert rel.max() < 1e-5, f"energy rel err {rel.max():.3e}"
    grads = torch.autograd.grad(e.sum(), [x] + [flat[k] for k in gkeys])
    _close(-grads[0], g["out"]["f64/forces"], 0.0, 1e-4, "forces")
    for k, gr in zip(gkeys, grads[1:]):
        ref = g["out"]["f64/grad:" + k]
        _close(gr, ref, 1e-3, 1e-4 * max(1.0, float(np.abs(ref).max())), "grad " + k)


@pytest.mark.parametrize("name", G.names("flow"))
def test_flow_golden(name):
    """H = 16 fixtures run on the generic fp32 engine, the H = 64 ones (round 2: LJ13-like 13 atoms / D = 3 and
    DW4-like 4 atoms / D = 2) on the tcgen05 engine (`auto`), including the likelihood loss of
    scripts/lj13_aug/run.py:39-43 and its parameter gradient."""
    import sake_b200
    import sake_b200.layers as L
    g = G.load(name)
    m = g["meta"]
    flow = sake_b200.flows.AugmentedFlowModel(depth=int(m["depth"]), mp_depth=int(m["mp_depth"]),
                                              hidden_features=int(m["H"]))
    p = _params(g)
    gkeys = [k[len("f64/grad:"):] for k in g["out"] if k.startswith("f64/grad:")]
    if gkeys:
        flat = L.flatten_tree(p)
        for k in gkeys:
            flat[k].requires_grad_(True)
        hh, xx, vv = (_dev(g["in"][k]) for k in ("h", "x", "v"))
        xb, vb, ldb = flow.apply({"params": p}, hh, xx, vv, method="f_backward")
        CG = sake_b200.flows.CenteredGaussian
        loss = (-CG.log_prob(xb) - CG.log_prob(vb) + ldb).mean()
        ref_loss = float(g["out"]["f64/loss"])
        assert abs(loss.item() - ref_loss) < 1e-5 * max(1.0, abs(ref_loss)), (loss.item(), ref_loss)
        grads = torch.autograd.grad(loss, [flat[k] for k in gkeys])
        for k, gr in zip(gkeys, grads):
            ref = g["out"]["f64/grad:" + k]
            _close(gr, ref, 1e-3, 2e-3 * float(np.abs(ref).max()) + 1e-7, "grad " + k)
        p = _params(g)
    h, x, v = (_dev(g["in"][k]) for k in ("h", "x", "v"))
    xf, vf, ld = flow.apply({"params": p}, h, x, v)
    _close(xf, g["out"]["f64/fwd_x"], 1e-4, 5e-5, "fwd_x")
    _close(vf, g["out"]["f64/fwd_v"], 1e-4, 5e-5, "fwd_v")
    _close(ld, g["out"]["f64/fwd_logdet"], 1e-4, 5e-5, "fwd_logdet")
    xb, vb, ldb = flow.apply({"params": p}, h, x, v, method="f_backward")
    _close(xb, g["out"]["f64/bwd_x"], 1e-4, 5e-5, "bwd_x")
    _close(vb, g["out"]["f64/bwd_v"], 1e-4, 5e-5, "bwd_v")
    # invertibility (sake/tests/test_augmented_flow.py:47-63)
    x2, v2, _ = flow.apply({"params": p}, h, xf, vf, method="f_backward")
    _close(x2, x.cpu().numpy(), 1e-4, 5e-5, "inv_x")
    _close(v2, v.cpu().numpy(), 1e-4, 5e-5, "inv_v")
