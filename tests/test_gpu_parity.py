"""GPU parity of the CUDA path against the CPU oracle on seeded synthetic molecules: energies,
forces and every parameter gradient; padded batches against the per-molecule unpadded oracle."""
import numpy as np
import pytest
import torch

from oracle import sake_oracle as O
from tests import synth

pytestmark = pytest.mark.gpu


def _oracle_params(p, dt=torch.float64):
    return O.tree_map(lambda t: t.detach().cpu().to(dt), p)


def _setup(H, depth, B, N, S, seed, engine, padded=False, n_min=2, update=True):
    import sake_b200
    h, x, mask, am = synth.molecules(seed, B, N, S, padded, n_min)
    model = sake_b200.DenseSAKEModel(hidden_features=H, out_features=1, depth=depth, update=update, engine=engine)
    dev = "cuda"
    ht, xt = torch.tensor(h, device=dev), torch.tensor(x, device=dev)
    variables = model.init(seed, ht, xt)
    # non-zero biases so that every bias path is exercised
    import sake_b200.layers as L
    flat = L.flatten_tree(variables["params"])
    g = torch.Generator().manual_seed(seed + 1)
    for k, t in flat.items():
        if k.endswith("bias"):
            t.add_(0.1 * torch.randn(t.shape, generator=g).to(dev))
    return model, variables["params"], ht, xt, (None if mask is None else torch.tensor(mask, device=dev)), \
        (None if am is None else torch.tensor(am, device=dev))


@pytest.mark.parametrize("engine", ["fp32", "auto", "f16x2"])
@pytest.mark.parametrize("H,depth,B,N", [(64, 2, 4, 21), (16, 3, 3, 9), (64, 1, 2, 70)])
def test_energy_forces_vs_oracle(engine, H, depth, B, N):
    if engine == "f16x2" and H != 64:
        pytest.skip("tcgen05 engines need H = 64")
    model, p, h, x, _, _ = _setup(H, depth, B, N, 8, 2666 + N, engine)
    e, f = model.energy_and_forces(p, h, x)
    po = _oracle_params(p)
    e0, f0 = O.energy_and_forces(po, h.cpu().double(), x.cpu().double())
    rel = ((e.cpu().double() - e0).abs() / e0.abs().clamp_min(1e-12)).max().item()
    ferr = (f.cpu().double() - f0).abs().max().item()
    assert rel < 1e-5, f"energy rel err {rel:.3e}"
    assert ferr < 1e-4, f"force abs err {ferr:.3e} (max |F| {f0.abs().max().item():.3e})"


@pytest.mark.parametrize("engine", ["fp32", "auto", "f16x2"])
def test_param_grads_vs_oracle(engine):
    import sake_b200.layers as L
    H, depth, B, N = 64, 2, 3, 12
    model, p, h, x, _, _ = _setup(H, depth, B, N, 6, 77, engine)
    flat = L.flatten_tree(p)
    for t in flat.values():
        t.requires_grad_(True)
    y = torch.tensor(np.random.default_rng(5).standard_normal(B).astype(np.float32), device="cuda")
    e = model.energy(p, h, x)
    loss = (e - y).abs().mean()                      # scripts/qm9/run.py:79-82
    grads = torch.autograd.grad(loss, list(flat.values()), allow_unused=True)
    po = _oracle_params(p)
    fo = O.tree_flatten(po)
    for t in fo.values():
        t.requires_grad_(True)
    e0 = O.energy(po, h.cpu().double(), x.cpu().double())
    loss0 = (e0 - y.cpu().double()).abs().mean()
    g0 = torch.autograd.grad(loss0, list(fo.values()), allow_unused=True)
    assert abs(loss.item() - loss0.item()) < 1e-5 * max(1.0, abs(loss0.item()))
    for (k, _), ga, gb in zip(flat.items(), grads, g0):
        if gb is None:
            assert ga is None or float(ga.abs().max()) == 0.0, k
            continue
        assert ga is not None, k
        scale = max(float(gb.abs().max()), 1e-6)
        err = float((ga.cpu().double() - gb).abs().max())
        assert err < 2e-3 * scale + 1e-6, f"{k}: err {err:.3e} scale {scale:.3e}"


@pytest.mark.parametrize("engine", ["fp32", "auto", "f16x2"])
def test_padded_batch_matches_unpadded_oracle(engine):
    """QM9-style padding (scripts/qm9/run.py:23-24,35): real atoms of every padded molecule match the
    oracle run on that molecule alone, unpadded and unmasked (sake/tests/test_mask.py:202-240)."""
    H, depth, B, N = 64, 2, 5, 13
    model, p, h, x, mask, am = _setup(H, depth, B, N, 5, 4242, engine, padded=True, n_min=3)
    e, f = model.energy_and_forces(p, h, x, mask=mask, atom_mask=am)
    assert torch.isfinite(e).all() and torch.isfinite(f).all()
    po = _oracle_params(p)
    for b in range(B):
        n = int(am[b].sum().item())
        e0, f0 = O.energy_and_forces(po, h[b, :n].cpu().double(), x[b, :n].cpu().double())
        assert abs(e[b].item() - e0.item()) < 1e-5 * max(1.0, abs(e0.item())), (b, e[b].item(), e0.item())
        assert (f[b, :n].cpu().double() - f0).abs().max().item() < 1e-4
        assert float(f[b, n:].abs().max()) == 0.0 if n < N else True


def test_equivariance():
    """E(3): h invariant, x equivariant (sake/tests/test_equivariance.py:3-45), on the CUDA path."""
    model, p, h, x, _, _ = _setup(64, 2, 2, 10, 4, 31, "auto")
    rng = np.random.default_rng(0)
    q, _ = np.linalg.qr(rng.standard_normal((3, 3)))
    Rm = torch.tensor(q.astype(np.float32), device="cuda")
    t = torch.tensor(rng.standard_normal((1, 3)).astype(np.float32), device="cuda")
    h0, x0, _ = model(p, h, x)
    h1, x1, _ = model(p, h, x @ Rm + t)
    assert torch.allclose(h0, h1, rtol=1e-4, atol=1e-4)
    assert torch.allclose(x0 @ Rm + t, x1, rtol=1e-4, atol=1e-4)


def test_error_paths():
    import sake_b200
    from sake_b200._lib import SakeError
    with pytest.raises(SakeError):               # only the closed-form cosine cutoff runs inside the kernels
        sake_b200.DenseSAKELayer(16, 16, cutoff=lambda d: d)
    layer = sake_b200.DenseSAKELayer(16, 16)
    h = torch.zeros(3, 16)
    x = torch.zeros(3, 3)
    with pytest.raises(SakeError):               # CPU tensors: no CPU fallback
        p = layer.init(0, h, x)
        layer.apply(p, h, x)


@pytest.mark.parametrize("engine", ["fp32", "auto"])
def test_runner_matches_autograd_path(engine):
    """The preallocated ModelRunner (bench / production path) equals the autograd-Function path:
    energies, forces, the L1 training loss and every parameter gradient."""
    import sake_b200
    import sake_b200.layers as L
    from sake_b200.runner import ModelRunner
    H, depth, B, N, S = 64, 2, 4, 11, 6
    model, p, h, x, mask, am = _setup(H, depth, B, N, S, 909, engine, padded=True, n_min=4)
    run = ModelRunner(model, p, B, N, S, masked=True, train=True)
    y = torch.tensor(np.random.default_rng(1).standard_normal(B).astype(np.float32), device="cuda")
    run.load_inputs(h, x, mask, am, y)
    e, f = run.energy_forces_step()
    e1, f1 = model.energy_and_forces(p, h, x, mask=mask, atom_mask=am)
    assert torch.allclose(e, e1, rtol=1e-5, atol=1e-5)
    assert torch.allclose(f, f1, rtol=1e-4, atol=1e-5)
    # training step: gradients before the Adam update are left in run.g
    flat = L.flatten_tree(p)
    for t in flat.values():
        t.requires_grad_(True)
    loss1 = (model.energy(p, h, x, mask=mask, atom_mask=am) - y).abs().mean()
    grads = torch.autograd.grad(loss1, list(flat.values()), allow_unused=True)
    before = run.flat_params.clone()
    loss = run.train_step()
    assert abs(loss.item() - loss1.item()) < 1e-5 * max(1.0, abs(loss1.item()))
    for (k, _), ga in zip(flat.items(), grads):
        gb = run.g[k]
        if ga is None:
            assert float(gb.abs().max()) == 0.0, k
            continue
        scale = max(float(ga.abs().max()), 1e-6)
        assert float((ga - gb).abs().max()) < 2e-3 * scale + 1e-7, k
    # Adam moved the parameters (first step: |delta| ~ lr for every non-zero gradient)
    delta = (run.flat_params - before).abs()
    assert float(delta.max()) <= 1.1e-3 and float(delta.max()) > 1e-4


def test_tcgen05_selftest():
    """Descriptor encodings, swizzled operand images and the TMEM load layout on the real device."""
    import ctypes as C
    from sake_b200 import _lib
    err = (C.c_float * 2)()
    rc = _lib.lib.sake_selftest_tcgen05(err, None)
    torch.cuda.synchronize()
    assert rc == 0, (_lib.lib.sake_last_error().decode(), err[0], err[1])
    assert err[0] < 1e-5 and err[1] < 5e-2


@pytest.mark.parametrize("engine,tol", [(2, 1e-4), (3, 1e-2)])
@pytest.mark.parametrize("P,xw,gw", [(64, 128, 64), (300, 256, 256), (1000, 64, 16), (5000, 256, 64), (777, 100, 128)])
def test_xtg_contraction_shapes(engine, tol, P, xw, gw):
    """out = X^T G (the weight-gradient contraction kernel) through sake_selftest_xtg: regular shapes (lean
    builder: whole 64-feature blocks, one narrow block) and irregular ones (generic builder), with the
    split-bf16 (engine 2) and the plain bf16 (engine 3) operand precision."""
    from sake_b200 import _lib
    gen = torch.Generator(device="cuda").manual_seed(P + xw)
    X = torch.randn(P, xw, device="cuda", generator=gen)
    G = torch.randn(P, gw, device="cuda", generator=gen)
    out = torch.zeros(xw, gw, device="cuda")
    rc = _lib.lib.sake_selftest_xtg(engine, P, xw, gw, X.data_ptr(), G.data_ptr(), out.data_ptr(), None)
    torch.cuda.synchronize()
    assert rc == 0, _lib.lib.sake_last_error().decode()
    ref = X.double().T @ G.double()
    err = (out.double() - ref).abs().max().item()
    assert err < tol * ref.abs().max().item(), (err, ref.abs().max().item())


@pytest.mark.parametrize("engine", ["tf32x3", "f16x2", "bf16"])
@pytest.mark.parametrize("B,N,padded", [(64, 29, True), (40, 21, False), (3, 200, False)])
def test_tc_engine_multi_tile_vs_generic(engine, B, N, padded):
    """More 128-pair tiles than SMs (persistent multi-tile pipeline, every ring/phase wrap) and rows
    split over several tiles (N > 128): tcgen05 engines against the generic fp32 CUDA engine."""
    import sake_b200
    import sake_b200.layers as L
    from sake_b200.runner import ModelRunner
    S, depth = 6, 2
    h, x, mask, am = synth.molecules(7 + N, B, N, S, padded, 5)
    y = np.random.default_rng(3).standard_normal(B).astype(np.float32)
    dev = "cuda"
    T = lambda a: None if a is None else torch.tensor(a, device=dev)
    outs = {}
    for eng in ("fp32", engine):
        model = sake_b200.DenseSAKEModel(hidden_features=64, out_features=1, depth=depth, engine=eng)
        params = model.init(5, T(h), T(x))["params"]
        run = ModelRunner(model, params, B, N, S, masked=padded, train=True)
        run.load_inputs(T(h), T(x), T(mask), T(am), T(y))
        e, f = run.energy_forces_step()
        e, f = e.clone(), f.clone()
        loss = run.train_step().clone()
        torch.cuda.synchronize()
        outs[eng] = (e, f, loss, run.flat_grads.clone(), {k: v.clone() for k, v in run.g.items()})
    e0, f0, l0, g0, gd0 = outs["fp32"]
    e1, f1, l1, g1, gd1 = outs[engine]
    etol, ftol, gtol = (1e-5, 1e-4, 2e-3) if engine in ("tf32x3", "f16x2") else (2e-2, 5e-2, 1e-1)
    assert torch.isfinite(e1).all() and torch.isfinite(f1).all() and torch.isfinite(g1).all()
    # relative to max(|E|, 0.1): a molecule whose atom terms (O(1) each) cancel to |E| ~ 0.03 carries the
    # absolute fp32 rounding of the sum (a few 1e-7), which is not a relative error of the engine
    assert ((e1 - e0).abs() / e0.abs().clamp_min(0.1)).max().item() < etol
    assert (f1 - f0).abs().max().item() < ftol * max(1.0, f0.abs().max().item() if engine == "bf16" else 1.0)
    for k in gd0:
        scale = max(float(gd0[k].abs().max()), 1e-6)
        assert float((gd1[k] - gd0[k]).abs().max()) < gtol * scale + 1e-7, k


@pytest.mark.parametrize("wscale", [3000.0, 1e-4])
@pytest.mark.parametrize("engine", ["tf32x3", "f16x2"])
def test_small_gradient_scale_and_large_features(engine, wscale):
    """Range robustness of the split-precision engines: edge features blown up to ~1e3 or shrunk to ~1e-4
    (the fp16-split engine normalises every operand row by an exact power of two) and cotangents scaled by 1e-6 (up-scaling
    path).  Saturated tanh makes this an ill-conditioned regime even for plain fp32 arithmetic, so the
    yardstick is the generic fp32 CUDA-core engine: against the fp64 oracle a split engine may be at
    most 10x worse than it (the split engines carry 21-22 mantissa bits per product against 24) (plus the stated 1e-5 / 1e-4 floors)."""
    import sake_b200
    B, N, S, depth = 8, 21, 6, 2
    h, x, mask, am = synth.molecules(123, B, N, S, False, 0)
    dev = "cuda"
    T = lambda a: None if a is None else torch.tensor(a, device=dev)
    errs = {}
    for eng in ("fp32", engine):
        model = sake_b200.DenseSAKEModel(hidden_features=64, out_features=1, depth=depth, engine=eng)
        params = model.init(3, T(h), T(x))["params"]
        params["d0"]["edge_model"]["mlp_out"]["layers_2"]["kernel"].mul_(wscale)
        params["d0"]["edge_model"]["mlp_out"]["layers_2"]["bias"].mul_(wscale)
        po = _oracle_params(params)
        e0, f0 = O.energy_and_forces(po, T(h).cpu().double(), T(x).cpu().double())
        fmax = max(1.0, f0.abs().max().item())
        for scale in (1.0, 1e-6):
            xx = T(x).requires_grad_(True)
            e = model.energy(params, T(h), xx)
            (g,) = torch.autograd.grad((e * scale).sum(), xx)
            assert torch.isfinite(g).all() and torch.isfinite(e).all()
            # relative to the batch's energy scale: single molecules whose atom terms cancel (|E| ~ 1e-3)
            # would turn absolute fp32 rounding into an arbitrary relative number
            erel = (e.cpu().double() - e0).abs().max().item() / max(1.0, e0.abs().max().item())
            ferr = (-g.cpu().double() / scale - f0).abs().max().item() / fmax
            errs[(eng, scale)] = (erel, ferr)
    for scale in (1.0, 1e-6):
        (er0, fr0), (er1, fr1) = errs[("fp32", scale)], errs[(engine, scale)]
        assert er1 < 10 * er0 + 1e-5, f"energy: {engine} {er1:.3e} vs fp32 engine {er0:.3e} (scale {scale})"
        assert fr1 < 10 * fr0 + 1e-4, f"forces: {engine} {fr1:.3e} vs fp32 engine {fr0:.3e} (scale {scale})"


@pytest.mark.parametrize("B,N,S,padded,n_min", [(256, 29, 10, True, 9), (1024, 63, 4, True, 20), (64, 200, 84, False, 0)])
def test_full_size_properties(B, N, S, padded, n_min):
    """BASELINE.json's full sizes (cfg2 / cfg3 / cfg4 shapes, depth 4, the default engine) through
    size-independent properties of the energy/force path, plus an oracle spot check:
      * forces of every molecule sum to zero (translation invariance of E);
      * energies are invariant and forces co-rotate under a random rotation + translation
        (sake/tests/test_equivariance.py, at batch scale);
      * four molecules of the batch, cut out and run unpadded through the fp64 oracle, agree
        (1e-5 relative energy, 1e-4 absolute force; sake/tests/test_mask.py:202-240 at full width)."""
    model, p, h, x, mask, am = _setup(64, 4, B, N, S, 2666, "auto", padded=padded, n_min=n_min)
    kw = {} if mask is None else {"mask": mask, "atom_mask": am}
    e, f = model.energy_and_forces(p, h, x, **kw)
    assert torch.isfinite(e).all() and torch.isfinite(f).all()
    fs = f.sum(1).abs().max().item()
    assert fs < 2e-3 * max(1.0, f.abs().max().item()), f"sum of forces {fs:.3e}"
    rng = np.random.default_rng(7)
    q, _ = np.linalg.qr(rng.standard_normal((3, 3)))
    Rm = torch.tensor(q.astype(np.float32), device="cuda")
    t = torch.tensor(rng.standard_normal((1, 1, 3)).astype(np.float32), device="cuda")
    x2 = x @ Rm + t
    if am is not None:
        x2 = x2 * am[..., None]                       # padded atoms stay at the origin (scripts/qm9/run.py:35)
    e2, f2 = model.energy_and_forces(p, h, x2, **kw)
    assert ((e2 - e).abs() / e.abs().clamp_min(1.0)).max().item() < 2e-5
    assert (f2 - f @ Rm).abs().max().item() < 2e-4 * max(1.0, f.abs().max().item())
    po = _oracle_params(p)
    for b in (0, 1, B // 2, B - 1):
        n = N if am is None else int(am[b].sum().item())
        e0, f0 = O.energy_and_forces(po, h[b, :n].cpu().double(), x[b, :n].cpu().double())
        assert abs(e[b].item() - e0.item()) < 1e-5 * max(1.0, abs(e0.item())), (b, e[b].item(), e0.item())
        assert (f[b, :n].cpu().double() - f0).abs().max().item() < 1e-4, b


@pytest.mark.parametrize("train", [False, True])
def test_cuda_graph_replay_matches_eager(train):
    """ModelRunner.capture(): replaying the recorded step gives bit-identical energies / forces / gradients
    (the library is enqueue-only on the caller's stream, SURVEY 8b 'CUDA-graph capturable')."""
    import sake_b200
    from sake_b200 import runner as R
    from sake_b200.init_params import _generator, init_model_params
    B, N, S = 16, 21, 6
    h, x, mask, am = synth.molecules(11, B, N, S, True, 5)
    model = sake_b200.DenseSAKEModel(hidden_features=64, out_features=1, depth=2, engine="auto")
    T = lambda a: torch.tensor(a, device="cuda")
    y = torch.randn(B, device="cuda")
    outs = []
    for use_graph in (False, True):
        run = R.ModelRunner(model, init_model_params(_generator(0), S, 64, 1, 2), B, N, S, masked=True, train=train)
        run.load_inputs(T(h), T(x), T(mask), T(am), y)
        if use_graph:
            assert run.capture() > 10
        if train:
            loss = run.train_step().clone()
            outs.append((loss, run.flat_grads.clone()))
        else:
            e, f = run.energy_forces_step()
            outs.append((e.clone(), f.clone()))
    if train:
        # gradients of the same step; a few weight gradients are accumulated with atomics (summation order
        # varies from run to run), so the comparison is to 1e-5 of the largest gradient, not bitwise
        assert torch.allclose(outs[0][0], outs[1][0], rtol=1e-6, atol=1e-7)
        gmax = outs[0][1].abs().max().item()
        assert (outs[0][1] - outs[1][1]).abs().max().item() < 1e-5 * gmax
    else:
        assert torch.equal(outs[0][0], outs[1][0]) and torch.equal(outs[0][1], outs[1][1])


@pytest.mark.parametrize("update,with_v,spatial,masked", [(True, True, True, True), (True, False, True, False),
                                                          (False, True, True, False), (True, True, False, True),
                                                          (False, False, True, True)])
def test_layer_variants_multi_tile_vs_generic(update, with_v, spatial, masked):
    """Every branch of the per-node tail (velocity gate, update on/off, spatial attention off, mask) on more
    than one 128-atom tile: the tcgen05 engine (tc_node.cu, tc_edge.cu, tc_mix.cu) against the generic fp32
    CUDA-core engine, outputs and input gradients (sake/layers.py:188-235; the reference tests these flags only
    for shapes, sake/tests/test_layers.py:27-36)."""
    import sake_b200
    B, N = 5, 31                                  # 155 atoms = 2 tiles, the second one partial
    g = torch.Generator(device="cuda").manual_seed(17)
    h = torch.randn(B, N, 64, device="cuda", generator=g)
    x = torch.randn(B, N, 3, device="cuda", generator=g) * 2.0
    v = torch.randn(B, N, 3, device="cuda", generator=g) if with_v else None
    mask = None
    if masked:
        n_real = torch.tensor([31, 20, 9, 31, 2], device="cuda")
        am = (torch.arange(N, device="cuda")[None, :] < n_real[:, None]).float()
        mask = am[:, :, None] * am[:, None, :]
        h, x = h * am[..., None], x * am[..., None]
        v = v * am[..., None] if v is not None else None
    outs = {}
    for eng in ("fp32", "auto"):
        layer = sake_b200.DenseSAKELayer(64, 64, update=update, use_spatial_attention=spatial, engine=eng)
        p = layer.init(3, h, x, v, mask)
        hh, xx = h.clone().requires_grad_(True), x.clone().requires_grad_(True)
        vv = v.clone().requires_grad_(True) if v is not None else None
        ho, xo, vo = layer.apply(p, hh, xx, vv, mask)
        s = (ho ** 2).sum() + (xo * 0.3).sum() + ((vo * vo).sum() if vo is not None else 0.0)
        ins = [hh, xx] + ([vv] if vv is not None else [])
        gr = torch.autograd.grad(s, ins, allow_unused=True)
        outs[eng] = [ho, xo] + ([vo] if vo is not None else []) + [t for t in gr if t is not None]
    for a, b in zip(outs["fp32"], outs["auto"]):
        assert torch.isfinite(b).all()
        scale = max(1.0, a.abs().max().item())
        assert (a - b).abs().max().item() < 2e-4 * scale, ((a - b).abs().max().item(), scale)
