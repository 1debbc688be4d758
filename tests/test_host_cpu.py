"""CPU-side checks: the C-ABI library loads and exports every symbol include/sake_b200.h declares,
argument validation works without a GPU, and the host parameter trees match the flax naming."""
import ctypes as C
import os
import re

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_header_symbols():
    from sake_b200 import _lib
    hdr = open(os.path.join(ROOT, "include", "sake_b200.h")).read()
    declared = set(re.findall(r"\b(sake_[a-z0-9_]+)\s*\(", hdr))
    declared.discard("sake_stream_t")
    assert declared, "no declarations parsed"
    for name in sorted(declared):
        assert hasattr(_lib.lib, name), f"{name} declared in the header but not exported"
    assert set(_lib.EXPORTS) == declared
    assert "sm_100a" in _lib.version()


def test_argument_validation_without_gpu():
    from sake_b200 import _lib
    d = _lib.SakeDims(1, 0, 64, 4, 50, 0, 0, 0)       # N = 0 is invalid
    assert _lib.lib.sake_layer_saved_bytes(C.byref(d)) == 0
    assert _lib.lib.sake_resolve_engine(C.byref(d)) == -1
    assert b"bad dims" in _lib.lib.sake_last_error()
    d = _lib.SakeDims(2, 5, 16, 4, 50, _lib.SAKE_UPDATE, _lib.ENGINE_TF32X3, 0)   # tcgen05 needs H=64
    assert _lib.lib.sake_resolve_engine(C.byref(d)) == -2
    d = _lib.SakeDims(2, 5, 16, 4, 50, _lib.SAKE_UPDATE, _lib.ENGINE_AUTO, 0)
    assert _lib.lib.sake_resolve_engine(C.byref(d)) == _lib.ENGINE_FP32
    assert _lib.lib.sake_layer_saved_bytes(C.byref(d)) > 0
    rc = _lib.lib.sake_layer_fwd(C.byref(d), None, None, None, None, None, None, None, None, None, None, None, 0, None,
                                 0, None)
    assert rc == -1
    # n_rbf: the tcgen05 edge kernel's K = 64 operand holds the RBF channels + 3 extra columns, and an engine is
    # tcgen05 for all kernels of a layer or for none (they share G8-layout buffers): beyond 58 AUTO -> generic engine
    d = _lib.SakeDims(2, 5, 64, 4, 58, 0, _lib.ENGINE_AUTO, 0)
    assert _lib.lib.sake_resolve_engine(C.byref(d)) == _lib.ENGINES["f16x2"]
    d = _lib.SakeDims(2, 5, 64, 4, 60, 0, _lib.ENGINE_AUTO, 0)
    assert _lib.lib.sake_resolve_engine(C.byref(d)) == _lib.ENGINE_FP32
    d = _lib.SakeDims(2, 5, 64, 4, 60, 0, _lib.ENGINES["f16x2"], 0)
    assert _lib.lib.sake_resolve_engine(C.byref(d)) == -2 and b"n_rbf" in _lib.lib.sake_last_error()
    # the saved buffer covers the G8 padding (rows rounded up to 8) and the weight images
    d = _lib.SakeDims(1, 3, 64, 4, 50, 0, _lib.ENGINE_AUTO, 0)
    assert _lib.lib.sake_layer_saved_bytes(C.byref(d)) > 2 * 1024 * 1024
    # ragged tables: size query works without a GPU, bad arguments are refused before any launch
    assert _lib.lib.sake_ragged_bytes(256, 29) > 256 * 29 * 32
    assert _lib.lib.sake_ragged_prepare(4, 200, None, None, 0, None) == -1
    assert _lib.lib.sake_ragged_bytes(4, 0) == 0


def test_param_tree_matches_flax_names():
    """Same leaves / shapes as the reference's init under the shim."""
    from tests import golden_util as G
    import sake_b200
    import sake_b200.layers as L
    g = G.load("model_h32_d3_n7_v_updlist")
    model = sake_b200.DenseSAKEModel(hidden_features=32, out_features=1, depth=3, update=[False, True, True])
    h = torch.zeros(3, 7, 5)
    x = torch.zeros(3, 7, 3)
    ours = L.flatten_tree(model.init(0, h, x, v=torch.zeros(3, 7, 3))["params"])
    ref = g["params"]
    assert set(ours) == set(ref), (set(ours) ^ set(ref))
    for k in ref:
        assert tuple(ours[k].shape) == tuple(ref[k].shape), k
    g = G.load("layer_h16_n5")
    layer = sake_b200.DenseSAKELayer(16, 16)
    ours = L.flatten_tree(layer.init(0, torch.zeros(5, 16), torch.zeros(5, 3))["params"])
    assert set(ours) == set(g["params"])


def test_functional_shapes():
    # sake/tests/test_functional.py:3-29
    import sake_b200
    x = torch.randn(5, 3)
    r = sake_b200.functional.get_x_minus_xt(x)
    assert r.shape == (5, 5, 3)
    assert sake_b200.functional.get_x_minus_xt_norm(r).shape == (5, 5, 1)
    assert sake_b200.functional.get_h_cat_ht(x).shape == (5, 5, 6)


def test_layer_sub_methods_match_oracle():
    """model.apply(params, ..., method=model.<sub-method>) as the reference's mask tests call it
    (sake/tests/test_mask.py:80,107,155,189): the host-side mirrors against the oracle's restatement, on CPU
    tensors (they are plain torch; the fused kernels never call them)."""
    from oracle import sake_oracle as O
    import sake_b200
    import sake_b200.functional as F
    torch.manual_seed(0)
    layer = sake_b200.DenseSAKELayer(16, 16)
    h = torch.rand(5, 16, dtype=torch.float64)
    x = torch.randn(5, 3, dtype=torch.float64)
    p = layer.init(3, h.float(), x.float(), v=torch.zeros(5, 3))
    p = {"params": O.tree_map(lambda t: t.double(), p["params"])}
    po = p["params"]
    r = F.get_x_minus_xt(x)
    n = F.get_x_minus_xt_norm(r)
    hc = F.get_h_cat_ht(h)
    e = layer.apply(p, hc, n, method=layer.edge_model)
    assert torch.allclose(e, O.edge_model(po["edge_model"], hc, n), rtol=1e-12, atol=1e-12)
    m = torch.cat([torch.ones(4), torch.zeros(1)]).double()
    mask = m[None, :] * m[:, None]
    for mk in (None, mask):
        sem = layer.apply(p, e, mask=mk, method=layer.semantic_attention)
        assert torch.allclose(sem, O.semantic_attention(po, e, mask=mk), rtol=1e-12, atol=1e-12)
        euc, sem2, comb = layer.apply(p, n, e, mask=mk, method="combined_attention")
        assert euc == 1.0 and torch.equal(sem2, sem)
        assert torch.allclose(comb, O.combined_attention(po, n, e, mask=mk), rtol=1e-12, atol=1e-12)
        hea = (e.unsqueeze(-1) * comb.unsqueeze(-2)).reshape(*e.shape[:-1], -1)
        hcomb, combs = layer.apply(p, hea, r, n, mask=mk, method=layer.spatial_attention)
        h0, c0 = O.spatial_attention(po, hea, r, n, mask=mk)
        assert torch.allclose(hcomb, h0, rtol=1e-12, atol=1e-12) and torch.allclose(combs, c0, rtol=1e-12, atol=1e-12)
        he = layer.apply(p, hea, mask=mk, method=layer.aggregate)
        assert he.shape == (5, 64)
        hn = layer.apply(p, h, he, hcomb, method=layer.node_model)
        assert hn.shape == (5, 16)
    v = layer.apply(p, torch.ones(5, 3).double(), h, method=layer.velocity_model)
    assert v.shape == (5, 3)
    from sake_b200._lib import SakeError
    with pytest.raises(SakeError):        # stale in the reference too: DenseSAKELayer has no euclidean_attention
        layer.apply(p, n, method="euclidean_attention")
