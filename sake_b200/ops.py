"""Thin torch-tensor wrappers over the C ABI (torch is plumbing: device memory + streams) and the
autograd Functions that play the role `jax.custom_vjp` plays on the JAX side: forward =
`sake_layer_fwd`, backward = `sake_layer_bwd`."""
import ctypes as C

import torch

from . import _lib
from ._lib import lib, check

# (path inside the flax layer tree, SakeLayerParams field)  — SURVEY Appendix C
LAYER_LEAVES = (
    ("edge_model/kernel/means", "rbf_means"),
    ("edge_model/kernel/betas", "rbf_betas"),
    ("edge_model/mlp_in/kernel", "mlp_in_kernel"),
    ("edge_model/mlp_in/bias", "mlp_in_bias"),
    ("edge_model/mlp_out/layers_0/kernel", "mlp_out0_kernel"),
    ("edge_model/mlp_out/layers_0/bias", "mlp_out0_bias"),
    ("edge_model/mlp_out/layers_2/kernel", "mlp_out2_kernel"),
    ("edge_model/mlp_out/layers_2/bias", "mlp_out2_bias"),
    ("semantic_attention_mlp/layers_0/kernel", "sem_kernel"),
    ("semantic_attention_mlp/layers_0/bias", "sem_bias"),
    ("x_mixing/layers_0/kernel", "x_mixing_kernel"),
    ("post_norm_mlp/layers_0/kernel", "post0_kernel"),
    ("post_norm_mlp/layers_0/bias", "post0_bias"),
    ("post_norm_mlp/layers_2/kernel", "post2_kernel"),
    ("post_norm_mlp/layers_2/bias", "post2_bias"),
    ("node_mlp/layers_0/kernel", "node0_kernel"),
    ("node_mlp/layers_0/bias", "node0_bias"),
    ("node_mlp/layers_2/kernel", "node2_kernel"),
    ("node_mlp/layers_2/bias", "node2_bias"),
    ("v_mixing/kernel", "v_mixing_kernel"),
    ("velocity_mlp/layers_0/kernel", "vel0_kernel"),
    ("velocity_mlp/layers_0/bias", "vel0_bias"),
    ("velocity_mlp/layers_2/kernel", "vel2_kernel"),
)
FIELD_OF = dict(LAYER_LEAVES)
PATH_OF = {f: p for p, f in LAYER_LEAVES}


def required_leaves(update, has_v, spatial=True):
    """Leaves DenseSAKELayer.__call__ reads for this configuration (mirrors check_leaves in csrc/api.cu)."""
    opt = {"v_mixing/kernel": update and spatial}
    for q in ("velocity_mlp/layers_0/kernel", "velocity_mlp/layers_0/bias", "velocity_mlp/layers_2/kernel"):
        opt[q] = update and has_v
    return [p for p, _ in LAYER_LEAVES if opt.get(p, True)]


def _ptr(t):
    return None if t is None else C.c_void_p(t.data_ptr())


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _f32c(t, name):
    if t is None:
        return None
    if not t.is_cuda:
        raise _lib.SakeError(f"{name} must be a CUDA tensor (sake_b200 has no CPU path)")
    if t.dtype != torch.float32:
        t = t.float()
    return t.contiguous()


def make_dims(B, N, H, A, K, update, has_v, has_mask, spatial=True, engine="auto", cutoff=None):
    """cutoff: None or (lower, upper) of sake.utils.cosine_cutoff (sake/utils.py:10-26)."""
    flags = ((_lib.SAKE_UPDATE if update else 0) | (_lib.SAKE_HAS_V if has_v else 0) |
             (_lib.SAKE_HAS_MASK if has_mask else 0) | (0 if spatial else _lib.SAKE_NO_SPATIAL) |
             (_lib.SAKE_COSINE_CUTOFF if cutoff is not None else 0))
    eng = _lib.ENGINES[engine] if isinstance(engine, str) else int(engine)
    lo, hi = (0.0, 5.0) if cutoff is None else (float(cutoff[0]), float(cutoff[1]))
    return _lib.SakeDims(int(B), int(N), int(H), int(A), int(K), flags, eng, 0, lo, hi)


def resolve_engine(dims):
    e = lib.sake_resolve_engine(C.byref(dims))
    if e < 0:
        check(e, "sake_resolve_engine")
    return _lib.ENGINE_NAMES[e]


def params_struct(flat, cls=_lib.SakeLayerParams):
    """flat: {flax path -> tensor}. Missing leaves become NULL."""
    s = cls()
    keep = []
    for path, field in LAYER_LEAVES:
        t = flat.get(path)
        if t is not None:
            keep.append(t)
            setattr(s, field, t.data_ptr())
    return s, keep


def saved_bytes(dims):
    return int(lib.sake_layer_saved_bytes(C.byref(dims)))


def scratch_bytes(dims, for_backward, with_grads):
    return int(lib.sake_layer_scratch_bytes(C.byref(dims), int(for_backward), int(with_grads)))


def _buf(nbytes, device):
    return torch.empty(max(int(nbytes), 256), dtype=torch.uint8, device=device)


def _pair_struct(pair):
    """pair: None or (u, p[, g_u, g_p]) CUDA tensors -> SakePairTerms (or None)"""
    if pair is None:
        return None
    t = list(pair) + [None] * (4 - len(pair))
    s = _lib.SakePairTerms()
    for name, ten in zip(("u", "p", "g_u", "g_p"), t):
        if ten is not None:
            setattr(s, name, ten.data_ptr())
    return C.byref(s)


def layer_fwd_raw(dims, pstruct, h, x, v, mask, h_out, x_out, v_out, saved, scratch, ragged=None, pair=None):
    rc = lib.sake_layer_fwd(C.byref(dims), C.byref(pstruct), _ptr(h), _ptr(x), _ptr(v), _ptr(mask), _ptr(ragged),
                            _pair_struct(pair), _ptr(h_out), _ptr(x_out), _ptr(v_out), _ptr(saved), saved.numel(),
                            _ptr(scratch), 0 if scratch is None else scratch.numel(), _stream())
    check(rc, "sake_layer_fwd")


def layer_bwd_raw(dims, pstruct, h, x, v, mask, saved, dh_out, dx_out, dv_out, dh, dx, dv, gstruct, scratch,
                  ragged=None, pair=None):
    rc = lib.sake_layer_bwd(C.byref(dims), C.byref(pstruct), _ptr(h), _ptr(x), _ptr(v), _ptr(mask), _ptr(ragged),
                            _pair_struct(pair), _ptr(saved), saved.numel(), _ptr(dh_out), _ptr(dx_out), _ptr(dv_out),
                            _ptr(dh), _ptr(dx), _ptr(dv),
                            None if gstruct is None else C.byref(gstruct),
                            _ptr(scratch), scratch.numel(), _stream())
    check(rc, "sake_layer_bwd")


def dense_fwd_raw(x, kernel, bias, y, act, ragged=None):
    rows = x.numel() // kernel.shape[0]
    check(lib.sake_dense_fwd(rows, kernel.shape[0], kernel.shape[1], int(act), _ptr(x), _ptr(kernel),
                             _ptr(bias), _ptr(y), _ptr(ragged), _stream()), "sake_dense_fwd")


def dense_bwd_raw(x, kernel, bias, dy, dx, dkernel, dbias, act, ragged=None):
    rows = x.numel() // kernel.shape[0]
    check(lib.sake_dense_bwd(rows, kernel.shape[0], kernel.shape[1], int(act), _ptr(x), _ptr(kernel),
                             _ptr(bias), _ptr(dy), _ptr(dx), _ptr(dkernel), _ptr(dbias), _ptr(ragged), _stream()),
          "sake_dense_bwd")


# ---- ragged batches (include/sake_b200.h: sake_ragged_*) ----------------------------------------------------
def ragged_bytes(B, N):
    return int(lib.sake_ragged_bytes(int(B), int(N)))


def ragged_prepare(B, N, n_real, blob):
    """n_real: int32 CUDA tensor [B]; blob: uint8 CUDA tensor of ragged_bytes(B, N)."""
    check(lib.sake_ragged_prepare(int(B), int(N), _ptr(n_real), _ptr(blob), blob.numel(), _stream()),
          "sake_ragged_prepare")


def ragged_gather(blob, B, N, width, padded, compact):
    check(lib.sake_ragged_gather(_ptr(blob), int(B), int(N), int(width), _ptr(padded), _ptr(compact), _stream()),
          "sake_ragged_gather")


def ragged_scatter(blob, B, N, width, compact, padded, alpha=1.0):
    check(lib.sake_ragged_scatter(_ptr(blob), int(B), int(N), int(width), float(alpha), _ptr(compact), _ptr(padded),
                                  _stream()), "sake_ragged_scatter")


class _DenseFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, kernel, bias, act):
        x = _f32c(x, "x")
        kernel = _f32c(kernel, "kernel")
        bias = _f32c(bias, "bias")
        y = torch.empty(*x.shape[:-1], kernel.shape[1], device=x.device, dtype=torch.float32)
        dense_fwd_raw(x, kernel, bias, y, act)
        ctx.save_for_backward(x, kernel, bias)
        ctx.act = act
        return y

    @staticmethod
    def backward(ctx, dy):
        x, kernel, bias = ctx.saved_tensors
        dy = _f32c(dy, "dy")
        dx = torch.empty_like(x) if ctx.needs_input_grad[0] else None
        dk = torch.zeros_like(kernel) if ctx.needs_input_grad[1] else None
        db = torch.zeros_like(bias) if (bias is not None and ctx.needs_input_grad[2]) else None
        dense_bwd_raw(x, kernel, bias, dy, dx, dk, db, ctx.act)
        return dx, dk, db, None


def dense(x, kernel, bias=None, act=0):
    """nn.Dense (+silu when act=1) through the CUDA library."""
    return _DenseFn.apply(x, kernel, bias, act)


class _LayerFn(torch.autograd.Function):
    """DenseSAKELayer.__call__ as one differentiable op (fwd / bwd = the two C-ABI entry points)."""

    @staticmethod
    def forward(ctx, cfg, h, x, v, mask, pair_u, pair_p, *leaves):
        paths = cfg["paths"]
        flat = {p: _f32c(t, p) for p, t in zip(paths, leaves)}
        h = _f32c(h, "h")
        x = _f32c(x, "x")
        v = _f32c(v, "v")
        mask = _f32c(mask, "mask")
        pair_u = _f32c(pair_u, "pair_u")
        pair_p = _f32c(pair_p, "pair_p")
        pair = None if pair_u is None else (pair_u, pair_p)
        lead = h.shape[:-2]
        N, H = h.shape[-2], h.shape[-1]
        B = 1
        for s in lead:
            B *= s
        dims = make_dims(B, N, H, cfg["A"], cfg["K"], cfg["update"], v is not None, mask is not None,
                         cfg["spatial"], cfg["engine"], cfg.get("cutoff"))
        ps, keep = params_struct(flat)
        dev = h.device
        saved = _buf(saved_bytes(dims), dev)
        scratch = _buf(scratch_bytes(dims, 0, 0), dev)
        h_out = torch.empty_like(h)
        x_out = torch.empty_like(x)
        v_out = torch.empty_like(x) if (cfg["update"] or v is not None) else None
        layer_fwd_raw(dims, ps, h, x, v, mask, h_out, x_out, v_out, saved, scratch, pair=pair)
        # leaves and the fwd->bwd buffer go through save_for_backward: autograd's version counters then catch an
        # in-place parameter update between forward and backward, and the buffer is freed with the graph
        ctx.cfg, ctx.dims = cfg, dims
        ctx.save_for_backward(h, x, v, mask, pair_u, pair_p, saved, *[flat[p] for p in paths])
        ctx.n_leaves = len(leaves)
        if v_out is None:
            v_out = x.new_zeros(())      # placeholder (reference returns None)
            ctx.mark_non_differentiable(v_out)
        return h_out, x_out, v_out

    @staticmethod
    def backward(ctx, dh_out, dx_out, dv_out):
        h, x, v, mask, pair_u, pair_p, saved_buf, *leaves = ctx.saved_tensors
        cfg, dims = ctx.cfg, ctx.dims
        flat = dict(zip(cfg["paths"], leaves))
        want_grads = any(ctx.needs_input_grad[7:])
        dev = h.device
        dh_out = _f32c(dh_out, "dh_out") if dh_out is not None else torch.zeros_like(h)
        dx_out = _f32c(dx_out, "dx_out") if dx_out is not None else None
        has_vout = cfg["update"] or v is not None
        dv_out = _f32c(dv_out, "dv_out") if (dv_out is not None and has_vout) else None
        dh = torch.empty_like(h)
        dx = torch.empty_like(x)
        dv = torch.empty_like(v) if v is not None else None
        gs, gflat = None, {}
        if want_grads:
            gflat = {p: torch.zeros_like(t) for p, t in flat.items()}
            gs, _ = params_struct(gflat, _lib.SakeLayerGrads)
        ps, keep = params_struct(flat)
        scratch = _buf(scratch_bytes(dims, 1, want_grads), dev)
        pair, g_u, g_p = None, None, None
        if pair_u is not None:
            g_u, g_p = torch.empty_like(pair_u), torch.empty_like(pair_p)
            pair = (pair_u, pair_p, g_u, g_p)
        layer_bwd_raw(dims, ps, h, x, v, mask, saved_buf, dh_out, dx_out, dv_out, dh, dx, dv, gs, scratch, pair=pair)
        grads = [gflat.get(p) if want_grads else None for p in cfg["paths"]]
        return (None, dh, dx, dv, None, g_u, g_p, *grads)


def sake_layer(flat_params, h, x, v=None, mask=None, *, n_heads=4, update=True, use_spatial_attention=True,
               engine="auto", cutoff=None, pair_u=None, pair_p=None):
    """flat_params: {flax path -> tensor} of one DenseSAKELayer.  Returns (h, x, v) like
    sake/layers.py:188-235 (v is None when the reference would return None)."""
    paths = tuple(p for p, _ in LAYER_LEAVES if p in flat_params)
    missing = [p for p in required_leaves(update, v is not None, use_spatial_attention) if p not in flat_params]
    if missing:
        # flax raises a missing-parameter error here (e.g. a tree initialised with v=None has no velocity_mlp,
        # sake/layers.py:226-229); the C ABI returns SAKE_EINVAL for the same condition
        raise _lib.SakeError("parameter tree lacks leaves this call needs: " + ", ".join(missing))
    K = flat_params["edge_model/kernel/means"].shape[0]
    cfg = {"paths": paths, "A": int(n_heads), "K": int(K), "update": bool(update),
           "spatial": bool(use_spatial_attention), "engine": engine, "cutoff": cutoff}
    h_out, x_out, v_out = _LayerFn.apply(cfg, h, x, v, mask, pair_u, pair_p, *[flat_params[p] for p in paths])
    if not (update or v is not None):
        v_out = None
    return h_out, x_out, v_out
