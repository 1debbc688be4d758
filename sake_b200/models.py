"""DenseSAKEModel with the reference's surface (sake/models.py:11-61)."""
import torch

from . import ops
from .layers import DenseSAKELayer, _generator, dense_init, flatten_tree, init_layer_params, tree_to


class DenseSAKEModel:
    """Drop-in for sake.models.DenseSAKEModel: embedding_in -> depth x DenseSAKELayer -> embedding_out."""

    def __init__(self, hidden_features, out_features, depth=4, activation=None, update=True,
                 use_semantic_attention=True, use_euclidean_attention=True, use_spatial_attention=True,
                 n_heads=4, cutoff=None, engine="auto"):
        self.hidden_features = hidden_features
        self.out_features = out_features
        self.depth = depth
        self.activation = activation
        self.update = update
        self.use_semantic_attention = use_semantic_attention
        self.use_euclidean_attention = use_euclidean_attention
        self.use_spatial_attention = use_spatial_attention
        self.n_heads = n_heads
        self.cutoff = cutoff
        self.engine = engine
        upd = [update] * depth if isinstance(update, bool) else list(update)   # models.py:33-36
        self.update_list = upd
        self.layers = [
            DenseSAKELayer(hidden_features, hidden_features, activation=activation, n_heads=n_heads,
                           update=upd[i], use_semantic_attention=use_semantic_attention,
                           use_euclidean_attention=use_euclidean_attention,
                           use_spatial_attention=use_spatial_attention, cutoff=cutoff, engine=engine)
            for i in range(depth)
        ]

    def init(self, key, h, x, v=None, mask=None, he=None):
        gen = _generator(key)
        H = self.hidden_features
        p = {"embedding_in": dense_init(gen, h.shape[-1], H),
             "embedding_out": {"layers_0": dense_init(gen, H, H), "layers_2": dense_init(gen, H, self.out_features)}}
        has_v = v is not None
        for i in range(self.depth):
            p["d%d" % i] = init_layer_params(
                gen, H, H, H, self.n_heads, self.update_list[i], has_v,
                log_gamma=self.use_semantic_attention and self.use_euclidean_attention,
                edge_features=0 if he is None else he.shape[-1])
            has_v = has_v or self.update_list[i]
        return {"params": tree_to(p, h.device)}

    def apply(self, variables, h, x, v=None, mask=None, he=None, method=None):
        if method is not None and (method if isinstance(method, str) else method.__name__) != "__call__":
            raise ops._lib.SakeError("DenseSAKEModel has no sub-methods (sake/models.py:11-61 defines __call__ only)")
        return self(variables["params"], h, x, v, mask, he)

    def __call__(self, params, h, x, v=None, mask=None, he=None):
        # sake/models.py:56-61
        e = params["embedding_in"]
        h = ops.dense(h, e["kernel"], e.get("bias"))
        for i, layer in enumerate(self.layers):
            h, x, v = layer(params["d%d" % i], h, x, v, mask, he)
        o = params["embedding_out"]
        h = ops.dense(h, o["layers_0"]["kernel"], o["layers_0"].get("bias"), act=1)
        h = ops.dense(h, o["layers_2"]["kernel"], o["layers_2"].get("bias"))
        return h, x, v

    # -- energy / force closures of the drivers (scripts/md17/run.py:46-58) -------------------
    def energy(self, params, h, x, mask=None, atom_mask=None):
        y, _, _ = self(params, h, x, mask=mask)
        if atom_mask is not None:
            y = y * atom_mask.unsqueeze(-1)
        return y.sum(dim=(-1, -2))

    def energy_and_forces(self, params, h, x, mask=None, atom_mask=None):
        x = x.detach().requires_grad_(True)
        e = self.energy(params, h, x, mask=mask, atom_mask=atom_mask)
        (g,) = torch.autograd.grad(e.sum(), x)
        return e.detach(), -g
