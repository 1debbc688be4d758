"""Augmented normalising flow on top of DenseSAKEModel (sake/flows.py:12-27,97-188).
The message-passing model is the CUDA path; the coupling glue (dummy atom, mean-centring,
scale MLP) is O(N) tensor plumbing in torch."""
import math

import torch

from . import ops
from .layers import _generator, dense_init, tree_to
from .models import DenseSAKEModel


class CenteredGaussian:
    @staticmethod
    def log_prob(value):
        # sake/flows.py:13-21
        n, d = value.shape[-2], value.shape[-1]
        r2 = (value ** 2).reshape(*value.shape[:-2], -1).sum(-1)
        return -0.5 * r2 - 0.5 * (n - 1) * d * math.log(2 * math.pi)

    @staticmethod
    def sample(key, shape, device="cuda"):
        # sake/flows.py:23-27
        x = torch.randn(tuple(shape), generator=_generator(key)).to(device)
        return x - x.mean(dim=-2, keepdim=True)


class AugmentedFlowLayer:
    """sake/flows.py:97-144"""

    def __init__(self, hidden_features=64, depth=3, activation=None, engine="auto"):
        self.hidden_features = hidden_features
        self.depth = depth
        self.sake_model = DenseSAKEModel(hidden_features=hidden_features, depth=depth, out_features=1,
                                         engine=engine)

    def init_params(self, gen, h_features, device):
        hx = torch.zeros(1, 2, h_features + 1, device=device)
        xx = torch.zeros(1, 2, 3, device=device)
        p = {"sake_model": self.sake_model.init(gen, hx, xx)["params"],
             "scale_mlp": tree_to({"layers_0": dense_init(gen, 1, self.hidden_features),
                                   "layers_2": dense_init(gen, self.hidden_features, 1, use_bias=False)}, device)}
        return p

    def mp(self, p, h, x):
        # sake/flows.py:118-129
        x0 = x
        h = torch.cat([h, (x ** 2).sum(-1, keepdim=True)], dim=-1)
        h = torch.cat([h, torch.zeros_like(h[..., -1:, :])], dim=-2)
        x = torch.cat([x, torch.zeros_like(x[..., -1:, :])], dim=-2)
        h, x, _ = self.sake_model(p["sake_model"], h, x)
        x = x[..., :-1, :]
        h = h[..., :-1, :]
        translation = x - x0
        translation = translation - translation.mean(dim=-2, keepdim=True)
        s = ops.dense(h, p["scale_mlp"]["layers_0"]["kernel"], p["scale_mlp"]["layers_0"]["bias"], act=1)
        s = torch.tanh(ops.dense(s, p["scale_mlp"]["layers_2"]["kernel"]))
        scale = s.mean(dim=-2, keepdim=True)
        return scale, translation

    def f_forward(self, p, h, x, v):
        scale, translation = self.mp(p, h, x)
        v = torch.exp(scale) * v + translation
        log_det = scale.sum((-1, -2)) * v.shape[-1] * v.shape[-2]
        return x, v, log_det

    def f_backward(self, p, h, x, v):
        scale, translation = self.mp(p, h, x)
        v = torch.exp(-scale) * (v - translation)
        log_det = scale.sum((-1, -2)) * v.shape[-1] * v.shape[-2]
        return x, v, log_det


class AugmentedFlowModel:
    """sake/flows.py:146-188"""

    def __init__(self, depth=3, mp_depth=3, hidden_features=64, activation=None, engine="auto"):
        self.depth = depth
        self.mp_depth = mp_depth
        self.hidden_features = hidden_features
        self.xv_layers = [AugmentedFlowLayer(hidden_features, mp_depth, engine=engine) for _ in range(depth)]
        self.vx_layers = [AugmentedFlowLayer(hidden_features, mp_depth, engine=engine) for _ in range(depth)]

    def init(self, key, h, x, v):
        gen = _generator(key)
        p = {}
        for i in range(self.depth):
            p["xv_%d" % i] = self.xv_layers[i].init_params(gen, h.shape[-1], h.device)
            p["vx_%d" % i] = self.vx_layers[i].init_params(gen, h.shape[-1], h.device)
        return {"params": p}

    def apply(self, variables, h, x, v, method=None):
        name = "f_forward" if method is None else (method if isinstance(method, str) else method.__name__)
        if name == "__call__":
            name = "f_forward"
        return getattr(self, name)(h, x, v, params=variables["params"])

    def f_forward(self, h, x, v, params=None):
        s = 0.0
        for i in reversed(range(self.depth)):
            x, v, ld = self.xv_layers[i].f_forward(params["xv_%d" % i], h, x, v)
            s = s + ld
            v, x, ld = self.vx_layers[i].f_forward(params["vx_%d" % i], h, v, x)
            s = s + ld
        return x, v, s

    def f_backward(self, h, x, v, params=None):
        s = 0.0
        for i in range(self.depth):
            v, x, ld = self.vx_layers[i].f_backward(params["vx_%d" % i], h, v, x)
            s = s + ld
            x, v, ld = self.xv_layers[i].f_backward(params["xv_%d" % i], h, x, v)
            s = s + ld
        return x, v, s

    def __call__(self, params, h, x, v):
        return self.f_forward(h, x, v, params=params)
