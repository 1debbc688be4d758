"""Preallocated, stream-ordered execution of AugmentedFlowModel (sake/flows.py:146-188): the sampling pass
(f_forward) and the log-likelihood pass (f_backward + the two prior log-probs, scripts/lj13_aug/run.py:39-43)
without autograd or per-step allocation.  Every coupling layer is one DenseSAKEModel forward on N+1 atoms
(ModelRunner.forward with that layer's ParamSet; all 2*depth models share one set of activation buffers) between
the two glue kernels of csrc/flow.cu, so a whole pass is a fixed launch sequence that can be replayed as a CUDA
graph.  (Training the flow — the gradient of the likelihood — goes through sake_b200.flows, the autograd path.)"""
import torch

from . import ops
from ._lib import lib, check
from .flows import AugmentedFlowModel
from .layers import _generator
from .runner import ModelRunner, ParamSet


class FlowRunner:
    def __init__(self, depth=4, mp_depth=4, B=4096, N=13, D=3, h_features=2, seed=0, engine="auto", device="cuda",
                 params=None, hidden_features=64):
        self.depth, self.B, self.N, self.D, self.Fh = depth, int(B), int(N), int(D), int(h_features)
        self.dev = dev = torch.device(device)
        self.flow = AugmentedFlowModel(depth=depth, mp_depth=mp_depth, hidden_features=hidden_features, engine=engine)
        if params is None:
            hx = torch.zeros(1, 2, h_features)
            params = self.flow.init(_generator(seed), hx, torch.zeros(1, 2, D), torch.zeros(1, 2, D))["params"]
        model = self.flow.xv_layers[0].sake_model
        self.Hs = hidden_features
        # one runner = the activation buffers of a DenseSAKEModel on [B, N+1] atoms; a ParamSet per coupling layer
        first = params["xv_0"]["sake_model"]
        self.run = ModelRunner(model, first, B, N + 1, h_features + 1, device=dev)
        self.psets, self.scale = {}, {}
        for i in range(depth):
            for nm in ("xv_%d" % i, "vx_%d" % i):
                # every coupling model gets its own ParamSet (also xv_0: the runner's built-in set would count as
                # "own" and skip the weight-image rebuild that sharing one set of `saved` buffers makes necessary)
                self.psets[nm] = ParamSet(model, params[nm]["sake_model"], dev)
                sm = params[nm]["scale_mlp"]
                self.scale[nm] = tuple(t.detach().to(dev).float().contiguous() for t in
                                       (sm["layers_0"]["kernel"], sm["layers_0"]["bias"], sm["layers_2"]["kernel"]))
        self.engine = self.run.engine
        f32 = torch.float32
        self.h = torch.zeros(B, N, h_features, device=dev, dtype=f32)       # scripts/lj13_aug/run.py:34: zeros
        self.x = torch.zeros(B, N, 3, device=dev, dtype=f32)
        self.v = torch.zeros(B, N, 3, device=dev, dtype=f32)
        self.x0, self.v0 = torch.zeros_like(self.x), torch.zeros_like(self.v)
        self.logdet = torch.zeros(B, device=dev, dtype=f32)
        self.out = torch.zeros(B, device=dev, dtype=f32)
        self.graph, self.graph_kind, self.graph_replays, self.graph_launches = None, None, 0, 0

    def load_inputs(self, x, v, h=None):
        """x, v: [B, N, D] (host pinned or device)."""
        if x.shape[-1] == 3:
            self.x.copy_(x, non_blocking=True)
            self.v.copy_(v, non_blocking=True)
        else:
            self.x[..., :x.shape[-1]].copy_(x, non_blocking=True)
            self.v[..., :v.shape[-1]].copy_(v, non_blocking=True)
        if h is not None:
            self.h.copy_(h, non_blocking=True)
        self.x0.copy_(self.x)
        self.v0.copy_(self.v)

    def restore_inputs(self):
        """Both passes transform (x, v) in place; this puts the loaded batch back (device-to-device)."""
        self.x.copy_(self.x0)
        self.v.copy_(self.v0)

    # one AugmentedFlowLayer: mp(h, pos) then the affine update of `other` (flows.py:118-142)
    def _coupling(self, name, pos, other, direction):
        r, B, N = self.run, self.B, self.N
        check(lib.sake_flow_pre(B, N, self.Fh, ops._ptr(self.h), ops._ptr(pos), ops._ptr(r.h_in), ops._ptr(r.x_in),
                                ops._stream()), "sake_flow_pre")
        r.forward(self.psets[name])
        w0, b0, w2 = self.scale[name]
        check(lib.sake_flow_post(B, N, self.D, self.Hs, direction, ops._ptr(r.xs[r.L]), ops._ptr(pos), ops._ptr(r.y),
                                 ops._ptr(w0), ops._ptr(b0), ops._ptr(w2), ops._ptr(other), ops._ptr(self.logdet),
                                 ops._stream()), "sake_flow_post")

    def _sample_body(self):
        # AugmentedFlowModel.f_forward, flows.py:168-176
        self.logdet.zero_()
        for i in reversed(range(self.depth)):
            self._coupling("xv_%d" % i, self.x, self.v, +1)
            self._coupling("vx_%d" % i, self.v, self.x, +1)

    def _loglik_body(self):
        # AugmentedFlowModel.f_backward (flows.py:178-186) + the loss terms of scripts/lj13_aug/run.py:39-43
        self.logdet.zero_()
        for i in range(self.depth):
            self._coupling("vx_%d" % i, self.v, self.x, -1)
            self._coupling("xv_%d" % i, self.x, self.v, -1)
        check(lib.sake_flow_logprob(self.B, self.N, self.D, ops._ptr(self.x), ops._ptr(self.v), ops._ptr(self.logdet),
                                    ops._ptr(self.out), ops._stream()), "sake_flow_logprob")

    def capture(self, kind="loglik"):
        body = self._loglik_body if kind == "loglik" else self._sample_body
        x0, v0 = self.x.clone(), self.v.clone()
        for _ in range(2):
            body()
        torch.cuda.synchronize(self.dev)
        n0 = lib.sake_launch_count()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            body()
        self.graph_launches = int(lib.sake_launch_count() - n0)
        self.graph, self.graph_kind = g, kind
        self.x.copy_(x0)
        self.v.copy_(v0)
        return self.graph_launches

    def _run(self, kind):
        if self.graph is not None and self.graph_kind == kind:
            self.graph.replay()
            self.graph_replays += 1
        elif kind == "loglik":
            self._loglik_body()
        else:
            self._sample_body()

    def log_likelihood_step(self):
        """In place: (x, v) -> latent; returns per-molecule -log p(x) - log p(v) + sum_log_det [B]."""
        self._run("loglik")
        return self.out

    def sample_step(self):
        """In place: latent (x, v) -> samples; returns the per-molecule sum_log_det [B]."""
        self._run("sample")
        return self.logdet
