"""Preallocated, stream-ordered execution of a DenseSAKEModel through the C ABI: the energy+forces
closure (scripts/md17/run.py:46-58) and the energy-L1 training step (scripts/qm9/run.py:74-96,
134-138) without autograd or per-step allocation.  Every buffer lives in HBM for the lifetime of
the runner; a step is a fixed sequence of library calls on the current stream, so it can be
captured in a CUDA graph."""
import ctypes as C

import torch

from . import _lib, ops
from ._lib import lib, check
from .layers import flatten_tree


class ParamSet:
    """The parameters of one DenseSAKEModel as ONE flat fp32 vector (+ gradient / Adam vectors when training: one
    all-reduce bucket) with per-leaf views and the per-layer C structs (SakeLayerParams / SakeLayerGrads)."""

    def __init__(self, model, params, device, train=False):
        dev, f32 = torch.device(device), torch.float32
        flat = flatten_tree(params)
        self.paths = list(flat.keys())
        sizes = [flat[k].numel() for k in self.paths]
        # every tensor starts on a 16-byte boundary (vectorised / cp.async weight loads in the node kernels);
        # the pad floats stay zero in the parameters, gradients and Adam moments
        offs, off = [], 0
        for n in sizes:
            offs.append(off)
            off += (n + 3) // 4 * 4
        self.n_params = off
        self.n_params_real = sum(sizes)
        self.flat_params = torch.zeros(self.n_params, device=dev, dtype=f32)
        self.p = {}
        self.span = {k: (off, off + (n + 3) // 4 * 4) for k, n, off in zip(self.paths, sizes, offs)}
        for k, n, off in zip(self.paths, sizes, offs):
            self.p[k] = self.flat_params[off:off + n].view(flat[k].shape)
            self.p[k].copy_(flat[k])
        self.g = {}
        self.flat_grads = self.adam_m = self.adam_v = None
        if train:
            self.flat_grads = torch.zeros(self.n_params, device=dev, dtype=f32)
            self.adam_m = torch.zeros_like(self.flat_grads)
            self.adam_v = torch.zeros_like(self.flat_grads)
            for k, n, off in zip(self.paths, sizes, offs):
                self.g[k] = self.flat_grads[off:off + n].view(flat[k].shape)
        self.K = flat["d0/edge_model/kernel/means"].shape[0]
        self.ps, self.gs, self._keep = [], [], []
        for l in range(model.depth):
            sub = {k[len("d%d/" % l):]: t for k, t in self.p.items() if k.startswith("d%d/" % l)}
            ps, keep = ops.params_struct(sub)
            self.ps.append(ps)
            self._keep.append(keep)
            if train:
                gsub = {k[len("d%d/" % l):]: t for k, t in self.g.items() if k.startswith("d%d/" % l)}
                gs, keep = ops.params_struct(gsub, _lib.SakeLayerGrads)
                self.gs.append(gs)
                self._keep.append(keep)


    def bucket(self, prefix):
        """(begin, end) of the contiguous slice of the flat vectors that holds every leaf under `prefix`."""
        sp = [v for k, v in self.span.items() if k.startswith(prefix)]
        b, e = min(a for a, _ in sp), max(c for _, c in sp)
        assert sum(c - a for a, c in sp) == e - b, "leaves under one prefix are contiguous in the flat vector"
        return b, e


class ModelRunner:
    def __init__(self, model, params, B, N, in_features, *, masked=False, ragged=False, train=False, device="cuda",
                 lr=1e-3, weight_decay=1e-5, max_delta=1.0, mean=0.0, std=1.0, defer_dw=False, async_prep=True,
                 defer_reduce=False):
        """masked: QM9-style padded batch with the reference's float mask = outer(m, m) (every kernel computes
        all N^2 pairs of the padded width and multiplies by the mask, like the reference).
        ragged: the same padded batch, but only the n_real[b] real atoms of every molecule are stored and
        computed (compact layout, include/sake_b200.h `sake_ragged_*`); inputs / forces stay padded at the API."""
        if ragged and masked:
            raise _lib.SakeError("ragged batches carry no float mask: pass masked=False")
        self.ragged = bool(ragged)
        self.model, self.B, self.N, self.F = model, int(B), int(N), int(in_features)
        self.H, self.L, self.A = model.hidden_features, model.depth, model.n_heads
        self.out = model.out_features
        self.masked, self.train, self.dev = masked, train, torch.device(device)
        self.lr, self.wd, self.max_delta, self.mean, self.std = lr, weight_decay, max_delta, mean, std
        self.step_count = 0
        self.graphs, self.graph_replays, self.graph_launches = {}, 0, 0
        dev, f32 = self.dev, torch.float32
        # ---- one flat fp32 parameter vector (+ grads, Adam moments): one all-reduce bucket --------
        self.pset = ps_ = ParamSet(model, params, dev, train)
        self.paths, self.n_params, self.n_params_real = ps_.paths, ps_.n_params, ps_.n_params_real
        self.flat_params, self.p, self.g = ps_.flat_params, ps_.p, ps_.g
        self.flat_grads, self.adam_m, self.adam_v = ps_.flat_grads, ps_.adam_m, ps_.adam_v
        self.ps, self.gs = ps_.ps, ps_.gs
        # ---- per-layer dims ---------------------------------------------------------------------
        K = ps_.K
        self.dims = []
        has_v = False
        self.has_v = []
        for l in range(self.L):
            upd = model.update_list[l]
            d = ops.make_dims(B, N, self.H, self.A, K, upd, has_v, masked, model.use_spatial_attention,
                              model.engine, model.layers[l]._cutoff)
            self.dims.append(d)
            self.has_v.append(has_v)
            has_v = has_v or upd
        self.engine = ops.resolve_engine(self.dims[0])
        # ---- activations -----------------------------------------------------------------------
        R = B * N
        self.h_in = torch.zeros(B, N, self.F, device=dev, dtype=f32)
        self.x_in = torch.zeros(B, N, 3, device=dev, dtype=f32)
        self.rg = None
        if self.ragged:
            # padded copies of the inputs as the caller hands them over; h_in / x_in above hold the compact rows
            self.h_pad = torch.zeros(B, N, self.F, device=dev, dtype=f32)
            self.x_pad = torch.zeros(B, N, 3, device=dev, dtype=f32)
            self.n_real = torch.full((B,), N, device=dev, dtype=torch.int32)
            self.rg = ops._buf(ops.ragged_bytes(B, N), dev)
            self.dx_pad = torch.zeros(B, N, 3, device=dev, dtype=f32)
        self.mask = torch.ones(B, N, N, device=dev, dtype=f32) if masked else None
        self.atom_mask = torch.ones(B, N, device=dev, dtype=f32) if masked else None
        self.target = torch.zeros(B, device=dev, dtype=f32)
        self.hs = [torch.empty(B, N, self.H, device=dev, dtype=f32) for _ in range(self.L + 1)]
        self.xs = [self.x_in] + [torch.empty(B, N, 3, device=dev, dtype=f32) for _ in range(self.L)]
        self.vs = [None] + [torch.empty(B, N, 3, device=dev, dtype=f32) for _ in range(self.L)]
        self.saved = [ops._buf(ops.saved_bytes(d), dev) for d in self.dims]
        # inference: the weights do not change between steps, so their tcgen05 operand images are built ONCE
        # (sake_layer_prepare) and every forward call skips the five small preparation launches per layer;
        # call refresh_weights() after changing self.p in place.  Training rebuilds them in every forward call.
        self.weights_prepared = False
        self.dims_unprepared = [_lib.SakeDims(d.B, d.N, d.H, d.A, d.K, d.flags, d.engine, d.reserved, d.cutoff_lower,
                                              d.cutoff_upper) for d in self.dims]
        if not train:
            for d in self.dims:
                d.flags |= _lib.SAKE_WEIGHTS_PREPARED
            self.refresh_weights()
        # training: the parameters change every step, so the images are rebuilt every step — on a side stream, layer
        # by layer, under the forward kernels of the earlier layers (layer l's forward waits for its own event)
        self.async_prep = bool(async_prep and train and self.engine != "fp32")
        if self.async_prep:
            self.prep_stream = torch.cuda.Stream(device=dev)
            self.prep_events = [torch.cuda.Event() for _ in range(self.L)]
            self.dims_prepared = [_lib.SakeDims(d.B, d.N, d.H, d.A, d.K, d.flags | _lib.SAKE_WEIGHTS_PREPARED, d.engine,
                                                d.reserved, d.cutoff_lower, d.cutoff_upper) for d in self.dims]
        nscr = max(max(ops.scratch_bytes(d, 0, 0), ops.scratch_bytes(d, 1, int(train))) for d in self.dims)
        self.scratch = ops._buf(nscr, dev)
        # defer_dw (opt-in): the weight-gradient contractions of layer l run on the library's side stream under the
        # backward of layer l-1 (SAKE_DEFER_DW), which needs a second scratch buffer to alternate with.  Measured on
        # B200 (profiles/r02_*): no gain — every big kernel of the backward wants whole SMs (200+ KB of shared memory,
        # 512 TMEM columns), so the two streams take turns on each SM instead of overlapping (cfg2 3.58 -> 3.83 ms).
        self.defer_dw = bool(defer_dw and train and self.engine != "fp32" and self.L > 1)
        self.scratches = [self.scratch, ops._buf(nscr, dev)] if self.defer_dw else [self.scratch, self.scratch]
        # defer_reduce (opt-in): only the partial-sum reduction that finishes a layer's weight gradients goes to the
        # side stream (SAKE_DEFER_REDUCE), under this layer's and the next layer's per-node kernels, which occupy a
        # quarter of the SMs.  Measured on B200 (same box, cfg2): 2.547 vs 2.538 ms without — the reduction's CTAs
        # delay the node kernels by as much as they save (k_tc_node_pre_bwd 17 -> 31 us).  Off by default.
        self.defer_reduce = bool(defer_reduce and train and self.engine != "fp32" and not self.defer_dw)
        self.dims_bwd = []
        for l, d in enumerate(self.dims):
            db = _lib.SakeDims(d.B, d.N, d.H, d.A, d.K, d.flags | (_lib.SAKE_DEFER_DW if self.defer_dw else 0) |
                               (_lib.SAKE_DEFER_REDUCE if self.defer_reduce else 0),
                               d.engine, l & 1, d.cutoff_lower, d.cutoff_upper)
            self.dims_bwd.append(db)
        self.y0 = torch.empty(B, N, self.H, device=dev, dtype=f32)
        self.y = torch.empty(B, N, self.out, device=dev, dtype=f32)
        self.energy = torch.empty(B, device=dev, dtype=f32)
        self.loss = torch.zeros(1, device=dev, dtype=f32)
        self.dy = torch.empty_like(self.y)
        self.dy0 = torch.empty_like(self.y0)
        self.dh = [torch.empty(B, N, self.H, device=dev, dtype=f32) for _ in range(2)]
        self.dx = [torch.empty(B, N, 3, device=dev, dtype=f32) for _ in range(2)]
        self.dv = [torch.empty(B, N, 3, device=dev, dtype=f32) for _ in range(2)]
        self.forces = torch.empty(B, N, 3, device=dev, dtype=f32)
        self.hbm_bytes = sum(t.numel() * t.element_size() for t in
                             [self.flat_params, *set(self.scratches), *self.saved, *self.hs, *self.xs[1:], *self.vs[1:]])

    def refresh_weights(self, pset=None):
        """(Re)build the weight operand images of every layer from the current parameters (inference runners)."""
        pset = self.pset if pset is None else pset
        for l in range(self.L):
            check(lib.sake_layer_prepare(C.byref(self.dims[l]), C.byref(pset.ps[l]), ops._ptr(self.saved[l]),
                                         self.saved[l].numel(), ops._stream()), "sake_layer_prepare")
        self.weights_prepared = True

    # -- inputs ------------------------------------------------------------------------------------
    def load_inputs(self, h, x, mask=None, atom_mask=None, target=None, n_real=None):
        """Copy one batch into the resident input buffers (host pinned or device tensors).
        ragged runners take the padded h [B,N,F], x [B,N,3] and n_real [B] (int32) instead of the masks."""
        if self.ragged:
            if n_real is None:
                raise _lib.SakeError("a ragged runner needs n_real")
            self.h_pad.copy_(h, non_blocking=True)
            self.x_pad.copy_(x, non_blocking=True)
            self.n_real.copy_(n_real, non_blocking=True)
        else:
            self.h_in.copy_(h, non_blocking=True)
            self.x_in.copy_(x, non_blocking=True)
        if self.masked:
            self.mask.copy_(mask, non_blocking=True)
            self.atom_mask.copy_(atom_mask, non_blocking=True)
        if target is not None:
            self.target.copy_(target, non_blocking=True)

    def input_bytes(self):
        n = self.h_in.numel() + self.x_in.numel() + (self.target.numel() if self.train else 0)
        if self.masked:
            n += self.mask.numel() + self.atom_mask.numel()
        if self.ragged:
            n += self.n_real.numel()
        return 4 * n

    def _ragged_inputs(self):
        """Tables from n_real (device side) and the compact copies of the padded inputs: first launches of a step,
        so a CUDA-graph replay picks up whatever batch load_inputs put into the resident buffers."""
        ops.ragged_prepare(self.B, self.N, self.n_real, self.rg)
        ops.ragged_gather(self.rg, self.B, self.N, self.F, self.h_pad, self.h_in)
        ops.ragged_gather(self.rg, self.B, self.N, 3, self.x_pad, self.x_in)

    # -- forward (sake/models.py:56-61) -------------------------------------------------------------
    def forward(self, pset=None):
        """pset: another ParamSet of the same architecture (the flow runs 2*depth models through one set of
        activation buffers); default: the runner's own parameters."""
        foreign = pset is not None and pset is not self.pset
        pset = self.pset if pset is None else pset
        p = pset.p
        rg = self.rg
        # another model's parameters through this runner's buffers (flows): the images in `saved` are not theirs
        dims = self.dims_unprepared if foreign else self.dims
        async_prep = self.async_prep and not foreign
        if async_prep:
            main = torch.cuda.current_stream()
            self.prep_stream.wait_stream(main)            # after the optimiser step that produced these parameters
            with torch.cuda.stream(self.prep_stream):
                for l in range(self.L):
                    check(lib.sake_layer_prepare(C.byref(self.dims[l]), C.byref(pset.ps[l]), ops._ptr(self.saved[l]),
                                                 self.saved[l].numel(), ops._stream()), "sake_layer_prepare")
                    self.prep_events[l].record()
            dims = self.dims_prepared
        if self.ragged:
            self._ragged_inputs()
        ops.dense_fwd_raw(self.h_in, p["embedding_in/kernel"], p.get("embedding_in/bias"), self.hs[0], 0, rg)
        for l in range(self.L):
            v_in = self.vs[l] if self.has_v[l] else None
            upd = self.model.update_list[l]
            v_out = self.vs[l + 1] if (upd or v_in is not None) else None
            if async_prep:
                torch.cuda.current_stream().wait_event(self.prep_events[l])
            ops.layer_fwd_raw(dims[l], pset.ps[l], self.hs[l], self.xs[l], v_in, self.mask,
                              self.hs[l + 1], self.xs[l + 1], v_out, self.saved[l], self.scratch, rg)
        ops.dense_fwd_raw(self.hs[self.L], p["embedding_out/layers_0/kernel"],
                          p.get("embedding_out/layers_0/bias"), self.y0, 1, rg)
        ops.dense_fwd_raw(self.y0, p["embedding_out/layers_2/kernel"], p.get("embedding_out/layers_2/bias"),
                          self.y, 0, rg)

    def _head(self, mode):
        check(lib.sake_energy_head(self.B, self.N, self.out, mode, ops._ptr(self.y), ops._ptr(self.atom_mask),
                                   ops._ptr(self.target) if mode == 1 else None, self.mean, self.std,
                                   ops._ptr(self.energy), ops._ptr(self.loss) if mode == 1 else None,
                                   ops._ptr(self.dy), ops._ptr(self.rg), ops._stream()), "sake_energy_head")

    # -- backward ------------------------------------------------------------------------------------
    # The backward pass in segments (readout, one per layer from the last to the first, embedding): a training step
    # on several GPUs starts the all-reduce of a layer's gradient bucket right after that layer's segment.
    # The cotangent buffers ping-pong: layer l reads slot (L-1-l) & 1 and writes the other one.
    def _bwd_readout(self, with_grads):
        p, g = self.p, self.g
        gk = (lambda k: g[k]) if with_grads else (lambda k: None)
        ops.dense_bwd_raw(self.y0, p["embedding_out/layers_2/kernel"], p.get("embedding_out/layers_2/bias"),
                          self.dy, self.dy0, gk("embedding_out/layers_2/kernel"),
                          gk("embedding_out/layers_2/bias") if "embedding_out/layers_2/bias" in p else None, 0, self.rg)
        ops.dense_bwd_raw(self.hs[self.L], p["embedding_out/layers_0/kernel"],
                          p.get("embedding_out/layers_0/bias"), self.dy0, self.dh[0],
                          gk("embedding_out/layers_0/kernel"), gk("embedding_out/layers_0/bias"), 1, self.rg)

    def _bwd_layer(self, l, with_grads):
        cur = (self.L - 1 - l) & 1
        nxt = 1 - cur
        last = l == self.L - 1
        v_in = self.vs[l] if self.has_v[l] else None
        dx_out = None if last else self.dx[cur]
        dv_out = None if (last or not self.has_v[l + 1]) else self.dv[cur]
        ops.layer_bwd_raw(self.dims_bwd[l] if with_grads else self.dims[l], self.ps[l], self.hs[l], self.xs[l], v_in,
                          self.mask, self.saved[l], self.dh[cur], dx_out, dv_out, self.dh[nxt], self.dx[nxt],
                          self.dv[nxt] if v_in is not None else None,
                          self.gs[l] if with_grads else None, self.scratches[l & 1] if with_grads else self.scratch,
                          self.rg)

    def _bwd_layer_joined(self, l):
        """One layer's backward as a self-contained segment (its gradient bucket is complete when it returns)."""
        self._bwd_layer(l, True)
        if self.defer_dw or self.defer_reduce:
            check(lib.sake_dw_sync(ops._stream()), "sake_dw_sync")

    def _bwd_embed(self, with_grads):
        cur = self.L & 1                       # slot the first layer wrote
        if with_grads:
            p, g = self.p, self.g
            ops.dense_bwd_raw(self.h_in, p["embedding_in/kernel"], p.get("embedding_in/bias"), self.dh[cur], None,
                              g["embedding_in/kernel"], g.get("embedding_in/bias"), 0, self.rg)
            if self.defer_dw or self.defer_reduce:       # join the side stream: gradients complete from here on
                check(lib.sake_dw_sync(ops._stream()), "sake_dw_sync")
        self._dx_final = self.dx[cur]

    def backward(self, with_grads):
        self._bwd_readout(with_grads)
        for l in reversed(range(self.L)):
            self._bwd_layer(l, with_grads)
        self._bwd_embed(with_grads)

    # -- the two driver closures ---------------------------------------------------------------------
    def _ef_body(self):
        self.forward()
        self._head(0)
        self.backward(False)
        if self.ragged:                      # compact -dE/dx -> padded forces (padding atoms: exactly 0)
            self.forces.zero_()
            ops.ragged_scatter(self.rg, self.B, self.N, 3, self._dx_final, self.forces, alpha=-1.0)
        else:
            torch.neg(self._dx_final, out=self.forces)

    def _train_body(self):
        self.flat_grads.zero_()
        self.loss.zero_()
        self.forward()
        self._head(1)
        self.backward(True)

    # the same step cut at the points where a gradient bucket becomes final (multi-GPU training)
    def _train_segments(self):
        def head():
            self.flat_grads.zero_()
            self.loss.zero_()
            self.forward()
            self._head(1)
            self._bwd_readout(True)
        segs = [("embedding_out/", head)]
        for l in reversed(range(self.L)):
            segs.append(("d%d/" % l, (lambda l=l: self._bwd_layer_joined(l))))
        segs.append(("embedding_in/", lambda: self._bwd_embed(True)))
        return segs

    def capture(self, train=None):
        """Record the enqueue-only part of a step (everything the library launches between the input copy and
        the optimiser / result read) into a CUDA graph; later steps of that kind replay it with one launch.
        All buffers are preallocated and the library never synchronises, so the capture is exact.  The gradient
        all-reduce and the Adam kernel (its step counter is a host argument) stay outside the graph.
        `train`: which step to record (default: the runner's own mode).  Returns the number of library launches
        one replay stands for."""
        train = self.train if train is None else bool(train)
        body = self._train_body if train else self._ef_body
        for _ in range(2):                       # first-use work (function attributes) must not happen inside a capture
            body()
        torch.cuda.synchronize(self.dev)
        n0 = lib.sake_launch_count()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            body()
        self.graph_launches = int(lib.sake_launch_count() - n0)
        self.graphs[train] = g
        if train:
            # the segmented form of the same step (one graph per gradient bucket), used when an all-reduce runs
            # between the segments; all graphs share one memory pool and the runner's static buffers
            self.seg_graphs = []
            for prefix, fn in self._train_segments():
                sg = torch.cuda.CUDAGraph()
                with torch.cuda.graph(sg, pool=g.pool()):
                    fn()
                self.seg_graphs.append((prefix, sg))
        return self.graph_launches

    def _run_body(self, train):
        g = self.graphs.get(train)
        if g is not None:
            g.replay()
            self.graph_replays += 1
        elif train:
            self._train_body()
        else:
            self._ef_body()

    def energy_forces_step(self):
        """E[b] and F = -dE/dx for the resident batch (scripts/md17/run.py:46-58)."""
        self._run_body(False)
        return self.energy, self.forces

    def train_step(self, allreduce=None, bucketed=False):
        """One energy-L1 training step (scripts/qm9/run.py:79-89): fwd, bwd with parameter grads,
        optional gradient all-reduce (lax.pmean, scripts/ani/run_gpu.py:130), AdamW-style chain."""
        scale = 1.0
        if bucketed and allreduce is not None and getattr(allreduce, "world", 1) > 1 and hasattr(allreduce, "start"):
            # per-layer buckets (lax.pmean over the same leaves, scripts/ani/run_gpu.py:130): the all-reduce of layer
            # l's gradients runs on the NCCL stream under the backward of layer l-1.  Opt-in: measured on 2 x B200 it
            # is SLOWER than one all-reduce after the backward (3.61 vs 3.4 ms per cfg2 step) — the collective's CTAs
            # take SMs away from persistent kernels that are sized one CTA per SM, which then need a second wave.
            segs = getattr(self, "seg_graphs", None)
            for prefix, fn in (segs if segs else self._train_segments()):
                if segs:
                    fn.replay()
                else:
                    fn()
                b, e = self.pset.bucket(prefix)
                allreduce.start(self.flat_grads[b:e])
            if segs:
                self.graph_replays += 1
            scale = allreduce.finish()
        else:
            self._run_body(True)
            if allreduce is not None:
                scale = allreduce(self.flat_grads)
        self.step_count += 1
        check(lib.sake_adam_step(self.n_params, ops._ptr(self.flat_params), ops._ptr(self.flat_grads),
                                 ops._ptr(self.adam_m), ops._ptr(self.adam_v), self.step_count, self.lr, 0.9,
                                 0.999, 1e-8, self.wd, self.max_delta, scale, ops._stream()), "sake_adam_step")
        return self.loss

    def params_tree(self):
        from .layers import unflatten_tree
        return unflatten_tree({k: v for k, v in self.p.items()})


def profile_begin(capacity=4096):
    check(lib.sake_profile_begin(int(capacity)), "sake_profile_begin")


def profile_collect(capacity=4096):
    ms = (C.c_float * capacity)()
    kind = (C.c_int32 * capacity)()
    pairs = (C.c_int64 * capacity)()
    n = lib.sake_profile_collect(ms, kind, pairs, capacity)
    return [(float(ms[i]), int(kind[i]), int(pairs[i])) for i in range(n)]
