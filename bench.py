#!/usr/bin/env python
"""Headline benchmark: SAKE DenseSAKEModel molecules/s (and atom-pairs/s) on B200.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--workload cfg1|cfg2|cfg3|cfg4|cfg5]
                  [--engine auto|fp32|tf32x3|bf16|f16x2] [--impl ours|reference]
                  [--padding ragged|masked] [--no-graphs] [--no-strong] [--no-cpu-baseline]

One "step" = one pass of the hot path over one batch of synthetic molecules:
  cfg1  MD17-aspirin-shaped, N=21, B=32, energy+forces            (the reference's CPU-runnable case)
  cfg2  QM9-shaped, padded to N=29, B=256, training step           (DEFAULT: the single-GPU config)
  cfg3  ANI-1x-shaped, padded to N=63, B=1024, energy+forces
  cfg4  OC20-slab-shaped, N=200, B=64, training step (+ grad all-reduce when N_gpus > 1)
  cfg5  LJ13 augmented flow (sake/flows.py), B=4096: log-likelihood pass (f_backward + log-probs); the sampling
        pass (f_forward) is reported beside it
Weak scaling (the headline line): every rank processes its own batch of B molecules; value = N_gpus * B / step time.
`strong` (same JSON line): the two configs BASELINE.json names for several GPUs, at FIXED total size — cfg3 with
1024 / N molecules per rank (no collective) and cfg4 with 64 / N per rank + NCCL gradient all-reduce.
Prints ONE JSON line (see the driver contract in the task statement / DESIGN.md section 6).
"""
import argparse
import importlib.util
import json
import math
import os
import subprocess
import sys
import threading
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: (B, N, species, padded, n_min, mode, description)
    "cfg1": (32, 21, 8, False, 0, "forces", "MD17-aspirin-shaped (21 atoms) energy+forces, batch 32"),
    "cfg2": (256, 29, 10, True, 9, "train", "QM9-shaped (padded to 29 atoms) training step, batch 256"),
    "cfg3": (1024, 63, 4, True, 20, "forces", "ANI-1x-shaped (padded to 63 atoms) energy+forces, batch 1024"),
    "cfg4": (64, 200, 84, False, 0, "train", "OC20-slab-shaped (200 atoms, dense all-pairs) training step, batch 64"),
    "cfg5": (4096, 13, 2, False, 0, "flow", "LJ13 augmented flow (13 atoms + dummy, depth 4 x mp_depth 4) log-likelihood pass, batch 4096"),
}
H, A, K, C = 64, 4, 50, 256
FLOP_PAIR = 2 * ((K + 1) * H + H * H + H * A + C * C + C)     # 146 816 (factored form)
FLOP_NODE = 2 * 67904                                         # 135 808
FLOP_MIX = 2 * C * C                                          # the x_mixing contraction alone, per pair
FLOP_EDGE = 2 * ((K + 1) * H + H * H + H * A)                 # edge MLP + logits, per pair
L2_NOTE = "GPU arm: L2 flushed between timed steps (256 MiB memset outside the per-step event pairs)"


def synth(seed, B, N, S, padded, n_min):
    """Synthetic molecules of the named shape: normal coords at molecular density,
    uniform species one-hots, QM9-style padding (scripts/qm9/run.py:23-24,35)."""
    rng = np.random.default_rng(seed)
    x = (rng.standard_normal((B, N, 3)) * 0.62 * N ** (1.0 / 3.0)).astype(np.float32)
    z = rng.integers(0, S, (B, N))
    h = np.eye(S, dtype=np.float32)[z]
    y = rng.standard_normal(B).astype(np.float32)
    if not padded:
        return h, x, None, None, y, np.full(B, N, dtype=np.int32)
    n_real = rng.integers(n_min, N + 1, B).astype(np.int32)
    am = (np.arange(N)[None, :] < n_real[:, None]).astype(np.float32)
    return h * am[..., None], x * am[..., None], am[:, :, None] * am[:, None, :], am, y, n_real


def _init_module():
    """sake_b200/init_params.py loaded BY PATH: torch-only flax-default initialisation, shared by both arms
    without importing the sake_b200 package (whose __init__ loads the CUDA library)."""
    spec = importlib.util.spec_from_file_location("_sake_init_params", os.path.join(ROOT, "sake_b200", "init_params.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def init_params_cpu(depth, S, seed):
    """flax-default initialisation, seeded; identical on every rank and in both arms."""
    ip = _init_module()
    return ip.init_model_params(ip._generator(seed), S, H, 1, depth, n_heads=A)


def make_config(wl, depth, world, padding, strong):
    B, N, S, padded, n_min, mode, desc = WORKLOADS[wl]
    model = (f"AugmentedFlowModel(depth=4, mp_depth=4, hidden=64)" if mode == "flow"
             else f"DenseSAKEModel(hidden=64, heads=4, depth={depth})")
    return {"workload": f"{wl}: {desc}; {model}", "molecules_per_gpu": B, "atoms_padded": N, "mode": mode,
            "parallelism": f"dp{world}" if world > 1 else "single", "l2": L2_NOTE}


# ------------------------------------------------------------------------------------------------
# CPU arm: the oracle port of the reference (JAX itself is not installable in this image).  Imports
# only torch, numpy, oracle/ and (by path) the torch-only init module — never the sake_b200 package.
# ------------------------------------------------------------------------------------------------
def cpu_step_fn(params_cpu, mode, batch):
    from oracle import sake_oracle as O
    h, x, mask, am, y = batch
    p = O.tree_map(lambda t: t.clone().float(), params_cpu)
    leaves = list(O.tree_flatten(p).values())
    if mode == "train":
        for t in leaves:
            t.requires_grad_(True)
        # the optimiser chain of scripts/qm9/run.py:134-138, exactly what the GPU arm's fused kernel applies:
        # additive_weight_decay(1e-5) -> clip(1.0, element-wise) -> adam(1e-3)
        m = [torch.zeros_like(t) for t in leaves]
        v = [torch.zeros_like(t) for t in leaves]
        state = {"t": 0}

        def step():
            e = O.energy(p, h, x, mask=mask, atom_mask=am)
            loss = (e - y).abs().mean()
            grads = torch.autograd.grad(loss, leaves, allow_unused=True)
            state["t"] += 1
            c1, c2 = 1.0 / (1.0 - 0.9 ** state["t"]), 1.0 / (1.0 - 0.999 ** state["t"])
            with torch.no_grad():
                for t, g, mi, vi in zip(leaves, grads, m, v):
                    g = torch.zeros_like(t) if g is None else g
                    g = (g + 1e-5 * t).clamp_(-1.0, 1.0)
                    mi.mul_(0.9).add_(g, alpha=0.1)
                    vi.mul_(0.999).addcmul_(g, g, value=0.001)
                    t.sub_(1e-3 * (mi * c1) / ((vi * c2).sqrt() + 1e-8))
            return float(loss.detach())
    else:
        def step():
            e, f = O.energy_and_forces(p, h, x, mask=mask, atom_mask=am)
            return float(e.sum())
    return step


def cpu_flow_step_fn(seed, Bs, N):
    """cfg5 on the CPU: the oracle's flow_backward + the two log-probs (scripts/lj13_aug/run.py:39-43, no gradient)."""
    from oracle import sake_oracle as O
    ip = _init_module()
    gen = ip._generator(seed)
    p = {}
    for i in range(4):
        for nm in ("xv_%d" % i, "vx_%d" % i):
            p[nm] = {"sake_model": ip.init_model_params(gen, 3, H, 1, 4, n_heads=A),
                     "scale_mlp": {"layers_0": ip.dense_init(gen, 1, H), "layers_2": ip.dense_init(gen, H, 1, use_bias=False)}}
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(Bs, N, 3, generator=g); x = x - x.mean(-2, keepdim=True)
    v = torch.randn(Bs, N, 3, generator=g); v = v - v.mean(-2, keepdim=True)
    h = torch.zeros(Bs, N, 2)

    def step():
        with torch.no_grad():
            xb, vb, ld = O.flow_backward(p, h, x, v)
            return float((-O.centered_gaussian_log_prob(xb) - O.centered_gaussian_log_prob(vb) + ld).mean())
    return step


def cpu_arm(wl, depth, steps, warmup, sample_B):
    B, N, S, padded, n_min, mode, desc = WORKLOADS[wl]
    torch.set_num_threads(os.cpu_count())
    Bs = min(B, sample_B)
    if mode == "flow":
        step = cpu_flow_step_fn(0, Bs, N)
    else:
        h, x, mask, am, y, n_real = synth(2666, Bs, N, S, padded, n_min)
        T = lambda a: None if a is None else torch.tensor(a)
        step = cpu_step_fn(init_params_cpu(depth, S, 0), mode, (T(h), T(x), T(mask), T(am), T(y)))
    for _ in range(warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(steps):
        step()
    dt = (time.perf_counter() - t0) / steps
    return Bs / dt, dt, Bs


CPU_KIND = ("torch-eager fp32 restatement of the reference (oracle/sake_oracle.py; JAX not installable), padded batch with "
            "the reference's float mask, same optimiser chain as the GPU arm")


# ------------------------------------------------------------------------------------------------
class ClockSampler(threading.Thread):
    def __init__(self, gpu_index):
        super().__init__(daemon=True)
        self.samples, self.reasons, self.stop_flag, self.idx = [], set(), False, gpu_index

    def run(self):
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap,power.draw")
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        while not self.stop_flag:
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.idx), "--query-gpu=" + q,
                                      "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5).stdout
                parts = [s.strip() for s in out.strip().split(",")]
                self.samples.append((float(parts[0]), float(parts[1]), float(parts[6]) if len(parts) > 6 else 0.0))
                for n, v in zip(names, parts[2:6]):
                    if v.lower().startswith("active"):
                        self.reasons.add(n)
            except Exception:
                pass
            time.sleep(0.05)

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": sorted(self.reasons)}
        sm = sorted(s[0] for s in self.samples)
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": self.samples[0][1], "reasons": sorted(self.reasons),
                "samples": len(sm), "sm_mhz_min": sm[0], "power_w_max": max(s[2] for s in self.samples)}


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        p = json.load(open(path))
        return p.get("bf16_tflops_sustained", 1396.8), p.get("bf16_tflops", 1665.4), p.get("hbm_gbs", 6547.2), "measured"
    return 1400.0, 1590.0, 6650.0, "fallback"


def traffic_lookup(workload, engine, kernel):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of `kernel`, parsed from the committed `ncu --set full`
    capture of the CURRENT build (profiles/traffic.json, written by scripts/ncu_traffic.py); None when not captured."""
    path = os.path.join(ROOT, "profiles", "traffic.json")
    try:
        t = json.load(open(path))
        return t.get(workload, {}).get(engine, {}).get(kernel)
    except Exception:
        return None


KIND_NAMES = {1: "mix_fwd", 2: "mix_bwd", 3: "mix_dw", 4: "edge_fwd", 5: "edge_bwd", 6: "node_post", 7: "node_post_bwd",
              8: "dw_small", 9: "node_pre", 10: "attn_fwd", 11: "pair_reduce", 12: "node_pre_bwd", 13: "attn_bwd",
              14: "dense"}
# algorithmic FLOPs per pair (kinds 1-5, 8) / per atom (6, 7) of one launch
KIND_FLOP = {1: FLOP_MIX, 2: FLOP_MIX, 3: FLOP_MIX, 4: FLOP_EDGE, 5: 2 * FLOP_EDGE, 6: FLOP_NODE, 7: 2 * FLOP_NODE,
             8: FLOP_EDGE}


# ------------------------------------------------------------------------------------------------
# GPU jobs
# ------------------------------------------------------------------------------------------------
class ModelJob:
    """cfg1-4: a DenseSAKEModel energy+forces or training step through sake_b200.runner.ModelRunner."""

    def __init__(self, wl, args, rank, world, dev, B_override=None, allreduce=True):
        import sake_b200
        from sake_b200 import runner as R
        self.R = R
        B, N, S, padded, n_min, mode, desc = WORKLOADS[wl]
        shard = None
        if B_override is not None:
            # strong scaling: this rank's share of the ONE global batch (seed 2666), split by the dense pair cost
            # sum(n_real^2) for padded batches and contiguously otherwise (scripts/ani/run_gpu.py:54-56)
            from sake_b200.parallel import balanced_partition, shard_range
            full = synth(2666, B, N, S, padded, n_min)
            if padded and world > 1:
                shard = balanced_partition(full[5], world)[rank]
            else:
                b0, b1 = shard_range(B, world, rank)
                shard = np.arange(b0, b1)
            B = len(shard)
        self.B, self.N, self.mode, self.padded = B, N, mode, padded
        self.ragged = padded and args.padding == "ragged" and N <= 128
        model = sake_b200.DenseSAKEModel(hidden_features=H, out_features=1, depth=args.depth, engine=args.engine)
        self.run = R.ModelRunner(model, init_params_cpu(args.depth, S, 0), B, N, S, masked=padded and not self.ragged,
                                 ragged=self.ragged, train=(mode == "train"), device=dev,
                                 defer_dw=args.defer_dw, defer_reduce=args.defer_reduce)
        if shard is not None:
            h, x, mask, am, y, n_real = (None if t is None else t[shard] for t in full)
        else:
            # rank 0's batch is the N = 1 batch; every other rank draws its own molecules (species, coordinates,
            # targets) with the SAME multiset of sizes in another order: weak scaling compares equal work per GPU
            h, x, mask, am, y, n_real = synth(2666 + rank, B, N, S, padded, n_min)
            if rank and padded:
                n_real = np.random.default_rng(777 + rank).permutation(synth(2666, B, N, S, padded, n_min)[5]).astype(np.int32)
                am = (np.arange(N)[None, :] < n_real[:, None]).astype(np.float32)
                hz, xz = synth(2666 + rank, B, N, S, False, 0)[:2]
                h, x, mask = hz * am[..., None], xz * am[..., None], am[:, :, None] * am[:, None, :]
        self.n_real = n_real
        pin = lambda a: None if a is None else torch.tensor(a).pin_memory()
        self.hp, self.xp, self.yp = pin(h), pin(x), pin(y)
        self.mp, self.ap, self.np_ = (None, None, pin(n_real)) if self.ragged else (pin(mask), pin(am), None)
        self.load()
        self.allreduce = None
        self.bucketed = bool(getattr(args, "bucketed_allreduce", False))
        if world > 1 and mode == "train" and allreduce:
            from sake_b200.parallel import GradAllReducer
            self.allreduce = GradAllReducer()   # NCCL sum of the flat grad bucket; 1/world folded into the Adam kernel
        self.out_host = torch.empty(1 if mode == "train" else B * (1 + 3 * N), dtype=torch.float32).pin_memory()
        self.engine = self.run.engine
        self.real_pairs = float((n_real.astype(np.float64) ** 2).sum())
        self.real_atoms = float(n_real.sum())
        self.flop_step = args.depth * (self.real_pairs * FLOP_PAIR + self.real_atoms * FLOP_NODE) * (3 if mode == "train" else 2)
        self.flop_step_padded = args.depth * B * (N * N * FLOP_PAIR + N * FLOP_NODE) * (3 if mode == "train" else 2)

    def load(self):
        if self.ragged:
            self.run.load_inputs(self.hp, self.xp, target=self.yp, n_real=self.np_)
        else:
            self.run.load_inputs(self.hp, self.xp, self.mp, self.ap, self.yp)

    def step(self):
        if self.mode == "train":
            return self.run.train_step(self.allreduce, bucketed=self.bucketed)
        return self.run.energy_forces_step()

    def e2e_step(self):
        self.load()
        r = self.step()
        if self.mode == "train":
            self.out_host.copy_(r, non_blocking=True)
        else:
            self.out_host[:self.B].copy_(r[0], non_blocking=True)
            self.out_host[self.B:].copy_(r[1].reshape(-1), non_blocking=True)

    def capture(self):
        return self.run.capture()

    def graph_launches(self):
        return self.run.graph_launches if self.run.graphs else 0

    def h2d_bytes(self):
        return self.run.input_bytes()

    def d2h_bytes(self):
        return self.out_host.numel() * 4

    def unit_counts(self, kind):
        """units (pairs or atoms) one launch of a profiled kernel kind processes: real ones when ragged"""
        per_pair = self.real_pairs if (self.ragged or not self.padded) else float(self.B * self.N * self.N)
        per_atom = self.real_atoms if (self.ragged or not self.padded) else float(self.B * self.N)
        return per_atom if kind in (6, 7) else per_pair


class FlowJob:
    """cfg5: LJ13 AugmentedFlowModel(depth=4, mp_depth=4), B=4096: one log-likelihood pass per step."""

    def __init__(self, wl, args, rank, world, dev):
        from sake_b200.flow_runner import FlowRunner
        B, N, S, padded, n_min, mode, desc = WORKLOADS[wl]
        self.B, self.N, self.mode = B, N, mode
        self.run = FlowRunner(depth=4, mp_depth=4, B=B, N=N, D=3, seed=0, engine=args.engine, device=dev)
        g = torch.Generator().manual_seed(2666 + rank)
        x = torch.randn(B, N, 3, generator=g); x = x - x.mean(-2, keepdim=True)
        v = torch.randn(B, N, 3, generator=g); v = v - v.mean(-2, keepdim=True)
        self.xp, self.vp = x.pin_memory(), v.pin_memory()
        self.run.load_inputs(self.xp, self.vp)
        self.out_host = torch.empty(B, dtype=torch.float32).pin_memory()
        self.engine = self.run.engine
        self.ragged, self.padded, self.allreduce = False, False, None
        n_models, n_layers = 8, 4
        self.real_pairs = float(B * (N + 1) ** 2) * n_models          # per "launch family": bench divides by launches
        self.real_atoms = float(B * (N + 1)) * n_models
        self.flop_step = n_models * n_layers * B * ((N + 1) ** 2 * FLOP_PAIR + (N + 1) * FLOP_NODE)
        self.flop_step_padded = self.flop_step
        self.n_real = np.full(B, N + 1, dtype=np.int32)

    def step(self):
        self.run.restore_inputs()            # the pass transforms (x, v) in place: every step sees the loaded batch
        return self.run.log_likelihood_step()

    def e2e_step(self):
        self.run.load_inputs(self.xp, self.vp)
        self.out_host.copy_(self.step(), non_blocking=True)

    def capture(self):
        return self.run.capture()

    def graph_launches(self):
        return self.run.graph_launches if self.run.graph is not None else 0

    def h2d_bytes(self):
        return 2 * self.B * self.N * 3 * 4

    def d2h_bytes(self):
        return self.B * 4

    def unit_counts(self, kind):
        return float(self.B * (self.N + 1)) if kind in (6, 7) else float(self.B * (self.N + 1) ** 2)


def timed_region(job, steps, flush, dist, dev):
    """EXACTLY `steps` steps, one CUDA event pair per step, L2 flushed between them; returns per-rank summed ms."""
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
    if dist is not None:
        dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for a, b in evs:
        if flush is not None:
            flush.zero_()
        a.record()
        job.step()
        b.record()
    torch.cuda.synchronize()
    if dist is not None:
        dist.barrier()
    wall = time.perf_counter() - t0
    dev_ms = sum(a.elapsed_time(b) for a, b in evs)
    tmax = torch.tensor([dev_ms], device=dev, dtype=torch.float64)
    if dist is not None:
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
    return tmax.item(), wall


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=None, help="timed steps (default: as many as fill ~2.5 s)")
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--workload", default="cfg2", choices=list(WORKLOADS))
    ap.add_argument("--engine", default="auto", choices=["auto", "fp32", "tf32x3", "bf16", "f16x2"])
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--depth", type=int, default=4)
    ap.add_argument("--padding", default="ragged", choices=["ragged", "masked"],
                    help="padded workloads: compute real atoms only (n_real-packed tiles) or all N^2 pairs with the float mask")
    ap.add_argument("--cpu-sample", type=int, default=16, help="molecules in the CPU-baseline sample batch")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-graphs", action="store_true", help="launch every kernel eagerly instead of replaying the step as a CUDA graph")
    ap.add_argument("--defer-dw", action="store_true",
                    help="run the weight-gradient contractions on the library's side stream (measured: no gain, see DESIGN.md)")
    ap.add_argument("--defer-reduce", action="store_true",
                    help="run the weight-gradient partial-sum reduction on the library's side stream (measured: no gain)")
    ap.add_argument("--bucketed-allreduce", action="store_true",
                    help="training on several GPUs: all-reduce per-layer gradient buckets between backward segments instead of one all-reduce after the backward")
    ap.add_argument("--no-strong", action="store_true", help="skip the fixed-total-size cfg3 / cfg4 records")
    ap.add_argument("--no-sustained", action="store_true", help="skip the extra timed rounds that extend the run to ~2.5 s")
    args = ap.parse_args()
    # The contract is ONE JSON line on stdout.  Libraries (NCCL prints its version banner to stdout) must not
    # get in the way: fd 1 is pointed at stderr for the whole run and the line goes to the saved descriptor.
    sys.stdout.flush()
    real_stdout = os.dup(1)
    os.dup2(2, 1)

    def emit(obj):
        os.write(real_stdout, (json.dumps(obj) + "\n").encode())
    B, N, S, padded, n_min, mode, desc = WORKLOADS[args.workload]
    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    local_rank = int(os.environ.get("LOCAL_RANK", 0))
    config = make_config(args.workload, args.depth, world, args.padding, not args.no_strong)

    # ---------------- reference arm: CPU oracle port, rank 0 only ----------------
    if args.impl == "reference":
        if rank != 0:
            return
        # honour K / W but keep the whole run within a few minutes (the sample batch bounds each step)
        steps = max(1, min(args.steps or 20, 30))
        warm = max(1, min(args.warmup, 5))
        sample_B = min(args.cpu_sample, 4) if mode == "flow" else args.cpu_sample
        val, dt, Bs = cpu_arm(args.workload, args.depth, steps, warm, sample_B)
        sample = f"{Bs} of {B} molecules of the same workload per step, {steps} timed steps"
        emit({
            "impl": "reference", "metric": "molecules_per_sec", "value": val, "unit": "molecules/s",
            "n_gpus": args.gpus, "steps": steps, "warmup": warm, "ms_per_step": dt * 1e3, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": config,
            "cpu_baseline": {"value": val, "unit": "molecules/s", "cores": os.cpu_count(), "kind": "port",
                             "sample": sample + "; " + CPU_KIND},
            "e2e": {"value": val, "unit": "molecules/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}})
        return

    # ---------------- our arm ----------------
    import sake_b200  # noqa: F401
    from sake_b200 import runner as R
    from sake_b200._lib import lib
    assert torch.cuda.is_available(), "bench.py needs a CUDA device (no CPU fallback)"
    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    dev = torch.device("cuda", local_rank)
    job = (FlowJob if mode == "flow" else ModelJob)(args.workload, args, rank, world, dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    warm = max(args.warmup, 3)
    for _ in range(warm):
        job.step()
    torch.cuda.synchronize()

    # ---- per-kernel durations for the roofline: a short EAGER pass of the same steps with the library's own
    # event pairs around every big launch (a graph replay makes no host calls, so they cannot be recorded there)
    n_prof = 3
    R.profile_begin(4096)
    g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    g0.record()
    for _ in range(n_prof):
        job.step()
    g1.record()
    torch.cuda.synchronize()
    prof, prof_ms = R.profile_collect(4096), g0.elapsed_time(g1)
    eager_launches0 = lib.sake_launch_count()
    job.step()
    torch.cuda.synchronize()
    launches_per_step = int(lib.sake_launch_count() - eager_launches0)
    graph_note = "off (every kernel launched eagerly)"
    if not args.no_graphs:
        n_graph = job.capture()
        graph_note = f"step body replayed as one CUDA graph ({n_graph} library launches per replay)"
        for _ in range(3):
            job.step()
        torch.cuda.synchronize()

    # ---- steps: the contract's K, or enough to fill ~2.5 s ------------------------------------------------
    if args.steps is None:
        c0, c1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        c0.record()
        for _ in range(3):
            job.step()
        c1.record()
        torch.cuda.synchronize()
        est = torch.tensor([c0.elapsed_time(c1) / 3], device=dev, dtype=torch.float64)
        if dist is not None:
            dist.all_reduce(est, op=dist.ReduceOp.MAX)
        args.steps = int(min(4000, max(20, math.ceil(2500.0 / max(est.item(), 1e-3)))))

    # ---- device-resident timed region -------------------------------------------------------------
    sampler = ClockSampler(local_rank)
    sampler.start()
    launches0 = lib.sake_launch_count()
    replays0 = getattr(job.run, "graph_replays", 0)
    dev_ms, t_wall = timed_region(job, args.steps, flush, dist, dev)
    launches = int(lib.sake_launch_count() - launches0) + (getattr(job.run, "graph_replays", 0) - replays0) * job.graph_launches()
    ms_step = dev_ms / args.steps
    value = world * B / (ms_step * 1e-3)
    # sustained: more rounds of the same K steps until ~4 s of timed work have run, so that the clocks are sampled
    # >= 20 times under load and the number is not a 0.1 s burst
    sustained = None
    if not args.no_sustained and dev_ms < 4000.0:
        rounds = int(min(400, math.ceil((4000.0 - dev_ms) / max(dev_ms, 1e-3))))
        tot = 0.0
        for _ in range(rounds):
            r_ms, _ = timed_region(job, args.steps, flush, dist, dev)
            tot += r_ms
        sustained = {"rounds": rounds, "steps": rounds * args.steps, "ms_per_step": tot / (rounds * args.steps),
                     "value": world * B / (tot / (rounds * args.steps) * 1e-3), "timed_s": tot * 1e-3}

    # ---- end-to-end: pinned host inputs -> H2D -> step -> D2H result, every step ---------------------
    for _ in range(2):
        job.e2e_step()
    torch.cuda.synchronize()
    if dist is not None:
        dist.barrier()
    ea, eb = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    ea.record()
    for _ in range(args.steps):
        job.e2e_step()
    eb.record()
    torch.cuda.synchronize()
    e2e_wall = time.perf_counter() - t0
    e2e_t = torch.tensor([max(ea.elapsed_time(eb) * 1e-3, e2e_wall)], device=dev, dtype=torch.float64)
    if dist is not None:
        dist.all_reduce(e2e_t, op=dist.ReduceOp.MAX)
    e2e_val = world * B * args.steps / e2e_t.item()
    sampler.stop_flag = True
    sampler.join(timeout=2)

    # everything the JSON line needs from the headline job (it is released before the strong-scaling jobs)
    head = {"engine": job.engine, "real_pairs": job.real_pairs, "flop_step": job.flop_step,
            "flop_step_padded": job.flop_step_padded, "h2d": job.h2d_bytes(), "d2h": job.d2h_bytes(),
            "units": {k: job.unit_counts(k) for k in KIND_NAMES},
            "padding": ("ragged: n_real-packed tiles, real atoms only (sum n^2 pairs)" if job.ragged else
                        ("masked: all N^2 pairs of the padded width times the float mask" if job.padded else "none (no padding in this workload)")),
            "flow": getattr(job, "flow_extra", None)}
    if mode == "flow":
        # the sampling direction of the same flow (f_forward), timed the same way, beside the likelihood pass
        s_ms, _ = timed_region(type("S", (), {"step": staticmethod(lambda: (job.run.restore_inputs(), job.run.sample_step()))})(), max(3, min(args.steps, 20)), flush, dist, dev)
        head["flow"] = {"sampling_ms_per_step": s_ms / max(3, min(args.steps, 20)),
                        "sampling_molecules_per_sec": world * B / (s_ms / max(3, min(args.steps, 20)) * 1e-3),
                        "layers_per_pass": 32, "atoms_per_model_call": N + 1}

    # ---- strong scaling of the named multi-GPU configs (fixed total size) --------------------------------
    strong = None
    if not args.no_strong and mode != "flow":
        strong = {}
        for swl in ("cfg3", "cfg4"):
            Bt = WORKLOADS[swl][0]
            job = None
            torch.cuda.empty_cache()
            sj = ModelJob(swl, args, rank, world, dev, B_override=Bt)
            for _ in range(3):
                sj.step()
            if not args.no_graphs:
                sj.capture()
                sj.step()
            torch.cuda.synchronize()
            ssteps = 10
            s_ms, _ = timed_region(sj, ssteps, flush, dist, dev)
            strong[swl] = {"molecules_total": Bt, "molecules_this_rank": sj.B, "mode": sj.mode,
                           "split": "one global batch (seed 2666) split over the ranks: balanced by sum(n_real^2)" if sj.padded
                                    else "one global batch (seed 2666) split contiguously",
                           "ms_per_step": s_ms / ssteps, "value": Bt / (s_ms / ssteps * 1e-3), "unit": "molecules/s",
                           "steps": ssteps, "collective": "ncclAllReduce of the flat gradient bucket" if sj.allreduce else "none",
                           "padding": "ragged" if sj.ragged else ("masked" if sj.padded else "none")}
            job = sj
        job = None

    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant kernel ---------------------------------------------------------------
    sus, burst, hbm, how = measured_peaks()
    by_kind = {}
    for ms, kind, pairs in prof:
        d = by_kind.setdefault(kind, [0.0, 0])
        d[0] += ms
        d[1] += 1
    roofline = None
    if by_kind:
        dom = max((k for k in by_kind if k in KIND_FLOP), key=lambda k: by_kind[k][0])
        tot, cnt = by_kind[dom]
        avg_ms = tot / cnt
        units = head["units"][dom]
        flop = KIND_FLOP[dom] * units
        achieved = flop / (avg_ms * 1e-3) / 1e12
        name = KIND_NAMES.get(dom, str(dom))
        roofline = {"bound": "tensor", "kernel": name, "achieved": achieved, "peak": burst, "unit": "TFLOP/s",
                    "frac": achieved / burst, "peak_sustained": sus, "frac_sustained": achieved / sus,
                    "traffic": traffic_lookup(args.workload, head["engine"], name),
                    "peak_source": f"bf16 dense burst (kernel-level timing at full clocks), {how} (MEASURED_PEAKS.json); "
                                   "a parity engine issues 3 MMAs per product: its ceiling is 1/3 (forward) or 1/6 "
                                   "(backward: recompute + dX) of this peak",
                    "avg_launch_ms": avg_ms, "launches_timed": cnt, "algorithmic_flop_per_launch": flop,
                    "units_per_launch": units, "units": "real atom pairs" if dom not in (6, 7) else "real atoms",
                    "timing": f"CUDA events around each launch on the launch stream, eager pass of {n_prof} steps",
                    "share_of_step": {KIND_NAMES.get(k, str(k)): by_kind[k][0] / prof_ms for k in by_kind}}
    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        sample_B = min(args.cpu_sample, 2) if mode == "flow" else args.cpu_sample
        val, dt, Bs = cpu_arm(args.workload, args.depth, 2, 1, sample_B)
        cpu = {"value": val, "unit": "molecules/s", "cores": os.cpu_count(), "kind": "port",
               "sample": f"{Bs} of {B} molecules per step, 2 timed steps; " + CPU_KIND}
    line = {
        "metric": "molecules_per_sec", "value": value, "unit": "molecules/s", "n_gpus": world, "steps": args.steps,
        "warmup": warm, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": {"fp32": "f32", "tf32x3": "tf32x3 (fp32-parity split)", "bf16": "bf16",
                  "f16x2": "f16x2 (fp32-parity split, fp16 hi/lo)"}[head["engine"]],
        "data": "synthetic", "config": config,
        "padding": head["padding"],
        "atom_pairs_per_sec": world * B * N * N * args.depth / (ms_step * 1e-3) if mode != "flow" else None,
        "real_atom_pairs_per_sec": world * head["real_pairs"] * (4 if mode == "flow" else args.depth) / (ms_step * 1e-3),
        "algorithmic_tflops": world * head["flop_step"] / (ms_step * 1e-3) / 1e12,
        "algorithmic_tflops_padded_count": world * head["flop_step_padded"] / (ms_step * 1e-3) / 1e12,
        "engine": head["engine"], "gpu_launches": int(launches), "launches_per_step_eager": launches_per_step,
        "wall_s_timed_region": t_wall, "sustained": sustained, "cuda_graph": graph_note,
        "clocks": sampler.summary(),
        "e2e": {"value": e2e_val, "unit": "molecules/s", "h2d_bytes_per_step": head["h2d"], "d2h_bytes_per_step": head["d2h"]},
        "roofline": roofline, "cpu_baseline": cpu, "strong": strong,
    }
    if mode == "flow":
        line["flow"] = head["flow"]
    emit(line)
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
