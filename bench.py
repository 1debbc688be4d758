#!/usr/bin/env python
"""Headline benchmark: SAKE DenseSAKEModel molecules/s (and atom-pairs/s) on B200.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--workload cfg1|cfg2|cfg3|cfg4]
                  [--engine auto|fp32|tf32x3|bf16] [--impl ours|reference]

One "step" = one pass of the hot path over one batch of synthetic molecules:
  cfg1  MD17-aspirin-shaped, N=21, B=32, energy+forces            (the reference's CPU-runnable case)
  cfg2  QM9-shaped, padded to N=29, B=256, training step           (DEFAULT: the single-GPU config)
  cfg3  ANI-1x-shaped, padded to N=63, B=1024, energy+forces
  cfg4  OC20-slab-shaped, N=200, B=64, training step (+ grad all-reduce when N_gpus > 1)
Weak scaling: every rank processes its own batch of B molecules; value = N_gpus * B / step time.
Prints ONE JSON line (see the driver contract in the task statement / DESIGN.md section 6).
"""
import argparse
import json
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
}
H, A, K, C = 64, 4, 50, 256
FLOP_PAIR = 2 * ((K + 1) * H + H * H + H * A + C * C + C)     # 146 816 (factored form)
FLOP_NODE = 2 * 67904                                         # 135 808
FLOP_MIX = 2 * C * C                                          # the x_mixing contraction alone, per pair


def synth(seed, B, N, S, padded, n_min):
    """Synthetic molecules of the named shape: normal coords at molecular density,
    uniform species one-hots, QM9-style padding (scripts/qm9/run.py:23-24,35)."""
    rng = np.random.default_rng(seed)
    x = (rng.standard_normal((B, N, 3)) * 0.62 * N ** (1.0 / 3.0)).astype(np.float32)
    z = rng.integers(0, S, (B, N))
    h = np.eye(S, dtype=np.float32)[z]
    y = rng.standard_normal(B).astype(np.float32)
    if not padded:
        return h, x, None, None, y, np.full(B, N)
    n_real = rng.integers(n_min, N + 1, B)
    am = (np.arange(N)[None, :] < n_real[:, None]).astype(np.float32)
    return h * am[..., None], x * am[..., None], am[:, :, None] * am[:, None, :], am, y, n_real


def init_params_cpu(depth, S, seed):
    """flax-default initialisation, seeded; identical on every rank."""
    from sake_b200.layers import _generator, dense_init, init_layer_params
    gen = _generator(seed)
    p = {"embedding_in": dense_init(gen, S, H),
         "embedding_out": {"layers_0": dense_init(gen, H, H), "layers_2": dense_init(gen, H, 1)}}
    has_v = False
    for i in range(depth):
        p["d%d" % i] = init_layer_params(gen, H, H, H, A, True, has_v)
        has_v = True
    return p


# ------------------------------------------------------------------------------------------------
# CPU arm: the oracle port of the reference (JAX itself is not installable in this image)
# ------------------------------------------------------------------------------------------------
def cpu_step_fn(params_cpu, mode, batch):
    from oracle import sake_oracle as O
    h, x, mask, am, y = batch
    p = O.tree_map(lambda t: t.clone().float(), params_cpu)
    leaves = list(O.tree_flatten(p).values())
    if mode == "train":
        for t in leaves:
            t.requires_grad_(True)
        opt = torch.optim.Adam(leaves, lr=1e-3)

        def step():
            opt.zero_grad(set_to_none=True)
            e = O.energy(p, h, x, mask=mask, atom_mask=am)
            loss = (e - y).abs().mean()
            loss.backward()
            opt.step()
            return float(loss.detach())
    else:
        def step():
            e, f = O.energy_and_forces(p, h, x, mask=mask, atom_mask=am)
            return float(e.sum())
    return step


def cpu_arm(wl, depth, steps, warmup, sample_B):
    B, N, S, padded, n_min, mode, desc = WORKLOADS[wl]
    torch.set_num_threads(os.cpu_count())
    Bs = min(B, sample_B)
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


# ------------------------------------------------------------------------------------------------
class ClockSampler(threading.Thread):
    def __init__(self, gpu_index):
        super().__init__(daemon=True)
        self.samples, self.reasons, self.stop_flag, self.idx = [], set(), False, gpu_index

    def run(self):
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        while not self.stop_flag:
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.idx), "--query-gpu=" + q,
                                      "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5).stdout
                parts = [s.strip() for s in out.strip().split(",")]
                self.samples.append((float(parts[0]), float(parts[1])))
                for n, v in zip(names, parts[2:]):
                    if v.lower().startswith("active"):
                        self.reasons.add(n)
            except Exception:
                pass
            time.sleep(0.1)

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": sorted(self.reasons)}
        sm = sorted(s[0] for s in self.samples)
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": self.samples[0][1], "reasons": sorted(self.reasons),
                "samples": len(sm)}


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        p = json.load(open(path))
        return p.get("bf16_tflops_sustained", 1396.8), p.get("bf16_tflops", 1665.4), p.get("hbm_gbs", 6547.2), "measured"
    return 1400.0, 1590.0, 6650.0, "fallback"


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--workload", default="cfg2", choices=list(WORKLOADS))
    ap.add_argument("--engine", default="auto", choices=["auto", "fp32", "tf32x3", "bf16", "f16x2"])
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--depth", type=int, default=4)
    ap.add_argument("--cpu-sample", type=int, default=16, help="molecules in the CPU-baseline sample batch")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--graphs", action="store_true",
                    help="replay the step as a CUDA graph (per-kernel roofline timings then come from a separate eager pass)")
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
    flop_step = args.depth * B * (N * N * FLOP_PAIR + N * FLOP_NODE) * (3 if mode == "train" else 2)
    config = {"workload": f"{args.workload}: {desc}; DenseSAKEModel(hidden=64, heads=4, depth={args.depth})",
              "molecules_per_gpu": B, "atoms_padded": N, "mode": mode, "parallelism": f"dp{world}" if world > 1 else "single"}

    # ---------------- reference arm: CPU oracle port, rank 0 only ----------------
    if args.impl == "reference":
        if rank != 0:
            return
        # honour K / W but keep the whole run within a few minutes (the sample batch bounds each step)
        steps = max(1, min(args.steps, 30))
        warm = max(1, min(args.warmup, 5))
        val, dt, Bs = cpu_arm(args.workload, args.depth, steps, warm, args.cpu_sample)
        sample = f"{Bs} of {B} molecules of the same workload per step, {steps} timed steps"
        emit({
            "impl": "reference", "metric": "molecules_per_sec", "value": val, "unit": "molecules/s",
            "n_gpus": args.gpus, "steps": steps, "warmup": warm, "ms_per_step": dt * 1e3, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": config,
            "cpu_baseline": {"value": val, "unit": "molecules/s", "cores": os.cpu_count(), "kind": "port",
                             "sample": sample + " (torch-eager fp32 restatement of the reference; JAX not installable)"},
            "e2e": {"value": val, "unit": "molecules/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}})
        return

    # ---------------- our arm ----------------
    import sake_b200
    from sake_b200 import runner as R
    from sake_b200._lib import lib
    assert torch.cuda.is_available(), "bench.py needs a CUDA device (no CPU fallback)"
    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    dev = torch.device("cuda", local_rank)
    model = sake_b200.DenseSAKEModel(hidden_features=H, out_features=1, depth=args.depth, engine=args.engine)
    params = init_params_cpu(args.depth, S, 0)
    run = R.ModelRunner(model, params, B, N, S, masked=padded, train=(mode == "train"), device=dev)
    h, x, mask, am, y, n_real = synth(2666 + rank, B, N, S, padded, n_min)
    pin = lambda a: None if a is None else torch.tensor(a).pin_memory()
    hp, xp, mp, ap_, yp = pin(h), pin(x), pin(mask), pin(am), pin(y)
    run.load_inputs(hp, xp, mp, ap_, yp)

    allreduce = None
    if world > 1 and mode == "train":
        from sake_b200.parallel import GradAllReducer
        allreduce = GradAllReducer()       # NCCL sum of the flat grad bucket; 1/world folded into the Adam kernel

    def step():
        if mode == "train":
            return run.train_step(allreduce)
        return run.energy_forces_step()

    flush = None
    ws_bytes = run.hbm_bytes
    if ws_bytes < 2 * 126e6:
        flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    config["l2"] = ("flushed between timed steps (256 MiB memset)" if flush is not None
                    else f"per-step working set {ws_bytes / 1e6:.0f} MB > 126 MB L2")

    for _ in range(max(args.warmup, 3)):
        step()
    torch.cuda.synchronize()
    graph_prof = None
    if args.graphs:
        # per-kernel durations for the roofline: a short eager, profiled pass of the same steps (a graph replay
        # makes no host calls, so the library's per-launch events cannot be recorded inside it)
        R.profile_begin(64 * 3)
        g0 = torch.cuda.Event(enable_timing=True); g1 = torch.cuda.Event(enable_timing=True)
        g0.record()
        for _ in range(3):
            step()
        g1.record()
        torch.cuda.synchronize()
        graph_prof = (R.profile_collect(64 * 3), g0.elapsed_time(g1) / 3)
        run.capture()
        config["cuda_graph"] = f"step body replayed as one CUDA graph ({run.graph_launches} library launches per replay)"
        for _ in range(3):
            step()
        torch.cuda.synchronize()

    # ---- device-resident timed region -------------------------------------------------------------
    sampler = ClockSampler(local_rank)
    sampler.start()
    R.profile_begin(64 * args.steps)
    launches0 = lib.sake_launch_count()
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    if dist is not None:
        dist.barrier()
    torch.cuda.synchronize()
    t_wall0 = time.perf_counter()
    for a, b in evs:
        if flush is not None:
            flush.zero_()
        a.record()
        step()
        b.record()
    torch.cuda.synchronize()
    if dist is not None:
        dist.barrier()
    t_wall = time.perf_counter() - t_wall0
    launches = lib.sake_launch_count() - launches0
    prof = R.profile_collect(64 * args.steps)
    if args.graphs:
        launches += args.steps * run.graph_launches      # kernels inside the replayed graphs
    dev_ms = sum(a.elapsed_time(b) for a, b in evs)
    tmax = torch.tensor([dev_ms], device=dev, dtype=torch.float64)
    if dist is not None:
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
    ms_step = tmax.item() / args.steps
    value = world * B / (ms_step * 1e-3)

    # ---- end-to-end: pinned host inputs -> H2D -> step -> D2H result, every step ---------------------
    out_host = torch.empty(1 if mode == "train" else B * (1 + 3 * N), dtype=torch.float32).pin_memory()
    def e2e_step():
        run.load_inputs(hp, xp, mp, ap_, yp)
        r = step()
        if mode == "train":
            out_host.copy_(r, non_blocking=True)
        else:
            out_host[:B].copy_(r[0], non_blocking=True)
            out_host[B:].copy_(r[1].reshape(-1), non_blocking=True)
    for _ in range(2):
        e2e_step()
    torch.cuda.synchronize()
    if dist is not None:
        dist.barrier()
    ea, eb = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    ea.record()
    for _ in range(args.steps):
        e2e_step()
    eb.record()
    torch.cuda.synchronize()
    e2e_wall = time.perf_counter() - t0
    e2e_t = torch.tensor([max(ea.elapsed_time(eb) * 1e-3, e2e_wall)], device=dev, dtype=torch.float64)
    if dist is not None:
        dist.all_reduce(e2e_t, op=dist.ReduceOp.MAX)
    e2e_val = world * B * args.steps / e2e_t.item()
    sampler.stop_flag = True
    sampler.join(timeout=2)

    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant kernel (x_mixing GEMM family) ---------------------------------------
    sus, burst, hbm, how = measured_peaks()
    by_kind = {}
    prof_step_ms = ms_step * args.steps
    if graph_prof is not None:
        prof, prof_step_ms = graph_prof[0], graph_prof[1] * 3
    for ms, kind, pairs in prof:
        d = by_kind.setdefault(kind, [0.0, 0, pairs])
        d[0] += ms
        d[1] += 1
    kind_names = {1: "mix_fwd", 2: "mix_bwd", 3: "mix_dw"}
    # dram__bytes_read.sum + dram__bytes_write.sum per launch of the kernel, from the committed `ncu --set full`
    # captures (profiles/r01e_ncu_cfg2_auto.txt, profiles/r01e_ncu_mix_bwd_cfg3_f16x2.txt); null where not captured
    traffic_table = {("cfg2", "f16x2", "mix_bwd"): 98.501120e6 + 231.125248e6,
                     ("cfg3", "f16x2", "mix_bwd"): 1.453833e9 + 1.117801e9}
    roofline = None
    if by_kind:
        dom = max(by_kind, key=lambda k: by_kind[k][0])
        tot, cnt, pairs = by_kind[dom]
        avg_ms = tot / cnt
        achieved = FLOP_MIX * pairs / (avg_ms * 1e-3) / 1e12
        roofline = {"bound": "tensor", "kernel": kind_names.get(dom, str(dom)), "achieved": achieved, "peak": sus,
                    "unit": "TFLOP/s", "frac": achieved / sus,
                    "traffic": traffic_table.get((args.workload, run.engine, kind_names.get(dom, str(dom)))),
                    "peak_source": f"bf16 dense sustained, {how} (MEASURED_PEAKS.json)",
                    "avg_launch_ms": avg_ms, "launches_timed": cnt, "algorithmic_flop_per_launch": FLOP_MIX * pairs,
                    "share_of_step": {kind_names.get(k, str(k)): by_kind[k][0] / prof_step_ms for k in by_kind}}
    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        val, dt, Bs = cpu_arm(args.workload, args.depth, 2, 1, args.cpu_sample)
        cpu = {"value": val, "unit": "molecules/s", "cores": os.cpu_count(), "kind": "port",
               "sample": f"{Bs} of {B} molecules per step, 2 timed steps, torch-eager fp32 oracle port (JAX not installable)"}
    d2h = out_host.numel() * 4
    line = {
        "metric": "molecules_per_sec", "value": value, "unit": "molecules/s", "n_gpus": world, "steps": args.steps,
        "warmup": max(args.warmup, 3), "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": {"fp32": "f32", "tf32x3": "tf32x3 (fp32-parity split)", "bf16": "bf16",
                  "f16x2": "f16x2 (fp32-parity split, fp16 hi/lo)"}[run.engine],
        "data": "synthetic", "config": config,
        "atom_pairs_per_sec": world * B * N * N * args.depth / (ms_step * 1e-3),
        "real_atom_pairs_per_sec": world * float((n_real.astype(np.float64) ** 2).sum()) * args.depth / (ms_step * 1e-3),
        "algorithmic_tflops": world * flop_step / (ms_step * 1e-3) / 1e12,
        "engine": run.engine, "gpu_launches": int(launches), "wall_s_timed_region": t_wall,
        "clocks": sampler.summary(),
        "e2e": {"value": e2e_val, "unit": "molecules/s", "h2d_bytes_per_step": run.input_bytes(),
                "d2h_bytes_per_step": d2h},
        "roofline": roofline, "cpu_baseline": cpu,
    }
    emit(line)
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
