"""world_size-2 gloo tests (CPU) of the host-side data-parallel logic: batch sharding, cost-balanced
partitioning, and the gradient all-reduce = lax.pmean semantics (scripts/ani/run_gpu.py:124-132),
checked against the oracle: mean of per-shard parameter gradients == gradient of the full batch."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out):
    import importlib.util
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, root)
    # parallel.py is pure host logic: import it without the CUDA library
    spec = importlib.util.spec_from_file_location("sake_parallel", os.path.join(root, "sake_b200", "parallel.py"))
    par = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(par)
    from oracle import sake_oracle as O
    from tests import golden_util as G
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    g = G.load("model_h16_d4_n5")
    p = O.params_to(G.unflatten(g["params"]), torch.float64)
    flat = O.tree_flatten(p)
    for t in flat.values():
        t.requires_grad_(True)
    rng = np.random.default_rng(0)
    B = 6
    h = torch.tensor(rng.uniform(size=(B, 5, 16)))
    x = torch.tensor(rng.standard_normal((B, 5, 3)))
    y = torch.tensor(rng.standard_normal(B))
    b0, b1 = par.shard_range(B, world, rank)
    e = O.energy(p, h[b0:b1], x[b0:b1])
    loss = (e - y[b0:b1]).abs().mean()
    grads = torch.autograd.grad(loss, list(flat.values()), allow_unused=True)
    bucket = torch.cat([(gr if gr is not None else torch.zeros_like(t)).reshape(-1)
                        for gr, t in zip(grads, flat.values())])
    # bucketed form (one all-reduce per layer, started as the layers finish): same result as the single call
    pieces = bucket.clone()
    red = par.GradAllReducer()
    n3 = pieces.numel() // 3
    for a, b in ((0, n3), (n3, 2 * n3), (2 * n3, pieces.numel())):
        red.start(pieces[a:b])
    s2 = red.finish()
    scale = par.GradAllReducer()(bucket)
    assert s2 == scale and torch.equal(pieces, bucket)
    bucket = bucket * scale
    tmax = par.max_over_ranks(float(rank + 1), "cpu")
    if rank == 0:
        # full-batch gradient (equal shard sizes -> mean of shard means == full mean)
        e_all = O.energy(p, h, x)
        loss_all = (e_all - y).abs().mean()
        g_all = torch.autograd.grad(loss_all, list(flat.values()), allow_unused=True)
        ref = torch.cat([(gr if gr is not None else torch.zeros_like(t)).reshape(-1)
                         for gr, t in zip(g_all, flat.values())])
        out.put((float((bucket - ref).abs().max()), float(ref.abs().max()), tmax))
    dist.barrier()
    dist.destroy_process_group()


def test_grad_allreduce_is_pmean_gloo():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    err, scale, tmax = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert err < 1e-12 * max(1.0, scale), (err, scale)
    assert tmax == 2.0


def test_shard_range_and_balance():
    import importlib.util
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    spec = importlib.util.spec_from_file_location("sake_parallel", os.path.join(root, "sake_b200", "parallel.py"))
    par = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(par)
    for n, w in [(1024, 8), (10, 4), (7, 8), (0, 2)]:
        spans = [par.shard_range(n, w, r) for r in range(w)]
        assert spans[0][0] == 0 and spans[-1][1] == n
        assert all(spans[i][1] == spans[i + 1][0] for i in range(w - 1))
        sizes = [b - a for a, b in spans]
        assert max(sizes) - min(sizes) <= 1
    n_real = np.random.default_rng(1).integers(9, 30, 256)
    parts = par.balanced_partition(n_real, 8)
    assert sorted(np.concatenate(parts).tolist()) == list(range(256))
    loads = [float((n_real[p].astype(np.float64) ** 2).sum()) for p in parts]
    assert max(loads) / min(loads) < 1.02
