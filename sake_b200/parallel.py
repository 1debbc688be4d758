"""Data parallelism for the SAKE hot path: molecules are independent, so a batch is split over
ranks with NO data-path collective (inference), and training adds exactly one collective — the
all-reduce of the flat weight-gradient bucket, `lax.pmean(grads, "batch")` in the reference
(scripts/ani/run_gpu.py:124-132, scripts/qm9_tpu/run.py:89-95).  One process per GPU;
torch.distributed (NCCL over NVLink on GPUs, gloo in the CPU tests) is the plumbing."""
import numpy as np
import torch
import torch.distributed as dist


def shard_range(n_items, world, rank):
    """Contiguous split of the leading batch axis, as the reference's host-side reshape to
    [n_devices, batch, ...] (scripts/ani/run_gpu.py:54-56).  Returns (begin, end)."""
    base, rem = divmod(int(n_items), int(world))
    begin = rank * base + min(rank, rem)
    return begin, begin + base + (1 if rank < rem else 0)


def balanced_partition(n_real, world):
    """Ragged / padded batches: balance ranks by the dense pair cost sum(n_real^2) instead of by
    molecule count (greedy longest-processing-time).  Returns a list of index arrays, one per rank."""
    cost = np.asarray(n_real, dtype=np.float64) ** 2
    order = np.argsort(-cost, kind="stable")
    loads = np.zeros(world)
    parts = [[] for _ in range(world)]
    for i in order:
        r = int(np.argmin(loads))
        parts[r].append(int(i))
        loads[r] += cost[i]
    return [np.asarray(sorted(p), dtype=np.int64) for p in parts]


class GradAllReducer:
    """All-reduce(sum) of one flat fp32 gradient bucket; returns the 1/world factor that the fused
    optimiser kernel (sake_adam_step, grad_scale) applies, i.e. lax.pmean without an extra pass."""

    def __init__(self, group=None):
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1

    def __call__(self, flat_grads: torch.Tensor) -> float:
        if self.world > 1:
            dist.all_reduce(flat_grads, op=dist.ReduceOp.SUM, group=self.group)
        return 1.0 / self.world

    # Bucketed form: start(bucket) as soon as one layer's gradients are final — the collective runs on the
    # backend's own stream (NCCL: ordered after the work already enqueued on the current stream) while the backward
    # of the next layer keeps the SMs busy — and finish() before the optimiser.  Buckets are views of the one flat
    # gradient vector, so the result is the same all-reduce, cut at layer boundaries.
    def start(self, bucket: torch.Tensor) -> None:
        if self.world > 1 and bucket.numel() > 0:
            if not hasattr(self, "_works"):
                self._works = []
            self._works.append(dist.all_reduce(bucket, op=dist.ReduceOp.SUM, group=self.group, async_op=True))

    def finish(self) -> float:
        for w in getattr(self, "_works", []):
            w.wait()                       # CUDA backends: makes the current stream wait, no host block
        self._works = []
        return 1.0 / self.world


def max_over_ranks(value: float, device) -> float:
    """Device-side timing is reported as the max over ranks."""
    t = torch.tensor([value], dtype=torch.float64, device=device)
    if dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())
