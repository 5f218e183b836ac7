"""Batch-dimension data parallelism for the allocation path: one process per GPU.

Every sample is an independent optimisation problem, so the hot path needs
no collective: rank r solves rows [r*B/G, (r+1)*B/G) of every input.  The only exchange in training
is the all-reduce of the *network parameter* gradients (the reference is single-process,
deepdow/experiments.py:349-354; its per-sample gradients are summed by autograd of the
repeat_interleave/broadcasts in deepdow/nn.py:563-575 before they reach a parameter).
"""
import torch
import torch.distributed as dist


def shard_bounds(n_samples, rank, world_size):
    """Contiguous row range of `rank`; sizes differ by at most one, every row is owned exactly once."""
    if not (0 <= rank < world_size):
        raise ValueError(f"rank {rank} outside [0, {world_size})")
    base, rem = divmod(n_samples, world_size)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def shard_batch(tensors, rank, world_size):
    """Slice every tensor of a batch along dim 0 to the rows owned by `rank`."""
    n = tensors[0].shape[0]
    for t in tensors:
        if t.shape[0] != n:
            raise ValueError("all tensors of a batch need the same leading dimension")
    lo, hi = shard_bounds(n, rank, world_size)
    return [t[lo:hi] for t in tensors]


class GradAllReducer:
    """Single flat-bucket all-reduce of parameter gradients (messages are tens of KB: latency bound)."""

    def __init__(self, parameters, average=True):
        self.params = [p for p in parameters if p.requires_grad]
        self.average = average

    def __call__(self):
        if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
            return
        grads = [p.grad if p.grad is not None else torch.zeros_like(p) for p in self.params]
        flat = torch.cat([g.reshape(-1) for g in grads])
        dist.all_reduce(flat, op=dist.ReduceOp.SUM)
        if self.average:
            flat /= dist.get_world_size()
        off = 0
        for p, g in zip(self.params, grads):
            n = g.numel()
            p.grad = flat[off:off + n].view_as(g).clone()
            off += n


def global_mean_loss(per_sample_loss, n_global):
    """Local contribution to the global mean: sum over the local rows / global sample count, so that
    summing the gradients over ranks reproduces the single-process mean (experiments.py:350-351)."""
    return per_sample_loss.sum() / n_global
