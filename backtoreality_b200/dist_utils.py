"""Host-side data-parallel plumbing: scene sharding and the single flat-gradient all-reduce.

Mirrors what the reference gets from DistributedSampler + DistributedDataParallel
(/root/reference/detection/GroupFree3D/train_GF_FSB.py:172-180, 86-96): scenes are independent, so
each rank takes a disjoint slice of the global batch, the forward needs no communication, BatchNorm
statistics stay per replica, and ONE sum all-reduce of the fp32 gradients per step keeps the
replicas identical.  Here the gradients live in one flat buffer so the collective is a single NCCL
call over NVLink (gloo on CPU in the tests).
"""
import torch
import torch.distributed as dist


def shard_scene_indices(global_batch, rank, world, step=0, first_seed=1000):
    """Scene seeds of this rank for `step`: contiguous, disjoint, covering the global batch."""
    if global_batch % world:
        raise ValueError("global batch %d is not divisible by world size %d" % (global_batch, world))
    per = global_batch // world
    base = first_seed + step * global_batch + rank * per
    return list(range(base, base + per))


class FlatGradBucket:
    """One contiguous fp32 buffer with a view per parameter.

    as_views=True (default): every p.grad IS its view, autograd accumulates into the flat buffer
    (one small add kernel per parameter per step) and `zero()` resets it.
    as_views=False: p.grad is left alone (None between steps, so autograd ASSIGNS fresh gradients
    without any add kernel); `reduce_from(grads)` packs them with one multi-tensor copy,
    all-reduces, and points every p.grad at its averaged view for the optimizer.
    """

    def __init__(self, params, as_views=True):
        self.params = [p for p in params]
        n = sum(p.numel() for p in self.params)
        dev = self.params[0].device if self.params else torch.device("cpu")
        self.flat = torch.zeros(n, dtype=torch.float32, device=dev)
        self.views = []
        off = 0
        for p in self.params:
            self.views.append(self.flat[off:off + p.numel()].view_as(p))
            off += p.numel()
        if as_views:
            for p, v in zip(self.params, self.views):
                p.grad = v

    def reduce_from(self, grads):
        """grads: one tensor per parameter (e.g. the p.grad list right after backward)."""
        torch._foreach_copy_(self.views, list(grads))
        self.allreduce_mean()
        for p, v in zip(self.params, self.views):
            p.grad = v

    def allreduce_mean(self):
        """Sum over ranks, divide by world size (DDP semantics).  No-op without a process group."""
        if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
            dist.all_reduce(self.flat)
            self.flat.mul_(1.0 / dist.get_world_size())

    def zero(self):
        self.flat.zero_()
