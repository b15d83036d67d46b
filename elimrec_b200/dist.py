"""Multi-GPU plumbing (one process per GPU, torch.distributed; NCCL on GPUs, gloo in the CPU tests).

SURVEY.md section 8e.  Two pieces are used today:
  * ``shard_range``   - contiguous user shards for the full-ranking evaluator (embarrassingly parallel;
                        the only collective is one all-reduce of the [n_metrics*K] metric sums);
  * ``GradBucket``    - data-parallel training replicas: every gradient of a step is packed into ONE flat
                        buffer and averaged with ONE all-reduce (30 MB at Tiktok shape), then Adam runs
                        identically on every rank.
"""
from __future__ import annotations

import torch
import torch.distributed as dist


def shard_range(n: int, rank: int, world: int):
    """[lo, hi) of the rank-th of ``world`` contiguous shards of ``n`` items (sizes differ by at most 1 block)."""
    if world <= 1:
        return 0, n
    per = (n + world - 1) // world
    return min(n, rank * per), min(n, (rank + 1) * per)


def is_dist():
    return dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1


class GradBucket:
    """Flat fp32 buffer holding every gradient tensor back to back; ``views[name]`` has the parameter's shape."""

    def __init__(self, shapes: dict, device, tail_flat=None, tail_views=None):
        """Part 0 (``head``): the tensors of ``shapes``, packed back to back into an owned flat buffer.  Part 1 (``tail``,
        optional): an EXISTING flat buffer whose views ``tail_views`` the backward pass writes directly (no packing).
        The parts are all-reduced separately (``all_reduce_mean_part``): the embedding-table gradients (part 0) are final
        before the weight gradients (part 1) are even started, so their all-reduce goes on the wire first and overlaps
        the rest of the backward pass."""
        self.head_names = list(shapes)
        sizes = [int(torch.Size(shapes[n]).numel()) for n in self.head_names]
        self.flat = torch.zeros(sum(sizes), dtype=torch.float32, device=device)
        self.views, o = {}, 0
        for n, sz in zip(self.head_names, sizes):
            self.views[n] = self.flat[o:o + sz].view(shapes[n])
            o += sz
        self.tail_flat = tail_flat
        self.tail_names = list(tail_views) if tail_views else []
        self.views.update(tail_views or {})
        self.names = self.head_names + self.tail_names

    def pack(self, grads: dict, names=None):
        for n in (self.names if names is None else names):
            if grads[n].data_ptr() != self.views[n].data_ptr():
                self.views[n].copy_(grads[n])          # strided gradient views are fine

    def _mean(self, t, async_op=False):
        if t is None or t.numel() == 0:
            return None
        if dist.get_backend() == "gloo":   # gloo has no AVG (CPU tests): synchronous
            dist.all_reduce(t, op=dist.ReduceOp.SUM)
            t.div_(dist.get_world_size())
            return None
        return dist.all_reduce(t, op=dist.ReduceOp.AVG, async_op=async_op)

    def all_reduce_mean_part(self, part: int, async_op: bool = False):
        """mean over ranks of part 0 (packed head) or part 1 (in-place tail); returns the work handle when async"""
        return self._mean(self.flat if part == 0 else self.tail_flat, async_op)

    def all_reduce_mean(self):
        self._mean(self.flat)
        self._mean(self.tail_flat)
        return self.views
