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

    def __init__(self, shapes: dict, device):
        self.names = list(shapes)
        sizes = [int(torch.Size(shapes[n]).numel()) for n in self.names]
        self.flat = torch.zeros(sum(sizes), dtype=torch.float32, device=device)
        self.views, o = {}, 0
        for n, sz in zip(self.names, sizes):
            self.views[n] = self.flat[o:o + sz].view(shapes[n])
            o += sz

    def pack(self, grads: dict):
        for n, v in self.views.items():
            v.copy_(grads[n])          # strided gradient views are fine

    def all_reduce_mean(self):
        if dist.get_backend() == "gloo":   # gloo has no AVG
            dist.all_reduce(self.flat, op=dist.ReduceOp.SUM)
            self.flat.div_(dist.get_world_size())
        else:
            dist.all_reduce(self.flat, op=dist.ReduceOp.AVG)
        return self.views
