"""Data-parallel plumbing of the path: the batch is sharded by rank, nothing on the data path
is exchanged, and the only collective is ONE all-reduce of the small CBAM weight gradients
(reference: DDP's bucketed all-reduce of every parameter, ultralytics/engine/trainer.py:241-252,367,
loss pre-multiplied by world_size :481-482 -- i.e. gradients are averaged over ranks).

Inside a reference DDP model nothing here is needed: gradients stay in `.grad` and DDP reduces
them.  `FlatGradReducer` is for running the blocks outside DDP (bench.py, module-only training):
all levels' gradients live in one flat fp32 buffer -> one NCCL call per step (n: 47 KB, x: 0.67 MB).
"""
from __future__ import annotations

from typing import Iterable, List

import torch
import torch.distributed as dist


def shard_range(n: int, rank: int, world: int) -> range:
    """Contiguous batch slice of rank `rank` (same rule as batch // world_size + DistributedSampler
    on an evenly divisible batch, ultralytics/engine/trainer.py:379)."""
    per, rem = divmod(n, world)
    lo = rank * per + min(rank, rem)
    return range(lo, lo + per + (1 if rank < rem else 0))


class FlatGradReducer:
    """Makes the `.grad` of every parameter a view into one flat buffer and all-reduces it once."""

    def __init__(self, params: Iterable[torch.nn.Parameter], average: bool = True):
        self.params: List[torch.nn.Parameter] = [p for p in params if p.requires_grad]
        if not self.params:
            raise ValueError("no trainable parameters")
        dev = self.params[0].device
        self.average = average
        self.numel = sum(p.numel() for p in self.params)
        self.flat = torch.zeros(self.numel, dtype=torch.float32, device=dev)
        self._bind(copy=False)

    def _bind(self, copy: bool) -> None:
        """Make every .grad a view into the flat buffer.  `optimizer.zero_grad()` / `Module.zero_grad()` default to
        set_to_none=True and the next backward then allocates fresh .grad tensors OUTSIDE the buffer: those are copied in
        and re-bound here, so the collective never reduces a stale buffer."""
        o = 0
        for p in self.params:
            n = p.numel()
            view = self.flat[o:o + n].view_as(p)
            g = p.grad
            if g is None:
                if copy:
                    view.zero_()  # no gradient this step: contributes zero
                p.grad = view
            elif g.data_ptr() != view.data_ptr() or g.dtype != torch.float32:
                view.copy_(g)
                p.grad = view
            o += n

    def zero(self) -> None:
        self.flat.zero_()

    def all_reduce(self, async_op: bool = False):
        self._bind(copy=True)
        if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
            return None
        work = dist.all_reduce(self.flat, op=dist.ReduceOp.SUM, async_op=async_op)
        if async_op:
            return work
        if self.average:
            self.flat.div_(dist.get_world_size())
        return None

    def finish(self, work) -> None:
        if work is not None:
            work.wait()
            if self.average:
                self.flat.div_(dist.get_world_size())
