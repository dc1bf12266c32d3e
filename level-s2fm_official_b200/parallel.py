"""Ray-parallel data parallelism: one process per GPU, full replica of both fields, each rank renders
its contiguous slice of the iteration's rays, ONE all-reduce (sum) of a flat fp32 gradient bucket per iteration.

The reference has no multi-GPU support (utils/options.py:110 asserts a single int GPU); this is new work on top of
``torch.distributed`` (NCCL over NVLink on the B200 box, gloo in the CPU tests)."""
from __future__ import annotations

from typing import Iterable, List

import torch
import torch.distributed as dist


def shard_rays(center: torch.Tensor, ray: torch.Tensor, rank: int, world: int):
    """Contiguous split of the R rays of every camera: [B,R,3] -> [B,R/world,3] (R must divide evenly)."""
    R = center.shape[1]
    if R % world:
        raise ValueError(f"{R} rays per camera do not split evenly over {world} ranks")
    n = R // world
    return center[:, rank * n:(rank + 1) * n].contiguous(), ray[:, rank * n:(rank + 1) * n].contiguous()


class GradBucket:
    """All parameter gradients as views into one flat fp32 buffer, laid out so the single all-reduce that follows the
    fused backward needs no packing step.  ``extra`` floats at the tail carry loss sums / counts."""

    def __init__(self, params: Iterable[torch.nn.Parameter], extra: int = 8):
        self.params: List[torch.nn.Parameter] = [p for p in params if p.requires_grad]
        total = sum(p.numel() for p in self.params)
        dev = self.params[0].device
        self.flat = torch.zeros(total + extra, dtype=torch.float32, device=dev)
        off = 0
        for p in self.params:
            p.grad = self.flat[off:off + p.numel()].view_as(p)
            off += p.numel()
        self.extra = self.flat[off:]

    def zero(self):
        self.flat.zero_()

    def rebind(self):
        """optimizers / zero_grad(set_to_none=True) may drop the views; call before backward."""
        off = 0
        for p in self.params:
            if p.grad is None or p.grad.data_ptr() != self.flat[off:off + p.numel()].data_ptr():
                p.grad = self.flat[off:off + p.numel()].view_as(p)
            off += p.numel()

    def allreduce(self, async_op: bool = False):
        if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
            return dist.all_reduce(self.flat, op=dist.ReduceOp.SUM, async_op=async_op)
        return None
