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

    def __init__(self, params: Iterable[torch.nn.Parameter], extra: int = 8, symmetric: bool = False, direct: bool = True):
        """direct: the fused backward kernels scatter the hash-table gradients straight into the bucket (``ops.grad_sink``) instead
        of into a zero-filled temporary that autograd then adds to ``.grad`` -- two passes over the 49 MB table less per backward.
        Valid for ``loss.backward()`` training loops (the bucket IS ``.grad``); leave it off when calling ``torch.autograd.grad``."""
        self.direct = direct
        self.params: List[torch.nn.Parameter] = [p for p in params if p.requires_grad]
        # every parameter starts on a 16-byte boundary of the bucket: the fused backward scatters into these views with 8-byte vector
        # atomics (red.global.add.v2.f32), and the multimem reduction moves whole 16-byte words
        self.offsets, total = [], 0
        for p in self.params:
            self.offsets.append(total)
            total += (p.numel() + 3) // 4 * 4
        dev = self.params[0].device
        extra += (-(total + extra)) % 4
        self.flat = None
        self.group_name = None      # set when the bucket lives in symmetric memory and the NVSwitch can reduce it (NVLS multicast)
        if (symmetric and dev.type == "cuda" and dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1
                and dist.get_backend() == "nccl"):
            # One process per GPU on one NVSwitch box: put the bucket in symmetric memory and let the switch do the sum
            # (multimem.ld_reduce / multimem.st: every GPU reads 1/N of the bucket reduced in the switch and broadcasts it) --
            # measured on 8 B200s for this 48.9 MB bucket: 0.13 ms vs 0.22 ms for ncclAllReduce (tools/allreduce_probe.py).
            # OPT-IN for now: inside the full step at 2 GPUs the symmetric-memory bucket was slower end to end (3.76 vs 3.53 ms
            # per step) and the GPU budget of round 1 ran out before the 8-GPU step could be measured with it.
            try:
                import torch.distributed._symmetric_memory as symm
                flat = symm.empty(total + extra, dtype=torch.float32, device=dev)
                hdl = symm.rendezvous(flat, dist.group.WORLD.group_name)
                if getattr(hdl, "multicast_ptr", 0):
                    # self-test before trusting it: every rank contributes rank + 1, the sum is known
                    world, gname = dist.get_world_size(), dist.group.WORLD.group_name
                    flat.fill_(float(dist.get_rank() + 1))
                    torch.ops.symm_mem.multimem_all_reduce_(flat, "sum", gname)
                    ok = torch.tensor([float(bool((flat == world * (world + 1) / 2).all()))], device=dev)
                    dist.all_reduce(ok, op=dist.ReduceOp.MIN)
                    if bool(ok.item()):
                        flat.zero_()
                        self.flat, self.group_name = flat, gname
            except Exception:
                self.flat, self.group_name = None, None
        if self.flat is None:
            self.flat = torch.zeros(total + extra, dtype=torch.float32, device=dev)
        for p, off in zip(self.params, self.offsets):
            p.grad = self.flat[off:off + p.numel()].view_as(p)
            if direct:
                p._ls2fm_grad_sink = p.grad
        self.extra = self.flat[total:]

    def zero(self):
        self.flat.zero_()

    def rebind(self):
        """optimizers / zero_grad(set_to_none=True) may drop the views; call before backward."""
        for p, off in zip(self.params, self.offsets):
            if p.grad is None or p.grad.data_ptr() != self.flat[off:off + p.numel()].data_ptr():
                p.grad = self.flat[off:off + p.numel()].view_as(p)
            if self.direct:
                p._ls2fm_grad_sink = p.grad

    def allreduce(self, async_op: bool = False):
        if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
            if self.group_name is not None and not async_op:
                torch.ops.symm_mem.multimem_all_reduce_(self.flat, "sum", self.group_name)
                return None
            return dist.all_reduce(self.flat, op=dist.ReduceOp.SUM, async_op=async_op)
        return None

    @property
    def collective(self) -> str:
        return "NVLS multimem all-reduce (symmetric memory)" if self.group_name is not None else "ncclAllReduce"
