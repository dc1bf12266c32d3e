"""8-GPU diagnostic: device time of the 48.9 MB gradient-bucket all-reduce under the NCCL configuration in the environment,
and of torch's symmetric-memory (NVLink SHARP / multimem) all-reduce when available.  torchrun --nproc-per-node N tools/allreduce_probe.py"""
import os, sys, time
import torch
import torch.distributed as dist

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device(f"cuda:{local}")
os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
dist.init_process_group("nccl", device_id=dev)
n = 12_230_000
x = torch.randn(n, device=dev)


def timeit(fn, iters=20):
    for _ in range(5):
        fn()
    torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters):
        fn()
    b.record()
    torch.cuda.synchronize()
    t = torch.tensor([a.elapsed_time(b) / iters], device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


tag = os.environ.get("PROBE_TAG", "default")
ms = timeit(lambda: dist.all_reduce(x))
if rank == 0:
    print(f"[{tag}] nccl all_reduce {n * 4 / 1e6:.1f} MB x{world}: {ms:.3f} ms  busbw {2 * (world - 1) / world * n * 4 / ms / 1e6:.0f} GB/s", flush=True)
if os.environ.get("PROBE_SYMM", "0") == "1":
    try:
        import torch.distributed._symmetric_memory as symm
        buf = symm.empty(n, device=dev)
        hdl = symm.rendezvous(buf, dist.group.WORLD.group_name)
        buf.copy_(x)
        for name in ("multimem_all_reduce_", "two_shot_all_reduce_", "one_shot_all_reduce"):
            op = getattr(torch.ops.symm_mem, name, None)
            if op is None:
                continue
            try:
                ms = timeit(lambda: op(buf, "sum", dist.group.WORLD.group_name))
                if rank == 0:
                    print(f"[{tag}] symm_mem.{name}: {ms:.3f} ms", flush=True)
            except Exception as e:
                if rank == 0:
                    print(f"[{tag}] symm_mem.{name} failed: {type(e).__name__}: {str(e)[:150]}", flush=True)
    except Exception as e:
        if rank == 0:
            print(f"[{tag}] symmetric memory unavailable: {type(e).__name__}: {str(e)[:200]}", flush=True)
dist.destroy_process_group()
