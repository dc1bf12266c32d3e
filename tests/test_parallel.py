"""Host logic of the ray-parallel multi-GPU path (levels2fm_b200/parallel.py) on CPU: gloo, world size 2."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import port


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port_no, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port_no))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from levels2fm_b200 import parallel
    torch.manual_seed(0)                       # identical replicas
    cfg = port.SceneCfg(n_levels=2, sample_intvs=8, log2_hashmap_size=10)
    sdf_sd, rad_sd = port.random_state(cfg, seed=1, table_std=0.2)
    params = [torch.nn.Parameter(v) for v in list(sdf_sd.values()) + list(rad_sd.values())]
    keys_s, keys_r = list(sdf_sd), list(rad_sd)
    bucket = parallel.GradBucket(params, extra=4)
    g = torch.Generator().manual_seed(5)
    center = torch.randn(1, 8, 3, generator=g) * 0.1 + torch.tensor([0.0, 0.0, -2.5])
    ray = torch.randn(1, 8, 3, generator=g) * 0.2 + torch.tensor([0.0, 0.0, 1.0])
    gt = torch.rand(1, 8, 3, generator=g)

    def loss_sum(c, r, t):
        sd_s = dict(zip(keys_s, params[:len(keys_s)]))
        sd_r = dict(zip(keys_r, params[len(keys_s):]))
        out = port.render_forward(c, r, sd_s, sd_r, cfg)
        # unnormalised sums: the mean over the GLOBAL ray count is applied by scaling with known denominators
        return (out["rgb"] - t).abs().sum() / (8 * 3) + (out["normals"].norm(dim=-1) - 1).abs().sum() / (8 * 8)

    c_loc, r_loc = parallel.shard_rays(center, ray, rank, world)
    gt_loc = gt[:, rank * 4:(rank + 1) * 4]
    bucket.zero()
    loss = loss_sum(c_loc, r_loc, gt_loc)
    loss.backward()
    assert all(p.grad.data_ptr() == bucket.flat[o:o + 1].data_ptr() for p, o in zip(bucket.params, bucket.offsets))
    assert all(o % 4 == 0 for o in bucket.offsets)          # 16-byte aligned views: the fused backward scatters with vector atomics
    bucket.extra[0] = loss.detach()
    bucket.allreduce()
    reduced = bucket.flat.clone()
    # single-process reference on all rays
    for p in params:
        p.grad = None
    full = loss_sum(center, ray, gt)
    full.backward()
    ref = torch.cat([p.grad.reshape(-1) for p in params])
    got = torch.cat([reduced[o:o + p.numel()] for p, o in zip(bucket.params, bucket.offsets)])
    err = (got - ref).abs().max().item() / ref.abs().max().item()
    n_flat = bucket.flat.numel() - bucket.extra.numel()
    q.put((rank, err, abs(reduced[n_flat].item() - full.item()) / abs(full.item())))
    dist.barrier()
    dist.destroy_process_group()


def test_gradient_bucket_allreduce_equals_full_batch():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port_no = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port_no, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=240) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, err, lerr in res:
        assert err < 1e-5 and lerr < 1e-5, (rank, err, lerr)


def test_shard_rays_partitions_exactly():
    from levels2fm_b200 import parallel
    c = torch.arange(2 * 12 * 3, dtype=torch.float32).view(2, 12, 3)
    parts = [parallel.shard_rays(c, c, r, 4)[0] for r in range(4)]
    assert torch.equal(torch.cat(parts, dim=1), c)
    try:
        parallel.shard_rays(c, c, 0, 5)
        assert False
    except ValueError:
        pass


def test_bucket_views_survive_zero_and_accumulate():
    from levels2fm_b200 import parallel
    ps = [torch.nn.Parameter(torch.randn(3, 4)), torch.nn.Parameter(torch.randn(5))]
    b = parallel.GradBucket(ps, extra=2)
    (ps[0].sum() * 2 + ps[1].sum() * 3).backward()
    assert b.offsets == [0, 12]
    assert torch.equal(b.flat[:12], torch.full((12,), 2.0)) and torch.equal(b.flat[12:17], torch.full((5,), 3.0))
    (ps[0].sum()).backward()                       # accumulates in place into the same storage
    assert torch.equal(b.flat[:12], torch.full((12,), 3.0))
    b.zero()
    assert b.flat.abs().sum() == 0 and ps[0].grad.data_ptr() == b.flat.data_ptr()
    ps[0].grad = None
    b.rebind()
    assert ps[0].grad.data_ptr() == b.flat.data_ptr()
