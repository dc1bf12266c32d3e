"""What the hash-grid gather / scatter of one C2 step costs when NOTHING else runs: the unfused tcnn-style encode kernels
(one thread per (point, level), 64 warps per SM) on the very sample set of the benchmark step.  This is the practical floor of
the memory side of the fused field kernels."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from levels2fm_b200 import _C, ops, synthetic  # noqa: E402
from levels2fm_b200.models.RadF import RadF  # noqa: E402
from levels2fm_b200.models.Renderer import Renderer  # noqa: E402
from levels2fm_b200.models.SDF import SDF  # noqa: E402

dev = "cuda:0"
lib = _C.get()
opt = bench.workload_opt("c2", dev)
torch.manual_seed(0)
sdf, rad, ren = SDF(opt).to(dev), RadF(opt).to(dev), Renderer(opt)
synthetic.init_fields(sdf, rad, "init")
center, ray = synthetic.make_rays(1, 4096, 1.0, 1200, 1600, seed=0)
center, ray = center.to(dev), ray.to(dev)
with torch.no_grad():
    t, _, _ = ren.volsdf_sampling(opt, center, ray, sdf)
x = center[:, :, None, :] + ray[:, :, None, :] * t[..., None]
u = ((x.reshape(-1, 3) + 1.0) / 2.0).contiguous()
grid = sdf.embed_fn.embedder_obj.grid
table = sdf.table().detach()
flush = torch.empty(256 * 1024 * 1024 // 4, device=dev)


def timed(fn, reps=10, cold=True):
    ts = []
    for _ in range(reps):
        if cold:
            flush.fill_(1.0)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return sorted(ts)[len(ts) // 2]


g_enc = torch.randn(u.shape[0], grid.n_output_dims, device=dev)
d_table = torch.zeros_like(table)
S = u.shape[0]
for name, pts in (("ray-ordered C2 samples", u), ("same points, shuffled", u[torch.randperm(S, device=dev)].contiguous())):
    for cold in (True, False):
        tg = timed(lambda: ops.grid_encode_raw(lib, grid, table, pts), cold=cold)
        ts = timed(lambda: ops.grid_encode_backward_raw(lib, grid, table, pts, g_enc, d_table), cold=cold)
        print(f"{name:26s} L2 {'flushed' if cold else 'warm   '}: gather {tg:.3f} ms ({S * 1024 / tg / 1e6:.0f} GB/s alg)  "
              f"scatter {ts:.3f} ms ({S * 1024 / ts / 1e6:.0f} GB/s alg)   [S = {S}]")
