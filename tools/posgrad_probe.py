"""Cost of the position-gradient route: one C2-sized render iteration (4096 rays x 128 uniform samples) with and without a
grad-requiring pose (rays from ops.GenerateRays).   python tools/posgrad_probe.py"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from levels2fm_b200 import ops, synthetic  # noqa: E402
from levels2fm_b200.models.RadF import RadF  # noqa: E402
from levels2fm_b200.models.Renderer import Renderer  # noqa: E402
from levels2fm_b200.models.SDF import SDF  # noqa: E402

dev = "cuda:0"
opt = bench.workload_opt("uniform128", dev)
torch.manual_seed(0)
sdf, rad, ren = SDF(opt).to(dev), RadF(opt).to(dev), Renderer(opt)
center, ray = synthetic.make_rays(1, 4096, 1.0, 1200, 1600, seed=0)
center, ray = center.to(dev), ray.to(dev)
gt = torch.rand(1, 4096, 3, device=dev)


def run(req):
    c = center.clone().requires_grad_(req)
    r = ray.clone().requires_grad_(req)
    for p in list(sdf.parameters()) + list(rad.parameters()):
        p.grad = None
    out = ren.forward(opt, c, r, sdf, rad)
    synthetic.render_loss_fused(out, gt).backward()
    return c.grad, r.grad


for req in (False, True):
    for _ in range(3):
        run(req)
    torch.cuda.synchronize()
    ops.KLOG.reset()
    ops.KLOG.timing = True
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(5):
        run(req)
    b.record()
    torch.cuda.synchronize()
    ops.KLOG.timing = False
    d = {k: sum(v) / 5 for k, v in ops.KLOG.durations_ms().items()}
    print(f"rays require grad = {req}: {a.elapsed_time(b) / 5:.3f} ms / iteration; field_backward {d.get('field_backward', 0):.3f} ms, "
          f"field_forward {d.get('field_forward', 0):.3f} ms")

if os.environ.get("LS2FM_PROFILE"):
    from torch.profiler import ProfilerActivity, profile
    with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
        for _ in range(3):
            run(True)
        torch.cuda.synchronize()
    print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=25, max_name_column_width=60))
