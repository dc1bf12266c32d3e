"""GPU diagnostic for the experimental warp-specialised values-only kernel: parity against the default kernel and device time on
the sampler's first launch shape (4096 rays x 64 uniform samples, C2 SDF network).
    python tools/ws_probe.py                              (4 gather warps)
    LS2FM_WS_GATHER_WARPS=8 python tools/ws_probe.py      (8 gather warps: two threads per sample)
    LS2FM_WS_DEPTH=4 python tools/ws_probe.py             (4 levels = 32 table reads in flight per gather thread)
Round 1, B200, 4 gather warps: bit-identical, 185 us vs 165 us for the default kernel."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import port
from tests import common
from levels2fm_b200 import _C, ops

lib = _C.get()
DEV = "cuda"
opt = common.make_opt("DTU", DEV, 16, (None, 64, 64, 64, 16), 64)
cfg = common.cfg_of(opt, 16)
sdf_sd, _ = port.random_state(cfg, seed=6, table_std=0.05)
sdf, _, _ = common.build_models(opt)
sdf.load_state_dict(sdf_sd)
spec, table = sdf.field_spec(), sdf.table().detach()
theta = sdf.SDF_MLP.theta().detach().contiguous()
center, ray = common.make_rays(1, 4096, 1.0, seed=5)
c2, r2 = center.reshape(-1, 3).contiguous().to(DEV), ray.reshape(-1, 3).contiguous().to(DEV)
t, _ = ops.sample_uniform_raw(lib, c2, r2, 64, cfg.bound_min, cfg.bound_max)
pts = ops._points(lib, None, c2, r2, t)
image = ops.field_prepare_raw(lib, spec, table, theta, None)


def run(ws):
    ops.FORWARD_WS = ws
    try:
        return ops.field_forward_raw(lib, spec, table, theta, pts, None, want_y=True, image=image)
    finally:
        ops.FORWARD_WS = False


ref = run(False)
out = run(True)
torch.cuda.synchronize()
print("parity: y rel err %.2e, sdf rel err %.2e" % (common.rel_err(out[0].cpu(), ref[0].cpu()), common.rel_err(out[1].cpu(), ref[1].cpu())), flush=True)
for ws in (False, True, False, True):
    for _ in range(3):
        run(ws)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda._sleep(int(4e6))
    a.record()
    for _ in range(20):
        run(ws)
    b.record()
    torch.cuda.synchronize()
    print(("ws     " if ws else "default"), "%.1f us per launch (%d samples)" % (a.elapsed_time(b) / 20 * 1e3, int(pts.n)), flush=True)
