"""Loss curves over 100 Adam steps: product vs oracle (fp32) vs oracle (fp64), same init -- how much of the difference is the
intrinsic sensitivity of Adam's sign-like updates to 1e-7-level gradient noise?   python tools/matched_loss_probe.py"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from levels2fm_b200 import synthetic  # noqa: E402
from oracle import port  # noqa: E402
from tests import common  # noqa: E402

DEV = "cuda"


def run(regime, lr_sdf, lr_col, eb, n_rays=512, steps=100):
    over = {"SDF.VolSDF.volsdf_sampling": True, "SDF.VolSDF.sample_intvs": 32, "SDF.VolSDF.final_sample_intvs": 32} if eb else {}
    opt = common.make_opt("DTU", DEV, 16, (None, 64, 64, 64, 16), 64, False, **over)
    cfg = common.cfg_of(opt, 16)
    if regime == "init":
        sdf_sd, rad_sd = port.random_state(cfg, seed=1, table_std=1e-4, generic_weights=False)
    else:
        sdf_sd, rad_sd = port.random_state(cfg, seed=1, table_std=0.02, generic_weights=False, hash_weight_std=0.05)
    sdf, rad, ren = common.build_models(opt)
    sdf.load_state_dict(sdf_sd)
    rad.load_state_dict(rad_sd)
    center, ray = synthetic.make_rays(1, n_rays, 1.0, 1200, 1600, seed=3)
    gt = torch.rand(1, n_rays, 3, generator=torch.Generator().manual_seed(4))
    center, ray, gt = center.to(DEV), ray.to(DEV), gt.to(DEV)
    curves = {}
    # ours
    o = torch.optim.Adam([{"params": sdf.parameters(), "lr": lr_sdf}, {"params": rad.parameters(), "lr": lr_col}])
    c = []
    for it in range(steps + 1):
        o.zero_grad(set_to_none=True)
        l = synthetic.render_loss(ren.forward(opt, center, ray, sdf, rad), gt)
        l.backward()
        c.append(float(l.detach()))
        o.step()
    curves["ours"] = c
    for name, dt in (("port32", torch.float32), ("port64", torch.float64)):
        s1 = {k: v.detach().clone().to(DEV).to(dt).requires_grad_(True) for k, v in sdf_sd.items()}
        r1 = {k: v.detach().clone().to(DEV).to(dt).requires_grad_(True) for k, v in rad_sd.items()}
        o = torch.optim.Adam([{"params": list(s1.values()), "lr": lr_sdf}, {"params": list(r1.values()), "lr": lr_col}])
        c = []
        for it in range(steps + 1):
            o.zero_grad(set_to_none=True)
            l = synthetic.render_loss(port.render_forward(center.to(dt), ray.to(dt), s1, r1, cfg), gt.to(dt))
            l.backward()
            c.append(float(l.detach()))
            o.step()
        curves[name] = c
    t = {k: torch.tensor(v, dtype=torch.float64) for k, v in curves.items()}
    rel = lambda a, b: ((t[a] - t[b]).abs() / t[b].abs())
    print(f"regime={regime} lr_sdf={lr_sdf} lr_color={lr_col} eb={eb}: loss {t['port64'][0]:.4f} -> {t['port64'][-1]:.4f}")
    for a, b in (("ours", "port32"), ("ours", "port64"), ("port32", "port64")):
        r = rel(a, b)
        print(f"   {a:7s} vs {b:7s}: step0 {r[0]:.2e}  max over curve {r.max():.2e} (step {int(r.argmax())})  step100 {r[-1]:.2e}  median {r.median():.2e}")


for regime, lr_sdf, lr_col in (("init", 1e-4, 1e-3), ("init", 1e-3, 1e-3), ("trained", 1e-4, 1e-3), ("trained", 1e-3, 1e-3)):
    for eb in (False, True):
        run(regime, lr_sdf, lr_col, eb)
