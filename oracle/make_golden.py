"""Generates tests/golden/*.npz from the UNMODIFIED reference python (build container only).

TEST INFRASTRUCTURE -- see oracle/__init__.py.  Runs /root/reference's own models/{SDF,RadF,Renderer}.py through
oracle/ref_shim.py (third-party tcnn / vren ops replaced by oracle/hashgrid.py / oracle/aabb.py) on seeded inputs
and stores inputs, outputs and gradients.  The fixtures pin (a) oracle/port.py and (b) the CUDA path.

    python -m oracle.make_golden            # rewrites tests/golden/c1_render.npz, c2_sampler.npz, c2_sampler_hard.npz, st_dtu.npz
"""
from __future__ import annotations

import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import port, ref_shim  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")


def _rays(n_cams, n_rays, bound, seed):
    g = torch.Generator().manual_seed(seed)
    center = (torch.randn(n_cams, n_rays, 3, generator=g) * 0.1 + torch.tensor([0.0, 0.0, -2.5])) * bound
    ray = torch.randn(n_cams, n_rays, 3, generator=g) * 0.2 + torch.tensor([0.0, 0.0, 1.0])
    return center, ray


def _load(mod, sd):
    mod.load_state_dict({k: v.detach().clone() for k, v in sd.items()})


def c1_render():
    """BASELINE config 1 (reduced ray count): 64 uniform samples, 2-layer-64 SDF MLP, L=4 hash grid, fp32."""
    L, N, R = 4, 64, 96
    opt = ref_shim.make_opt("DTU", **{"SDF.VolSDF.sample_intvs": N})
    hash_cfg = dict(otype="HashGrid", n_levels=L, n_features_per_level=2, log2_hashmap_size=19, base_resolution=16,
                    per_level_scale=1.38)
    sdf, rad, ren = ref_shim.build_models(opt, hash_config=hash_cfg)
    cfg = port.SceneCfg(n_levels=L, sample_intvs=N, iters_max_st=10)
    sdf_sd, rad_sd = port.random_state(cfg, seed=0, table_std=0.05)
    _load(sdf, sdf_sd)
    _load(rad, rad_sd)
    center, ray = _rays(1, R, 1.0, seed=0)
    out = ren.forward(opt, center, ray, sdf, rad)
    g = torch.Generator().manual_seed(1)
    gt = torch.rand(1, R, 3, generator=g)
    # the reference's rendering losses (pipelines/rendering_refine.py:99-121): 10^3 L1(rgb) + 10^2 eikonal
    loss = 1e3 * (out["rgb"] - gt).abs().mean() + 1e2 * (out["normals"].norm(dim=-1) - 1).abs().mean()
    loss.backward()
    d = {"center": center, "ray": ray, "gt": gt, "loss": loss.detach().reshape(1)}
    for k in ("rgb", "sdfs_volume", "normals", "depth_mlp", "normal_mlp"):
        d["out." + k] = out[k].detach()
    gi = torch.Generator().manual_seed(2)
    for nm, mod in (("sdf", sdf), ("rad", rad)):
        for k, p in mod.named_parameters():
            gr = p.grad.detach()
            if gr.numel() > 100000:
                idx = torch.randperm(gr.numel(), generator=gi)[:4096]
                nz = gr.nonzero()[:, 0]
                idx = torch.cat([idx, nz[torch.randperm(nz.numel(), generator=gi)[:4096]]])
                d[f"gradidx.{nm}.{k}"] = idx
                d[f"gradval.{nm}.{k}"] = gr[idx]
                d[f"gradnorm.{nm}.{k}"] = gr.double().norm().reshape(1)
            else:
                d[f"grad.{nm}.{k}"] = gr
    np.savez_compressed(os.path.join(OUT, "c1_render.npz"), **{k: v.numpy() for k, v in d.items()})
    print("c1_render: loss", float(loss))


def st_dtu():
    """sphere_tracing + get_surface_pts of the reference SDF on a near-sphere field (geometric init + hash perturbation)."""
    L, M = 16, 128
    opt = ref_shim.make_opt("DTU")
    sdf, rad, ren = ref_shim.build_models(opt)
    cfg = port.SceneCfg(n_levels=L, iters_max_st=10)
    sdf_sd, _ = port.random_state(cfg, seed=4, table_std=0.02, generic_weights=False, hash_weight_std=0.05)
    _load(sdf, sdf_sd)
    center, ray = _rays(2, M, 1.0, seed=7)
    d_pred, sdf_last, _, finish = sdf.sphere_tracing(center, ray, sdf)
    d_pred.sum().backward()
    pts = center.reshape(-1, 3) * 0.2
    surf, nv = sdf.get_surface_pts(pts.clone())
    d = {"center": center, "ray": ray, "d_pred": d_pred.detach(), "sdf_last": sdf_last.detach(),
         "finish_mask": finish.to(torch.uint8), "pts": pts, "surf": surf.detach(), "nv": nv.detach()}
    for k, p in sdf.named_parameters():
        if p.grad is not None and p.numel() < 100000:
            d["grad." + k] = p.grad.detach()
    np.savez_compressed(os.path.join(OUT, "st_dtu.npz"), **{k: v.numpy() for k, v in d.items()})
    print("st_dtu: finished", int(finish.sum()), "of", finish.numel())


def c2_sampler():
    """BASELINE config 2's sampler: the reference's error-bounded volsdf_sampling (with the SURVEY a5 fixes)."""
    L, R = 16, 64
    opt = ref_shim.make_opt("DTU", **{"SDF.VolSDF.sample_intvs": 64, "SDF.VolSDF.final_sample_intvs": 64,
                                      "SDF.VolSDF.volsdf_sampling": True, "SDF.arch.layers": [None, 64, 64, 64, 16]})
    ref_shim.apply_c2_fixes(opt)
    sdf, rad, ren = ref_shim.build_models(opt)
    cfg = port.SceneCfg(n_levels=L, sample_intvs=64, final_sample_intvs=64, volsdf_sampling=True,
                        sdf_layers=(None, 64, 64, 64, 16))
    sdf_sd, rad_sd = port.random_state(cfg, seed=6, table_std=0.02, generic_weights=False, hash_weight_std=0.05)
    _load(sdf, sdf_sd)
    _load(rad, rad_sd)
    center, ray = _rays(1, R, 1.0, seed=9)
    with torch.no_grad():
        t, beta_plus, iters = ren.volsdf_sampling(opt, center, ray, SDF_Field=sdf)
    out = ren.forward(opt, center, ray, sdf, rad)
    d = {"center": center, "ray": ray, "t": t, "beta_plus": beta_plus, "iters": iters,
         "out.rgb": out["rgb"].detach(), "out.depth_mlp": out["depth_mlp"].detach()}
    np.savez_compressed(os.path.join(OUT, "c2_sampler.npz"), **{k: v.numpy() for k, v in d.items()})
    print("c2_sampler: t", tuple(t.shape), "iters", iters.unique().tolist())


def c2_sampler_hard():
    """The reference's error-bounded sampler where it struggles: a tight eps, few samples, a rough field -- several up-sampling
    rounds, the bisection on beta+ (models/Renderer.py:281-291) and rays that NEVER converge (iters = -1, Renderer.py:309-321),
    plus rays that miss the box.  Pins oracle/port.volsdf_sampling's bisection / give-up path against the reference itself."""
    L, R, N = 16, 48, 16
    opt = ref_shim.make_opt("DTU", **{"SDF.VolSDF.sample_intvs": N, "SDF.VolSDF.final_sample_intvs": 24, "SDF.VolSDF.volsdf_sampling": True,
                                      "SDF.VolSDF.eps": 0.002, "SDF.VolSDF.max_upsample_iter": 4})
    ref_shim.apply_c2_fixes(opt)
    sdf, rad, ren = ref_shim.build_models(opt)
    cfg = port.SceneCfg(n_levels=L, sample_intvs=N, final_sample_intvs=24, volsdf_sampling=True, eps=0.002, max_upsample_iter=4)
    sdf_sd, _ = port.random_state(cfg, seed=8, table_std=0.05, generic_weights=False, hash_weight_std=0.1)
    _load(sdf, sdf_sd)
    center, ray = _rays(1, R, 1.0, seed=21)
    center[0, :3] += 10                       # rays that miss the box
    with torch.no_grad():
        t, beta_plus, iters = ren.volsdf_sampling(opt, center, ray, SDF_Field=sdf)
    d = {"center": center, "ray": ray, "t": t, "beta_plus": beta_plus, "iters": iters}
    np.savez_compressed(os.path.join(OUT, "c2_sampler_hard.npz"), **{k: v.numpy() for k, v in d.items()})
    print("c2_sampler_hard: t", tuple(t.shape), "iters", {int(k): int((iters == k).sum()) for k in iters.unique()})


if __name__ == "__main__":
    os.makedirs(OUT, exist_ok=True)
    which = sys.argv[1:] or ["c1_render", "st_dtu", "c2_sampler", "c2_sampler_hard"]
    for w in which:
        globals()[w]()
