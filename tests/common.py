"""Shared builders for the parity tests: seeded scenes, our modules loaded with the oracle's parameters."""
from __future__ import annotations

import json
import os
import tempfile

import torch

from oracle import port

_HASH_FILES = {}


def hash_cfg_file(n_levels: int) -> str:
    if n_levels not in _HASH_FILES:
        f = tempfile.NamedTemporaryFile("w", suffix=".json", delete=False)
        json.dump({"encoding": {"otype": "HashGrid", "n_levels": n_levels, "n_features_per_level": 2,
                                "log2_hashmap_size": 19, "base_resolution": 16, "per_level_scale": 1.38}}, f)
        f.close()
        _HASH_FILES[n_levels] = f.name
    return _HASH_FILES[n_levels]


def make_opt(dataset="DTU", device="cpu", n_levels=16, sdf_layers=(None, 64, 16), n_samples=128, dual=False, **extra):
    from levels2fm_b200.config import default_opt
    over = {"SDF.Hash_config.config_file": hash_cfg_file(n_levels), "SDF.arch.layers": list(sdf_layers),
            "SDF.VolSDF.sample_intvs": n_samples, "Ablate_config.dual_field": dual}
    over.update(extra)
    return default_opt(dataset, device=device, **over)


def cfg_of(opt, n_levels) -> port.SceneCfg:
    v = opt.SDF.VolSDF
    return port.SceneCfg(
        bound_min=tuple(float(x) for x in opt.data.bound_min), bound_max=tuple(float(x) for x in opt.data.bound_max),
        inside=bool(opt.data.inside), bgcolor=tuple(float(x) for x in opt.data.bgcolor), scale_mlp=float(opt.SDF.NN_Init.scale_mlp),
        rescale=float(v.rescale), beta_speed=float(v.beta_speed), beta_init=float(v.beta_init), sdf_threshold=float(v.sdf_threshold),
        iters_max_st=int(v.iters_max_st), res=int(opt.Res), sample_intvs=int(v.sample_intvs),
        final_sample_intvs=int(v.final_sample_intvs), volsdf_sampling=bool(v.volsdf_sampling),
        max_upsample_iter=int(v.max_upsample_iter), max_bisection_itr=int(v.max_bisection_itr), eps=float(v.eps),
        n_levels=n_levels, sdf_layers=tuple(opt.SDF.arch.layers), rad_layers=tuple(opt.RadF.arch.layers),
        dual_field=bool(opt.Ablate_config.dual_field))


def make_rays(n_cams, n_rays, bound, seed=3, device="cpu"):
    g = torch.Generator().manual_seed(seed)
    center = (torch.randn(n_cams, n_rays, 3, generator=g) * 0.1 + torch.tensor([0.0, 0.0, -2.5])) * bound
    ray = torch.randn(n_cams, n_rays, 3, generator=g) * 0.2 + torch.tensor([0.0, 0.0, 1.0])
    return center.to(device), ray.to(device)


def build_models(opt):
    from levels2fm_b200.models.RadF import RadF
    from levels2fm_b200.models.Renderer import Renderer
    from levels2fm_b200.models.SDF import SDF
    return SDF(opt).to(opt.device), RadF(opt).to(opt.device), Renderer(opt)


def loss_fn(out, gw):
    """A generic scalar over every output the reference returns + the eikonal term on the per-sample normals."""
    return sum((out[k] * gw[k]).sum() for k in gw) + 3.0 * (out["normals"].norm(dim=-1) - 1).abs().mean()


def render_parity_case(opt, n_levels, n_cams, n_rays, seed=5, table_std=0.2, device="cpu"):
    """Runs oracle and product on the same seeded scene.  Returns dict of (ours, ref) outputs and grads."""
    cfg = cfg_of(opt, n_levels)
    sdf_sd, rad_sd = port.random_state(cfg, seed=seed, table_std=table_std)
    sdf, rad, ren = build_models(opt)
    assert sorted(sdf.state_dict().keys()) == sorted(sdf_sd.keys())
    assert sorted(rad.state_dict().keys()) == sorted(rad_sd.keys())
    sdf.load_state_dict(sdf_sd)
    rad.load_state_dict(rad_sd)
    for sd in (sdf_sd, rad_sd):
        for k in sd:
            sd[k] = sd[k].clone().requires_grad_(True)
    center, ray = make_rays(n_cams, n_rays, float(opt.data.bound_max[0]))
    ref = port.render_forward(center, ray, sdf_sd, rad_sd, cfg)
    out = ren.forward(opt, center.to(device), ray.to(device), sdf, rad)
    g = torch.Generator().manual_seed(11)
    keys = ["rgb", "depth_mlp", "normal_mlp", "sdfs_volume"]
    gw = {k: torch.randn(ref[k].shape, generator=g) for k in keys}
    loss_fn(ref, gw).backward()
    loss_fn(out, {k: v.to(device) for k, v in gw.items()}).backward()
    outs = {k: (out[k].detach().cpu(), ref[k].detach()) for k in ["rgb", "sdfs_volume", "normals", "depth_mlp", "normal_mlp"]}
    grads = {}
    for mod, sd, nm in ((sdf, sdf_sd, "sdf"), (rad, rad_sd, "rad")):
        for k, p in mod.named_parameters():
            grads[f"{nm}.{k}"] = (p.grad.detach().cpu(), sd[k].grad)
    return outs, grads


def rel_err(a, b):
    return ((a - b).abs().max() / b.abs().max().clamp_min(1e-30)).item()


def cosine(a, b):
    return torch.nn.functional.cosine_similarity(a.reshape(-1).double(), b.reshape(-1).double(), dim=0).item()


def cell_ids(x, cfg):
    """Integer cell coordinates of world points x [M,3] at every hash-grid level -> int64 [M, L, 3] (the floor() of SURVEY A.1).
    The field is trilinear INSIDE a cell: its gradient jumps across cell faces, so two evaluations of the 'same' point that
    differ by an ulp can legitimately return different normals iff their cell ids differ at some level."""
    bmin = torch.tensor(cfg.bound_min, dtype=x.dtype)
    bmax = torch.tensor(cfg.bound_max, dtype=x.dtype)
    u = ((x - bmin) / (bmax - bmin)).reshape(-1, 3)
    out = []
    for lvl in cfg.grid().levels:
        p = (u.double() * lvl.scale + 0.5).float()
        out.append(torch.floor(p).to(torch.int64))
    return torch.stack(out, dim=1)


def same_cells(xa, xb, cfg):
    """bool [M]: both point sets fall into the same cell at every level."""
    return (cell_ids(xa, cfg) == cell_ids(xb, cfg)).all(dim=-1).all(dim=-1)
