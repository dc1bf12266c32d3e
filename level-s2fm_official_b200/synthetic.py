"""Synthetic scenes for the benchmark / smoke runs (no dataset is available offline).

Cameras sit on a sphere of radius 2.5 x half-extent looking at the origin, pinhole focal = 1.2 W; rays are the
UN-NORMALISED world-space pixel directions the reference's ``utils/camera.py:230-252`` (get_center_and_ray) produces
(camera-space z = 1), pixels drawn without replacement; only rays that hit the scene box are kept.
"""
from __future__ import annotations

import math

import torch


def look_at_cameras(n_cams: int, half: float, gen: torch.Generator):
    """-> rotation cam->world [B,3,3], camera centres [B,3]."""
    d = torch.nn.functional.normalize(torch.randn(n_cams, 3, generator=gen), dim=-1)
    pos = d * (2.5 * half)
    fwd = -d
    up = torch.tensor([0.0, 1.0, 0.0]).expand_as(fwd)
    right = torch.nn.functional.normalize(torch.cross(up, fwd, dim=-1), dim=-1)
    up2 = torch.cross(fwd, right, dim=-1)
    rot = torch.stack([right, up2, fwd], dim=-1)          # columns: camera x, y, z axes in world coordinates
    return rot, pos


def make_rays(n_cams: int, rays_per_cam: int, half: float, H: int, W: int, seed: int = 0):
    """-> center [B,R,3], ray [B,R,3] (fp32, CPU).  Every ray hits [-half, half]^3."""
    gen = torch.Generator().manual_seed(seed)
    rot, pos = look_at_cameras(n_cams, half, gen)
    f = 1.2 * W
    centers, rays = [], []
    for b in range(n_cams):
        keep_c, keep_r, have = [], [], 0
        while have < rays_per_cam:
            pix = torch.randperm(H * W, generator=gen)[: 2 * rays_per_cam]
            u = (pix % W).float() + 0.5
            v = (pix // W).float() + 0.5
            dir_cam = torch.stack([(u - W / 2) / f, (v - H / 2) / f, torch.ones_like(u)], dim=-1)
            r = dir_cam @ rot[b].T
            o = pos[b].expand_as(r)
            inv = 1.0 / r
            lo, hi = (-half - o) * inv, (half - o) * inv
            t1 = torch.minimum(lo, hi).max(dim=-1).values
            t2 = torch.maximum(lo, hi).min(dim=-1).values
            hit = (t1 <= t2) & (t2 > 0)
            keep_c.append(o[hit])
            keep_r.append(r[hit])
            have += int(hit.sum())
        centers.append(torch.cat(keep_c)[:rays_per_cam])
        rays.append(torch.cat(keep_r)[:rays_per_cam])
    return torch.stack(centers).contiguous(), torch.stack(rays).contiguous()


def init_fields(sdf, rad, regime: str = "init", seed: int = 0):
    """Reference initialisation is already applied by the constructors (geometric sphere init for the SDF MLP,
    default nn.Linear + weight-norm for the radiance MLP, table U(-1e-4, 1e-4)).  ``regime='trained'`` swaps in
    N(0, 0.05) hash tables and lets the first layer see them, so the level set is a perturbed sphere."""
    if regime == "init":
        return
    g = torch.Generator().manual_seed(seed)
    with torch.no_grad():
        tab = sdf.embed_fn.embedder_obj.params
        tab.copy_((torch.randn(tab.shape, generator=g) * 0.05).to(tab.device))
        l0 = sdf.SDF_MLP.mlp[0]
        w = l0.weight_v
        w[:, 3:] = (torch.randn(w.shape[0], w.shape[1] - 3, generator=g) * 0.05).to(w.device)
        l0.weight_g.copy_(w.norm(dim=1, keepdim=True))


def render_loss_fused(out, gt):
    """Same value and gradients as render_loss, through the fused loss kernel (ops.RenderLoss)."""
    from . import ops
    return ops.RenderLoss.apply(out["rgb"], gt, out["normals"], 1e3, 1e2)[0]


def render_loss(out, gt):
    """The rendering losses of the reference's refine stage (pipelines/rendering_refine.py:99-121 with the
    log10 weights rgb: 3, eikonal_loss: 2 of options/LevelS2fM.yaml:120-122)."""
    rgb_l1 = (out["rgb"] - gt).abs().mean()
    eik = (out["normals"].norm(dim=-1) - 1).abs().mean()
    return 1e3 * rgb_l1 + 1e2 * eik
