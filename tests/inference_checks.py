"""Forward-only entry points of SURVEY 8(f) row 4 against the oracle: full-image rendering in slices
(/root/reference/pipelines/Camera.py:275-311) and the marching-cubes SDF volume (/root/reference/utils/util.py:392-430).
Shared by the emulator tests (device "cpu") and the GPU tests (device "cuda")."""
from __future__ import annotations

import numpy as np
import torch

from oracle import port

from . import common


def reference_grid_points(N, volume_size=2.0, bound_max=None, bound_min=None):
    """utils/util.py:392-411 restated with numpy exactly as the reference writes it (np.int -> int): float64 arithmetic, TRUE
    division of the flat index -- the y / x columns are not integer grid indices."""
    s = volume_size
    voxel_grid_origin = [-s / 2., -s / 2., -s / 2.]
    if bound_max is not None:
        voxel_grid_origin = bound_min
    overall_index = np.arange(0, N ** 3, 1).astype(int)
    xyz = np.zeros([N ** 3, 3])
    xyz[:, 2] = overall_index % N
    xyz[:, 1] = (overall_index / N) % N
    xyz[:, 0] = ((overall_index / N) / N) % N
    xyz[:, 0] = (xyz[:, 0] * (s / (N - 1))) + voxel_grid_origin[2]
    xyz[:, 1] = (xyz[:, 1] * (s / (N - 1))) + voxel_grid_origin[1]
    xyz[:, 2] = (xyz[:, 2] * (s / (N - 1))) + voxel_grid_origin[0]
    return torch.from_numpy(xyz).float()


def grid_case(device, N=12, dataset="DTU", volume_size=2.0, use_bounds=False, chunk=1000):
    from levels2fm_b200 import _C, ops
    opt = common.make_opt(dataset, device, 16, (None, 64, 16), 16)
    cfg = common.cfg_of(opt, 16)
    sdf_sd, _ = port.random_state(cfg, seed=8, table_std=0.2)
    sdf, _, _ = common.build_models(opt)
    sdf.load_state_dict(sdf_sd)
    bmax, bmin = (list(cfg.bound_max), list(cfg.bound_min)) if use_bounds else (None, None)
    ref_pts = reference_grid_points(N, volume_size, bmax, bmin)
    # the generated points are the reference's, bit for bit (ragged chunk: not a divisor of N^3)
    s = float(volume_size)
    origin = bmin if use_bounds else [-s / 2.0] * 3
    got = torch.cat([ops.grid_points_raw(_C.get(), N, s / (N - 1), (origin[2], origin[1], origin[0]), b, min(chunk, N ** 3 - b), device).cpu()
                     for b in range(0, N ** 3, chunk)])
    assert torch.equal(got, ref_pts)
    vol = sdf.infer_sdf_grid(N=N, volume_size=volume_size, bound_max=bmax, bound_min=bmin, chunk=chunk)
    assert vol.shape == (N, N, N)
    ref = port.infer_sdf(ref_pts, sdf_sd, cfg).reshape(N, N, N)
    assert common.rel_err(vol.cpu(), ref.detach()) < 1e-4
    # and it is what the reference's batchify loop would have produced through the drop-in infer_sdf
    direct = torch.cat([sdf.infer_sdf(ref_pts[i:i + 777].to(device)).detach().cpu() for i in range(0, N ** 3, 777)]).reshape(N, N, N)
    assert common.rel_err(vol.cpu(), direct) < 2e-6


def image_case(device, H=12, W=16, dataset="DTU", dual=False, slice_rays=50, eb=False):
    from levels2fm_b200 import rays as rays_mod
    over = {"SDF.VolSDF.volsdf_sampling": True, "SDF.VolSDF.sample_intvs": 8, "SDF.VolSDF.final_sample_intvs": 8} if eb else {}
    opt = common.make_opt(dataset, device, 16, (None, 64, 16), 16, dual, **over)
    opt.H, opt.W = H, W
    opt.Renderer.rand_rays = slice_rays
    cfg = common.cfg_of(opt, 16)
    sdf_sd, rad_sd = port.random_state(cfg, seed=8, table_std=0.1 if not eb else 0.01)
    sdf, rad, ren = common.build_models(opt)
    sdf.load_state_dict(sdf_sd)
    rad.load_state_dict(rad_sd)
    half = float(cfg.bound_max[0])
    pose = torch.tensor([[[1.0, 0.0, 0.0, 0.05 * half], [0.0, 1.0, 0.0, -0.02 * half], [0.0, 0.0, 1.0, 2.5 * half]]], device=device)
    intr = torch.tensor([[1.2 * W, 0.0, W / 2], [0.0, 1.2 * W, H / 2], [0.0, 0.0, 1.0]], device=device)
    center, ray = rays_mod.get_center_and_ray(opt, pose, intr=intr[None])
    assert center.shape == (1, H * W, 3)
    out = ren.render_image(opt, center, ray, sdf, rad)                  # ragged last slice: H*W is not a multiple of slice_rays
    assert out["rgb"].shape == (1, H * W, 3) and out["depth"].shape == (1, H * W, 1) and out["norm"].shape == (1, H * W, 3)
    assert not out["rgb"].requires_grad
    # (1) exactly what Camera.render_img_by_slices gets from one Renderer.forward per slice (pipelines/Camera.py:290-310)
    with torch.no_grad():
        parts = [ren.forward(opt, center[:, s:s + slice_rays], ray[:, s:s + slice_rays], sdf, rad) for s in range(0, H * W, slice_rays)]
    for key, k2 in (("rgb", "rgb"), ("depth", "depth_mlp"), ("norm", "normal_mlp")):
        cat = torch.cat([p[k2].reshape(1, -1, out[key].shape[-1]) for p in parts], dim=1)
        assert torch.equal(out[key], cat), key
    # (2) the oracle on the same rays
    ref = port.render_forward(center.cpu(), ray.cpu(), sdf_sd, rad_sd, cfg)
    for key, k2 in (("rgb", "rgb"), ("depth", "depth_mlp"), ("norm", "normal_mlp")):
        a, b = out[key].cpu().reshape(H * W, -1), ref[k2].detach().reshape(H * W, -1)
        if not eb:
            assert common.rel_err(a, b) < 1e-4, key
        else:
            # error-bounded depths are discontinuous functions of the SDF (threshold tests, bisection): the same statistical bar as
            # golden_checks.check_c2 -- typical ray at 1e-4, nearly all at 1e-3
            e = (a - b).abs().amax(dim=-1) / b.abs().max()
            assert e.median().item() < 1e-4 and (e < 1e-3).float().mean().item() > 0.8, (key, e.median().item(), (e < 1e-3).float().mean().item())


def values_only_kernel_case(device, n, layers, n_levels=16, ray_mode=False):
    """The two-tiles-in-flight values-only kernel (sampler rounds, infer_sdf, the SDF volume): several tile pairs per CTA, ragged
    tails, every MLP depth (1, 2, 3 hidden layers), explicit points and ray samples -- against the exact fp32-SIMT kernel of the
    same library (2e-6) and the oracle (1e-4)."""
    from levels2fm_b200 import _C, ops
    lib = _C.get()
    opt = common.make_opt("DTU", device, n_levels, layers, 16)
    cfg = common.cfg_of(opt, n_levels)
    sdf_sd, _ = port.random_state(cfg, seed=6, table_std=0.2)
    sdf, _, _ = common.build_models(opt)
    sdf.load_state_dict(sdf_sd)
    spec, table = sdf.field_spec(), sdf.table().detach()
    theta = sdf.SDF_MLP.theta().detach().contiguous()
    g = torch.Generator().manual_seed(n)
    if ray_mode:
        n_per_ray = 7
        n_rays = (n + n_per_ray - 1) // n_per_ray
        center = (torch.rand(n_rays, 3, generator=g) * 0.6 - 0.3).to(device).contiguous()
        ray = torch.nn.functional.normalize(torch.randn(n_rays, 3, generator=g), dim=-1).to(device).contiguous()
        t = (torch.rand(n_rays, n_per_ray, generator=g) * 0.6).to(device).contiguous()
        pts = ops._points(lib, None, center, ray, t)
        x = (center[:, None] + ray[:, None] * t[..., None]).reshape(-1, 3)
    else:
        x = (torch.rand(n, 3, generator=g) * 1.6 - 0.8).to(device).contiguous()
        pts = ops._points(lib, x, None, None, None)
    image = ops.field_prepare_raw(lib, spec, table, theta, None)
    y, s, _, _ = ops.field_forward_raw(lib, spec, table, theta, pts, None, want_y=True, image=image)
    y_ref, s_ref, _, _ = ops.field_forward_raw(lib, spec, table, theta, pts, None, want_y=True, simt=True)
    assert common.rel_err(y, y_ref) < 2e-6 and common.rel_err(s, s_ref) < 2e-6
    o_sdf, _ = port.infer_sdf(x.cpu(), sdf_sd, cfg, "ret_all")
    assert common.rel_err(s.cpu(), o_sdf.reshape(-1)) < 1e-4


def empty_batch_case(device):
    """Zero rays / zero points through every public entry of the path: shapes follow the reference's conventions, backward runs and
    leaves all-zero parameter gradients (no launch with an empty grid, no reshape ambiguity)."""
    for eb in (False, True):
        over = {"SDF.VolSDF.volsdf_sampling": True, "SDF.VolSDF.sample_intvs": 8, "SDF.VolSDF.final_sample_intvs": 8} if eb else {}
        opt = common.make_opt("DTU", device, 16, (None, 64, 16), 16, eb, **over)      # (dual field on the second pass)
        sdf, rad, ren = common.build_models(opt)
        center, ray = torch.zeros(1, 0, 3, device=device), torch.zeros(1, 0, 3, device=device)
        out = ren.forward(opt, center, ray, sdf, rad)
        n = 16
        assert out["rgb"].shape == (1, 0, 3) and out["normals"].shape == (1, 0, n, 3) and out["sdfs_volume"].shape == (1, 0, n, 1)
        (out["rgb"].sum() + out["normals"].sum() + out["depth_mlp"].sum()).backward()
        for p in list(sdf.parameters()) + list(rad.parameters()):
            assert p.grad is None or float(p.grad.abs().max()) == 0.0
        img = ren.render_image(opt, center, ray, sdf, rad)
        assert img["rgb"].shape == (1, 0, 3) and img["depth"].shape == (1, 0, 1)
    x = torch.zeros(0, 3, device=device)
    assert sdf.infer_sdf(x).shape == (0, 1)
    assert sdf.gradient(x.clone().requires_grad_(True)).shape == (0, 3)
    pts, _ = sdf.get_surface_pts(x)[:2]
    assert pts.shape == (0, 3)
