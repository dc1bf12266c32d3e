"""Container-only: the oracle port against the UNMODIFIED reference python (imported read-only from /root/reference through
oracle/ref_shim.py).  Skipped where the reference tree does not exist (the GPU box); the committed fixtures under
tests/golden carry the same pin there."""
import os

import pytest
import torch

from oracle import port, ref_shim

pytestmark = pytest.mark.skipif(not ref_shim.available(), reason="/root/reference not present")


@pytest.fixture(scope="module")
def ref():
    return ref_shim.load()


def test_ray_generation_matches_reference(ref):
    g = torch.Generator().manual_seed(0)
    B, N = 2, 50
    R, _ = torch.linalg.qr(torch.randn(B, 3, 3, generator=g))
    pose = torch.cat([R, torch.randn(B, 3, 1, generator=g)], dim=-1)
    intr = torch.tensor([[1920.0, 0.0, 800.0], [0.0, 1900.0, 600.0], [0.0, 0.0, 1.0]]).repeat(B, 1, 1)
    opt = ref_shim.make_opt("DTU")
    xy_full = ref.camera.mesh_grid(opt)
    idx = torch.randperm(xy_full.shape[0], generator=g)[:N]
    c_ref, r_ref = ref.camera.get_center_and_ray(opt, pose, intr=intr, rays_idx=idx, xy_grid=xy_full)
    c, r = port.get_center_and_ray(pose, intr, xy_full[idx])
    assert torch.equal(c, c_ref) and torch.equal(r, r_ref)


def test_render_forward_backward_matches_reference(ref):
    L, N, Rn = 4, 24, 20
    opt = ref_shim.make_opt("bmvs", **{"SDF.VolSDF.sample_intvs": N, "Ablate_config.dual_field": True})
    hash_cfg = dict(otype="HashGrid", n_levels=L, n_features_per_level=2, log2_hashmap_size=19, base_resolution=16, per_level_scale=1.38)
    sdf, rad, ren = ref_shim.build_models(opt, hash_config=hash_cfg)
    cfg = port.SceneCfg(bound_min=(-2.0,) * 3, bound_max=(2.0,) * 3, inside=True, bgcolor=(1.0, 1.0, 1.0), scale_mlp=3.0,
                        n_levels=L, sample_intvs=N, dual_field=True)
    sdf_sd, rad_sd = port.random_state(cfg, seed=3, table_std=0.1)
    assert sorted(sdf.state_dict()) == sorted(sdf_sd) and sorted(rad.state_dict()) == sorted(rad_sd)
    sdf.load_state_dict(sdf_sd)
    rad.load_state_dict(rad_sd)
    for sd in (sdf_sd, rad_sd):
        for k in sd:
            sd[k] = sd[k].clone().requires_grad_(True)
    g = torch.Generator().manual_seed(1)
    center = (torch.randn(1, Rn, 3, generator=g) * 0.1 + torch.tensor([0.0, 0.0, -2.5])) * 2
    ray = torch.randn(1, Rn, 3, generator=g) * 0.2 + torch.tensor([0.0, 0.0, 1.0])
    out_ref = ren.forward(opt, center, ray, sdf, rad)
    out = port.render_forward(center, ray, sdf_sd, rad_sd, cfg)
    for k in ("rgb", "sdfs_volume", "normals", "depth_mlp", "normal_mlp"):
        assert torch.allclose(out[k], out_ref[k], rtol=1e-5, atol=1e-6), k
    def loss(o):
        return o["rgb"].sum() + (o["normals"].norm(dim=-1) - 1).abs().mean() + o["depth_mlp"].sum()
    loss(out_ref).backward()
    loss(out).backward()
    for mod, sd in ((sdf, sdf_sd), (rad, rad_sd)):
        for k, p in mod.named_parameters():
            assert torch.allclose(p.grad, sd[k].grad, rtol=1e-4, atol=1e-6), k


def test_se3_to_SE3_matches_reference():
    """oracle/port.se3_to_SE3 vs the reference's own Lie.se3_to_SE3 (utils/camera.py:85-96), values and gradients."""
    ref = ref_shim.load()
    g = torch.Generator().manual_seed(0)
    wu = torch.randn(9, 6, generator=g) * torch.tensor([0.7, 0.7, 0.7, 2.0, 2.0, 2.0])
    wu[0, :3] = 0.0                          # the identity rotation (theta = 0)
    wu[1, :3] *= 1e-4
    w_ref, w_port = wu.clone().requires_grad_(True), wu.clone().requires_grad_(True)
    a, b = ref.camera.lie.se3_to_SE3(w_ref), port.se3_to_SE3(w_port)
    assert torch.allclose(a, b, atol=1e-6, rtol=1e-6)
    c = torch.randn(a.shape, generator=g)
    (a * c).sum().backward()
    (b * c).sum().backward()
    assert torch.allclose(w_ref.grad[1:], w_port.grad[1:], atol=1e-5, rtol=1e-5)
