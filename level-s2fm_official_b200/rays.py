"""Ray generation in front of the renderer (reference: utils/camera.py:230-252, the caller side of the hot path;
SURVEY.md 8f row 2)."""
from __future__ import annotations

import torch

from . import ops


def mesh_grid(opt):
    """Pixel-centre grid [H*W, 2] (x, y), utils/camera.py:253-261."""
    y = torch.arange(opt.H, dtype=torch.float32, device=opt.device) + 0.5
    x = torch.arange(opt.W, dtype=torch.float32, device=opt.device) + 0.5
    Y, X = torch.meshgrid(y, x, indexing="ij")
    return torch.stack([X, Y], dim=-1).view(-1, 2)


def get_center_and_ray(opt, pose, intr=None, rays_idx=None, xy_grid=None):
    """Same signature and results as the reference's get_center_and_ray: pose [B,3,4] world->camera, intr [B,3,3]
    -> (center [B,N,3], ray [B,N,3]), rays un-normalised.  One kernel; differentiable w.r.t. the pose."""
    assert opt.camera.model == "perspective" if hasattr(opt, "camera") else True
    with torch.no_grad():
        xy = mesh_grid(opt) if xy_grid is None else xy_grid[rays_idx, :]
        kinv = intr.inverse()
    # the reference broadcasts the intrinsics through a matmul: its own callers pass intr[None] ([1,3,3]) next to a
    # multi-camera pose [B,3,4] (pipelines/Camera.py:472); the kernel indexes one K^-1 per camera
    kinv = kinv.reshape(-1, 3, 3)
    if kinv.shape[0] not in (1, len(pose)):
        raise ValueError(f"intr has {kinv.shape[0]} matrices for {len(pose)} cameras")
    kinv = kinv.expand(len(pose), 3, 3).contiguous()
    return ops.GenerateRays.apply(pose, kinv, xy)


def se3_to_SE3(wu):
    """``camera.lie.se3_to_SE3`` (utils/camera.py:85-96): wu [...,6] -> [...,3,4], one kernel, differentiable."""
    return ops.Se3ToSE3.apply(wu)
