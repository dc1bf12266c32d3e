"""Seam B: drop-in stand-ins for the two native packages of the reference, for a maintainer who keeps the
reference's own ``models/`` and only swaps the third-party kernels:

    import levels2fm_b200.compat as compat
    sys.modules["tinycudann"] = compat.tinycudann      # models/base.py:5,17,37   (tcnn.Encoding)
    sys.modules["vren"] = compat.vren                  # utils/custom_functions.py:4,31   (vren.ray_aabb_intersect)

Both run on ``libls2fm_sm100.so`` (``ls2fm_grid_encode`` / ``ls2fm_grid_encode_backward``, ``ls2fm_ray_aabb``); with this seam the MLPs
and the renderer stay eager PyTorch, so only the hash grid and the slab test are B200-native (seam A is the fast path)."""
from __future__ import annotations

import types

import torch

from . import _C, ops
from .models.base import Encoding


def ray_aabb_intersect(rays_o, rays_d, center, half_size, max_hits):
    """``vren.ray_aabb_intersect`` for the only configuration the reference uses (one box, ``max_hits = 1``):
    -> (hit_cnt int32 [M], hits_t float32 [M,1,2] = (max(t1,0), t2) or (-1,-1), hits_voxel_idx int64 [M,1])."""
    if max_hits != 1 or center.reshape(-1, 3).shape[0] != 1:
        raise NotImplementedError("ray_aabb_intersect: one voxel, max_hits = 1 (utils/custom_functions.py:31 call sites)")
    o, d = rays_o.detach().float().contiguous(), rays_d.detach().float().contiguous()
    hits, cnt = ops.ray_aabb_raw(_C.get(), o, d, center.reshape(3).tolist(), half_size.reshape(3).tolist())
    idx = torch.where(cnt[:, None] > 0, 0, -1).long()
    return cnt, hits.view(-1, 1, 2), idx


tinycudann = types.ModuleType("tinycudann")
tinycudann.Encoding = Encoding
vren = types.ModuleType("vren")
vren.ray_aabb_intersect = ray_aabb_intersect
