"""Oracle: ray / axis-aligned-box slab test (``vren.ray_aabb_intersect``). [EXT]

TEST INFRASTRUCTURE -- see oracle/__init__.py.  PARITY UNPINNED for this op:
vren 2.0 (= kwea123/ngp_pl ``models/csrc``; reference env.yaml:251, install
recipe README.md:29-43) is not under /root/reference.  Restated from its
published kernel and anchored on the reference wrapper
``utils/custom_functions.py:10-31`` and call sites ``models/Renderer.py:178-179``,
``models/SDF.py:120-121`` (N_voxels = 1, max_hits = 1).
"""
import torch


def ray_aabb_t(rays_o, rays_d, center, half_size):
    """Differentiable core.  rays_o/rays_d [M,3]; center/half_size [1,3] or [3].

    Returns (t_near [M], t_far [M]) with the vren conventions:
    hit  iff  t1 <= t2 and t2 > 0  ->  (max(t1, 0), t2);  else (-1, -1).
    """
    center = center.reshape(1, 3).to(rays_o)
    half = half_size.reshape(1, 3).to(rays_o)
    inv_d = 1.0 / rays_d
    t_lo = (center - half - rays_o) * inv_d
    t_hi = (center + half - rays_o) * inv_d
    t1 = torch.minimum(t_lo, t_hi).max(dim=-1).values
    t2 = torch.maximum(t_lo, t_hi).min(dim=-1).values
    hit = (t1 <= t2) & (t2 > 0)
    minus1 = torch.full_like(t1, -1.0)
    t_near = torch.where(hit, torch.clamp_min(t1, 0.0), minus1)
    t_far = torch.where(hit, t2, minus1)
    return t_near, t_far


def ray_aabb_intersect(rays_o, rays_d, center, half_size, max_hits):
    """The vren signature: -> (hit_cnt i32 [M], hits_t f32 [M,max_hits,2], idx i64 [M,max_hits])."""
    assert center.reshape(-1, 3).shape[0] == 1, "oracle restates the single-voxel case only"
    t_near, t_far = ray_aabb_t(rays_o, rays_d, center, half_size)
    M = rays_o.shape[0]
    hit = t_far > 0
    hits_t = torch.full((M, max_hits, 2), -1.0, dtype=rays_o.dtype, device=rays_o.device)
    hits_t[:, 0, 0] = t_near
    hits_t[:, 0, 1] = t_far
    idx = torch.full((M, max_hits), -1, dtype=torch.int64, device=rays_o.device)
    idx[:, 0] = torch.where(hit, torch.zeros_like(idx[:, 0]), idx[:, 0])
    return hit.to(torch.int32), hits_t, idx
