"""The point-only block of the reference's bundle adjustment (``BA.run_ba``, pipelines/BA.py:117-151): up to
10 x 1000 iterations per registered image on a few thousand tracked points -- small batches, launch-bound in the reference
(3 SDF evaluations with their autograd graphs, a Lie exponential per point, to_hom / matmul / division chains, two boolean-mask
gathers with host synchronisations).  Here one iteration is 7 launches forward (surface projection, SDF at the projected points,
se3 -> SE3, reprojection sums) and as many backward, has no host synchronisation, and can therefore be replayed as one CUDA graph
(``graph.GraphedStep``)."""
from __future__ import annotations

import torch

from . import ops, rays


def surface_ba_terms(sdf_func, xyzs, se3_refine, pose_idx, intrinsic, kypts2D, sdf_threshold, eps: float = 1e-6):
    """BA.py:123-148 + the "sfm" branch of compute_loss (BA.py:199-203).

    xyzs [n,3] tracked 3-D points, se3_refine [C,6], pose_idx [n] (long: the camera of every observation), intrinsic [3,3],
    kypts2D [n,2], sdf_threshold = (bound_max - bound_min)[0] / 10 / opt.Res (BA.py:115).
    -> dict: xyzs_new [n,3], sdfs [n,1], gradients [n,1] (|grad sdf| at the input points), uvs [n,2], mask_surf [n] bool,
       reproj_loss, sdf_surf (= L1(sdfs, 0)), eikonal_loss (= L1(gradients, 1)) -- all differentiable where the reference's are."""
    xyzs_new, normals_value = sdf_func.get_surface_pts(xyzs)                                  # BA.py:124 (one fused launch)
    sdfs = sdf_func.infer_sdf(xyzs_new, mode="ret_sdf").view(-1, 1)                           # BA.py:125
    poses = rays.se3_to_SE3(se3_refine[pose_idx])                                             # BA.py:127 (one launch)
    reproj, uvs, mask_surf, _ = ops.ReprojLoss.apply(xyzs_new, poses, intrinsic, kypts2D, sdfs, 2.0 * float(sdf_threshold), eps)
    return {"xyzs_new": xyzs_new, "sdfs": sdfs, "gradients": normals_value, "uvs": uvs, "mask_surf": mask_surf,
            "reproj_loss": reproj, "sdf_surf": sdfs.abs().mean(), "eikonal_loss": (normals_value - 1.0).abs().mean()}
