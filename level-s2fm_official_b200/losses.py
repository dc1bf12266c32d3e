"""Loss tail of the render path on the fused ``ls2fm_render_tail`` kernels.

``camera_render_tail`` computes what the tail of the reference's ``CameraSet.render`` (pipelines/Camera.py:506-537) adds to
``ret`` -- ``mask_bg``, ``rgb_loss``, ``DC_loss``, ``PSNR`` -- together with the stage's eikonal term (BA.py:192-193 masked by
``mask_bg``; rendering_refine.py:101-102 / Initialization.py:256-257 over all samples) and their weighted sum
``10**w_rgb * rgb_loss + 10**w_eik * eikonal + 10**w_dc * DC_loss`` (``summarize_loss``, BA.py:206-220), in two launches with all
gradients produced in the same passes."""
from __future__ import annotations

import torch

from . import ops


def camera_render_tail(ret, rgbs_gt, d_points=None, mask_finish=None, log10_weights=(3.0, 2.0, 0.0), eik_masked=True):
    """ret: the dict ``Renderer.forward`` returned (``rgb``, ``normals``, ``depth_mlp``); rgbs_gt [B,R,3];
    d_points [B,R] / mask_finish [B*R,1] from ``SDF.sphere_tracing`` (optional).
    log10_weights = (rgb, eikonal_loss, DC_Loss) of ``opt.loss_weight.<stage>`` (None = term not weighted in).
    -> dict(loss, rgb_loss, eikonal_loss, DC_loss, PSNR, mask_bg [B,R], mask_finish [B,R])."""
    w = [0.0 if x is None else 10.0 ** float(x) for x in log10_weights]
    B, R = ret["rgb"].shape[:2]
    total, l1, eik, dc, psnr, m_bg, m_fin = ops.RenderTail.apply(
        ret["rgb"], rgbs_gt, ret.get("normals"), ret.get("depth_mlp") if d_points is not None else None, d_points, mask_finish,
        w[0], w[1], w[2], bool(eik_masked))
    return {"loss": total, "rgb_loss": l1, "eikonal_loss": eik, "DC_loss": dc, "PSNR": psnr,
            "mask_bg": m_bg.view(B, R), "mask_finish": m_fin.view(B, R)}
