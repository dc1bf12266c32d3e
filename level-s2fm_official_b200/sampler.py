"""Error-bounded (VolSDF) depth sampler front end (reference: Renderer.volsdf_sampling, models/Renderer.py:186-328)."""
from __future__ import annotations

import torch

from . import _C, ops


@torch.no_grad()
def error_bounded(renderer, opt, center, ray, SDF_Field, prepared=None):
    """center/ray [M,3] contiguous fp32 on the device -> (t [M, N+Nf], beta_plus [M], iters [M]).

    The whole algorithm (bound evaluation, bisection on beta+, inverse-CDF upsampling, merging, device-side compaction
    of the active rays, SDF evaluation of the new samples) runs inside libls2fm without a host synchronisation."""
    v = opt.SDF.VolSDF
    theta = (prepared["theta"] if prepared else SDF_Field.SDF_MLP.theta()).detach().contiguous()
    image = prepared["image"] if prepared else None
    max_bis = getattr(v, "max_bisection_itr", None) or 10      # undefined in the reference; VolSDF's default
    return ops.sample_error_bounded_raw(
        _C.get(), SDF_Field.field_spec(), SDF_Field.table().detach(), theta, SDF_Field.beta.detach(), center, ray,
        int(v.sample_intvs), int(v.final_sample_intvs), int(v.max_upsample_iter), int(max_bis), float(v.eps),
        float(SDF_Field.beta_speed), image=image)
