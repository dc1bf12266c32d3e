"""Per-ray SDF volume renderer of Level-S2fM on the fused sm_100a kernels (reference: models/Renderer.py).

``Renderer.forward(opt, center, ray, SDF_Field, Rad_Field)`` returns the reference's dict
(rgb, sdfs_volume, normals, depth_mlp, normal_mlp).  Launches per call: depth sampler, [second field when
dual_field], fused field kernel (hash + MLP + normals + radiance), compositing.  Backward: compositing backward,
fused field backward.
"""
from __future__ import annotations

import numpy as np
import torch

from .. import _C, ops


class Renderer:
    def __init__(self, opt):
        self.opt = opt
        dev = opt.device
        self.bound_max = torch.tensor(np.array(opt.data.bound_max), dtype=torch.float32, device=dev)[None, None, :]
        self.bound_min = torch.tensor(np.array(opt.data.bound_min), dtype=torch.float32, device=dev)[None, None, :]
        self.center = (self.bound_max + self.bound_min) / 2
        self.half_size = (self.bound_max - self.bound_min) / 2
        scene = opt.data[f"{opt.data.scene}"] if f"{opt.data.scene}" in opt.data else None
        bgcolor = getattr(scene, "bgcolor", None) if scene is not None else None
        if bgcolor is None:
            bgcolor = opt.data.bgcolor
        self.bgcolor = torch.tensor(np.array(bgcolor), dtype=torch.float32, device=dev)
        self._bg = tuple(float(x) for x in np.array(bgcolor).reshape(-1)[:3])

    # ------------------------------------------------------------------ samplers
    def sample_depth(self, opt, min_d=None, max_d=None):
        """(i + 0.5) / N * (max_d - min_d) + min_d, [..., N, 1] (models/Renderer.py:118-127)."""
        n = opt.SDF.VolSDF.sample_intvs
        i = 0.5 + torch.arange(n, device=min_d.device)[None, None, :, None].float()
        return i / n * (max_d[..., None, :] - min_d[..., None, :]) + min_d[..., None, :]

    def _prepare(self, SDF_Field, Rad_Field=None):
        """Effective weights (weight-norm, composed radiance map) and their tensor-core operand image, built once per
        call and shared by every launch that evaluates the same parameters (sampler rounds + render forward)."""
        from .base import prepare_params
        lib = _C.get()
        theta, w_eff, b_eff = prepare_params(SDF_Field.SDF_MLP, Rad_Field.Rad_dec if Rad_Field is not None else None)
        rad = None
        if Rad_Field is not None:
            rad = ops._rad(lib, Rad_Field.rad_spec(), w_eff.detach(), b_eff.detach(), None)
        image = ops.field_prepare_raw(lib, SDF_Field.field_spec(), SDF_Field.table().detach(), theta.detach().contiguous(), rad)
        return {"theta": theta, "w_eff": w_eff, "b_eff": b_eff, "image": image}

    def volsdf_sampling(self, opt, center, ray, SDF_Field, det=True, prepared=None):
        """Depth samples [B,R,N] (returned three times like the reference's default branch).

        Gradients: the reference's RayAABBIntersector defines no backward (utils/custom_functions.py:10-31), so a loss that
        reaches grad-requiring rays through the uniform depths raises there.  Here the slab test has its analytic VJP
        (``ops.UniformDepths``; SURVEY 8a defect iii) unless ``opt.Renderer.aabb_grad`` is set to False, in which case the
        reference's behaviour (NotImplementedError) is reproduced.  The error-bounded branch runs under no_grad in the
        reference (models/Renderer.py:185) and yields constant depths here too."""
        B, R = center.shape[:2]
        v = opt.SDF.VolSDF
        if v.volsdf_sampling == True:   # noqa: E712
            from .. import sampler
            c2, r2 = center.detach().reshape(-1, 3).float().contiguous(), ray.detach().reshape(-1, 3).float().contiguous()
            prepared = prepared or self._prepare(SDF_Field)
            t, beta_plus, iters = sampler.error_bounded(self, opt, c2, r2, SDF_Field, prepared)
            return t.view(B, R, t.shape[-1]), beta_plus.view(B, R), iters.view(B, R)
        bmin, bmax = [float(x) for x in opt.data.bound_min], [float(x) for x in opt.data.bound_max]
        if torch.is_grad_enabled() and (center.requires_grad or ray.requires_grad):
            if getattr(getattr(opt, "Renderer", None), "aabb_grad", True) == False:   # noqa: E712
                raise NotImplementedError("RayAABBIntersector has no backward in the reference (utils/custom_functions.py:10-31); "
                                          "set opt.Renderer.aabb_grad = True for the analytic slab-test gradient")
            t = ops.UniformDepths.apply(center.reshape(-1, 3).float(), ray.reshape(-1, 3).float(), int(v.sample_intvs), bmin, bmax)
        else:
            c2, r2 = center.detach().reshape(-1, 3).float().contiguous(), ray.detach().reshape(-1, 3).float().contiguous()
            t, _ = ops.sample_uniform_raw(_C.get(), c2, r2, int(v.sample_intvs), bmin, bmax)
        t = t.view(B, R, t.shape[-1])       # (explicit sample count: an empty ray batch keeps its [B, 0, N] shape)
        return t, t, t

    # ------------------------------------------------------------------ the hot path
    def forward(self, opt, center, ray, SDF_Field, Rad_Field):
        prepared = self._prepare(SDF_Field, Rad_Field)
        t, _, _ = self.volsdf_sampling(opt, center, ray, SDF_Field=SDF_Field, prepared=prepared)
        return self.render_with_depths(opt, center, ray, t, SDF_Field, Rad_Field, prepared=prepared)

    def render_with_depths(self, opt, center, ray, t, SDF_Field, Rad_Field, prepared=None):
        """Everything of Renderer.forward after the depth sampler (models/Renderer.py:57-116) for given depths t [B,R,N].
        Differentiable w.r.t. the field parameters and w.r.t. center / ray / t (sample positions x = center + ray t, the Fourier
        embedding of the direction and the interval lengths |ray| dt)."""
        if SDF_Field._bg_sdf():
            raise NotImplementedError("opt.data.bg_sdf (background-sphere clamp, models/SDF.py:68-69) is not fused into the render "
                                      "kernels; no shipped config sets it")
        prepared = prepared or self._prepare(SDF_Field, Rad_Field)
        B, R = center.shape[:2]
        N = t.shape[-1]
        c2, r2 = center.reshape(-1, 3).float(), ray.reshape(-1, 3).float()
        t2 = t.reshape(B * R, N)
        geo2 = None
        if Rad_Field.dual_field:
            theta2 = Rad_Field.Geo_enc.theta()
            table2 = Rad_Field.embed_fn.embedder_obj.params
            image2 = ops.field_prepare_raw(_C.get(), Rad_Field.field_spec(), table2.detach(), theta2.detach().contiguous(), None)
            _, geo2, _, _ = ops.FieldEval.apply(Rad_Field.field_spec(), None, table2, theta2, None, None, None, None, c2, r2, t2, 0, None,
                                                True, False, image2)
        sdf, _, nrm, rgbs = ops.FieldEval.apply(SDF_Field.field_spec(), Rad_Field.rad_spec(), SDF_Field.table(),
                                                prepared["theta"], prepared["w_eff"], prepared["b_eff"], geo2, None, c2, r2, t2,
                                                0, None, False, True, prepared["image"])
        rgb, depth, normal, _ = ops.Composite.apply(r2, t2, sdf.view(B * R, N), rgbs.view(B * R, N, 3),
                                                    nrm.view(B * R, N, 3), SDF_Field.beta, float(SDF_Field.beta_speed),
                                                    self._bg)
        return {"rgb": rgb.view(B, R, 3), "sdfs_volume": sdf.view(B, R, N, 1), "normals": nrm.view(B, R, N, 3),
                "depth_mlp": depth.view(B, R, 1), "normal_mlp": normal.view(B, R, 3)}

    @torch.no_grad()
    def render_image(self, opt, center, ray, SDF_Field, Rad_Field, slice_rays=None):
        """No-grad rendering of MANY rays (a whole image) in slices: what ``Camera.render_img_by_slices`` does with one
        ``Renderer.forward`` per ``opt.Renderer.rand_rays`` rays (pipelines/Camera.py:275-311), returning its dict
        ``depth [B,HW,1]``, ``norm [B,HW,3]``, ``rgb [B,HW,3]``.  Differences from calling ``forward`` per slice: the effective
        weights and their tensor-core operand image are built once for the whole image, nothing per-sample survives a slice,
        no autograd tape."""
        if SDF_Field._bg_sdf():
            raise NotImplementedError("opt.data.bg_sdf is not fused into the render kernels")
        lib = _C.get()
        B, HW = center.shape[:2]
        n_slice = int(slice_rays or opt.Renderer.rand_rays)
        prepared = self._prepare(SDF_Field, Rad_Field)
        theta = prepared["theta"].detach().contiguous()
        w_eff, b_eff = prepared["w_eff"].detach().contiguous(), prepared["b_eff"].detach().contiguous()
        spec, rs = SDF_Field.field_spec(), Rad_Field.rad_spec()
        table = SDF_Field.table().detach()
        dual = Rad_Field.dual_field
        if dual:
            spec2, table2 = Rad_Field.field_spec(), Rad_Field.embed_fn.embedder_obj.params.detach()
            theta2 = Rad_Field.Geo_enc.theta().detach().contiguous()
            image2 = ops.field_prepare_raw(lib, spec2, table2, theta2, None)
        dev = table.device
        depth = torch.empty(B, HW, 1, device=dev)
        norm = torch.empty(B, HW, 3, device=dev)
        rgb = torch.empty(B, HW, 3, device=dev)
        for start in range(0, HW, n_slice):
            end = min(start + n_slice, HW)
            c2 = center[:, start:end].reshape(-1, 3).float().contiguous()
            r2 = ray[:, start:end].reshape(-1, 3).float().contiguous()
            t, _, _ = self.volsdf_sampling(opt, c2[None], r2[None], SDF_Field, prepared=prepared)
            t2 = t.reshape(c2.shape[0], -1).contiguous()
            geo2 = None
            if dual:
                pts2 = ops._points(lib, None, c2, r2, t2)
                geo2, _, _, _ = ops.field_forward_raw(lib, spec2, table2, theta2, pts2, None, want_y=True, want_sdf=False, image=image2)
            pts = ops._points(lib, None, c2, r2, t2)
            rad = ops._rad(lib, rs, w_eff, b_eff, geo2)
            _, sdf, nrm, rgbs = ops.field_forward_raw(lib, spec, table, theta, pts, rad, want_nrm=True, want_rgb=True,
                                                      image=prepared["image"])
            N = t2.shape[-1]
            o_rgb, o_depth, o_nrm, _ = ops.composite_forward_raw(lib, r2, t2, sdf.view(-1, N), rgbs.view(-1, N, 3), nrm.view(-1, N, 3),
                                                                 SDF_Field.beta.detach(), float(SDF_Field.beta_speed), self._bg)
            depth[:, start:end] = o_depth.view(B, end - start, 1)
            norm[:, start:end] = o_nrm.view(B, end - start, 3)
            rgb[:, start:end] = o_rgb.view(B, end - start, 3)
        return {"depth": depth, "norm": norm, "rgb": rgb}

    render_rays = forward    # north-star alias

    # ------------------------------------------------------------------ API-compatible pieces
    # The reference exposes its building blocks as methods; the fused kernels do not call them (the density, the
    # compositing, the error bound and the inverse-CDF sampling all live inside the sampler / compositing kernels), but
    # they stay callable with the reference's signatures and semantics for callers that use them on their own tensors.
    def sdf_to_sigma(self, sdf, alpha, beta):
        """Laplace-CDF density (models/Renderer.py:164-167)."""
        e = 0.5 * torch.exp(-torch.abs(sdf) / beta)
        return alpha * torch.where(sdf >= 0, e, 1 - e)

    def composite(self, ray, rgb_samples, density_samples, depth_samples):
        """(models/Renderer.py:33-49)  ray [B,R,3], rgb [B,R,N,3], density [B,R,N], depth [B,R,N,1] -> rgb [B,R,3], prob [B,R,N-1,1]."""
        ray_length = ray.norm(dim=-1, keepdim=True)
        dist = (depth_samples[..., 1:, 0] - depth_samples[..., :-1, 0]) * ray_length
        sd = density_samples[..., :-1] * dist
        alpha = 1 - torch.exp(-sd)
        T = torch.exp(-torch.cat([torch.zeros_like(sd[..., :1]), sd], dim=2).cumsum(dim=2))[..., :-1]
        prob = (T * alpha)[..., None]
        return (rgb_samples[..., :-1, :] * prob).sum(dim=2), prob

    def error_bound(self, d_vals, sdf, alpha, beta):
        """VolSDF opacity error bound per interval (models/Renderer.py:330-360): [..., M] -> [..., M-1]."""
        sigma = self.sdf_to_sigma(sdf, alpha, beta)
        delta = d_vals[..., 1:] - d_vals[..., :-1]
        R = torch.cat([torch.zeros_like(sdf[..., :1]), torch.cumsum(sigma[..., :-1] * delta, dim=-1)], dim=-1)[..., :-1]
        d_star = torch.clamp_min(0.5 * (sdf.abs()[..., :-1] + sdf.abs()[..., 1:] - delta), 0.0)
        E = torch.cumsum(alpha / (4 * beta) * delta ** 2 * torch.exp(-d_star / beta), dim=-1)
        bound = torch.exp(-R) * (torch.exp(E) - 1.0)
        return torch.where(torch.isnan(bound), torch.full_like(bound, float("inf")), bound)

    def sample_pdf(self, bins, weights, N_importance, det=False, eps=1e-5):
        """NeRF-style inverse-CDF sampling (models/Renderer.py:362-399)."""
        w = weights + 1e-5
        pdf = w / w.sum(-1, keepdim=True)
        cdf = torch.cat([torch.zeros_like(pdf[..., :1]), torch.cumsum(pdf, -1)], -1)
        if det:
            u = torch.linspace(0.0, 1.0, N_importance, device=bins.device, dtype=bins.dtype).expand(*cdf.shape[:-1], N_importance)
        else:
            u = torch.rand(*cdf.shape[:-1], N_importance, device=bins.device, dtype=bins.dtype)
        u = u.contiguous()
        inds = torch.searchsorted(cdf.detach(), u, right=False)
        below, above = (inds - 1).clamp_min(0), inds.clamp_max(cdf.shape[-1] - 1)
        c0, c1 = cdf.gather(-1, below), cdf.gather(-1, above)
        b0, b1 = bins.gather(-1, below), bins.gather(-1, above)
        denom = c1 - c0
        denom = torch.where(denom < eps, torch.ones_like(denom), denom)
        return b0 + (u - c0) / denom * (b1 - b0)
