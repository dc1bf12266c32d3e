"""SDF field of Level-S2fM on the fused sm_100a kernels (reference: models/SDF.py).

Same constructor, attributes, method names and state-dict keys as the reference class
(``beta``, ``embed_fn.embedder_obj.params``, ``SDF_MLP.mlp.{i}.{bias,weight_g,weight_v}``), so the reference's
``pipelines/`` can use it unchanged.  Every evaluation is ONE launch of the fused hash-grid + MLP (+ normals)
kernel; autograd sees a single node whose backward is the fused backward kernel (second-order path included).
"""
from __future__ import annotations

import numpy as np
import torch
import torch.nn as nn

from .. import ops
from .base import Geometry, get_Embedder, get_layer_dims


class SDF(nn.Module):
    def __init__(self, opt):
        super().__init__()
        self.opt = opt
        dev = opt.device
        self.bound_max = torch.tensor(np.array(opt.data.bound_max), dtype=torch.float32, device=dev)[None, None, :]
        self.bound_min = torch.tensor(np.array(opt.data.bound_min), dtype=torch.float32, device=dev)[None, None, :]
        self.center = (self.bound_max + self.bound_min) / 2
        self.half_size = (self.bound_max - self.bound_min) / 2
        v = opt.SDF.VolSDF
        self.rescale = v.rescale
        self.beta_speed = v.beta_speed
        beta_init = np.log(v.beta_init) / self.beta_speed
        self.beta = nn.Parameter(torch.tensor([beta_init], dtype=torch.float32, device=dev))
        self.sdf_threshold = float(v.sdf_threshold)
        self.iters_max = int(v.iters_max_st)
        self.scale_mlp = opt.SDF.NN_Init.scale_mlp
        # sphere_tracing without its one host read-back (see the method): off by default = the reference's exact output shapes
        self.st_sync_free = False
        self.define_network(opt)

    def define_network(self, opt):
        self.embed_fn = get_Embedder(opt=opt, input_dim=3)
        self.SDF_MLP = Geometry(opt=opt, input_dim=self.embed_fn.out_dim, skip=opt.SDF.arch.skip,
                                tf_init=opt.SDF.NN_Init.tf_init, layers=get_layer_dims(opt.SDF.arch.layers))

    # ------------------------------------------------------------------ kernel plumbing
    def field_spec(self) -> ops.FieldSpec:
        inside = self.opt.data.inside == True   # noqa: E712  (the reference compares with == True)
        return ops.FieldSpec(self.embed_fn.embedder_obj.grid, [float(x) for x in self.opt.data.bound_min],
                             [float(x) for x in self.opt.data.bound_max], self.SDF_MLP.dims(), float(self.rescale),
                             100.0, 20.0, 1.0 if inside else -1.0, float(self.scale_mlp))

    def table(self):
        return self.embed_fn.embedder_obj.params

    def _bg_sdf(self) -> bool:
        return self.opt.data.inside == True and getattr(self.opt.data, "bg_sdf", False) == True  # noqa: E712

    def _eval(self, xyz, want_y=False, want_nrm=False, sdf_x_detached=False):
        """One fused launch: sdf [...,1], raw MLP output [...,k+1] | None, d sdf / d x [...,3] | None.  Differentiable w.r.t. the
        field parameters AND the positions (like tcnn's encoding in the reference)."""
        shp = xyz.shape[:-1]
        flat = xyz.reshape(-1, 3).float()
        theta = self.SDF_MLP.theta()
        image = None
        if flat.shape[0] >= ops.TC_BACKWARD_MIN_SAMPLES:      # large point sets: tensor-core operand image (streamed weights, tcgen05 backward)
            from .. import _C
            image = ops.field_prepare_raw(_C.get(), self.field_spec(), self.table().detach(), theta.detach().contiguous(), None)
        sdf, y, nrm, _ = ops.FieldEval.apply(self.field_spec(), None, self.table(), theta, None, None, None,
                                             flat, None, None, None, 0, None, want_y, want_nrm, image, sdf_x_detached)
        sdf = sdf.view(*shp, 1)
        nrm = nrm.view(*shp, 3) if want_nrm else None
        if self._bg_sdf():
            # min(sdf, bg_rad - |x|) (models/SDF.py:68-69); where the sphere term is active its gradient is -x/|x|
            xs = xyz.detach() if sdf_x_detached else xyz
            r = xs.norm(dim=-1, keepdim=True)
            bg = self.opt.data.bg_rad - r
            if nrm is not None:
                nrm = torch.where(bg < sdf, -xyz / xyz.norm(dim=-1, keepdim=True), nrm)
            sdf = torch.min(sdf, bg)
        return sdf, (y.view(*shp, -1) if want_y else None), nrm

    # ------------------------------------------------------------------ reference surface
    def infer_sdf(self, xyz, mode="ret_sdf"):
        sdf, feat, _ = self._eval(xyz, want_y=mode != "ret_sdf")
        if mode == "ret_sdf":
            return sdf
        if mode == "ret_feat":
            return feat
        return sdf, feat

    forward = infer_sdf      # north-star alias (the reference class defines no forward; its sampler calls one)

    def forward_ab(self):
        beta = torch.exp(self.beta * self.beta_speed)
        return 1.0 / beta, beta

    def sdf_to_sigma(self, sdf, alpha, beta):
        e = 0.5 * torch.exp(-torch.abs(sdf) / beta)
        return alpha * torch.where(sdf >= 0, e, 1 - e)

    def density(self, opt, xyzs):
        alpha, beta = self.forward_ab()
        return self.sdf_to_sigma(self.infer_sdf(xyzs), alpha, beta)

    def gradient(self, p):
        """d sdf / d p by the kernel's analytic reverse sweep (models/SDF.py:102-114).  Like the reference
        (create_graph=True) the result stays differentiable w.r.t. the field parameters and w.r.t. ``p`` itself
        (``p.requires_grad_`` is set in place on the caller's tensor, as the reference does)."""
        with torch.enable_grad():
            p.requires_grad_(True)
            return self._eval(p, want_nrm=True)[2]

    def sdf_and_gradient(self, p):
        sdf, _, nrm = self._eval(p, want_nrm=True)
        return sdf, nrm

    def get_surface_pts(self, pts):
        """One Newton projection onto the zero level set (models/SDF.py:95-100); one fused launch.  As in the reference the sdf
        is evaluated at ``pts.detach()`` and the normals at ``pts`` (which gets requires_grad set in place)."""
        with torch.enable_grad():
            pts.requires_grad_(True)
            sdf, _, normals = self._eval(pts, want_nrm=True, sdf_x_detached=True)
        normals_value = torch.norm(normals, dim=-1, keepdim=True)
        surf_pts = pts - normals / normals_value.detach() * sdf
        return surf_pts, normals_value

    @torch.no_grad()
    def infer_sdf_grid(self, N=512, volume_size=2.0, bound_max=None, bound_min=None, chunk=1 << 22):
        """The [N,N,N] SDF volume the reference's marching-cubes export evaluates (utils/util.py:392-430 ``extract_mesh``:
        N^3 numpy points, ``infer_sdf`` in 16 k chunks through host memory, reshape) -- generated and evaluated on the device
        in ``chunk``-point launches of the values-only tensor-core kernel; nothing but the N^3 floats ever leaves the GPU.
        Same point arithmetic as the reference (float64, its true-division quirk included), so
        ``infer_sdf_grid(N, s, bmax, bmin)`` == ``extract_mesh``'s ``out`` array."""
        from .. import _C
        lib = _C.get()
        s = float(volume_size)
        origin = [-s / 2.0] * 3
        if bound_max is not None:
            origin = [float(v) for v in bound_min]
        step = s / (N - 1)
        org = (origin[2], origin[1], origin[0])            # util.py:408-410 adds origin[2] to x, origin[0] to z
        dev = self.table().device
        spec, table = self.field_spec(), self.table().detach()
        theta = self.SDF_MLP.theta().detach().contiguous()
        image = ops.field_prepare_raw(lib, spec, table, theta, None)
        total = N ** 3
        out = torch.empty(total, device=dev)
        for begin in range(0, total, chunk):
            cnt = min(chunk, total - begin)
            xyz = ops.grid_points_raw(lib, N, step, org, begin, cnt, dev)
            pts = ops._points(lib, xyz, None, None, None)
            _, sdf, _, _ = ops.field_forward_raw(lib, spec, table, theta, pts, None, image=image)
            if self._bg_sdf():
                sdf = torch.min(sdf, self.opt.data.bg_rad - xyz.norm(dim=-1))
            out[begin:begin + cnt] = sdf
        return out.view(N, N, N)

    def sphere_tracing(self, ray0, ray_direction, model=None, c=None, tau=0.5, n_steps=(128, 129), n_secant_steps=8,
                       depth_range=(0.0, 2.4), max_points=3500000, rad=1.0, iter=0):
        """Bidirectional sphere tracing (models/SDF.py:116-226).  Returns (d_pred [B,M] differentiable w.r.t. the
        field parameters, sdf_last [B*M], sampled_pts [1,*,3] (random eikonal points), finish_mask [B*M,1])."""
        from .. import _C
        o = ray0.detach().reshape(-1, 3).float().contiguous()
        d = ray_direction.detach().reshape(-1, 3).float().contiguous()
        # the whole no-grad march (both fronts, thresholds, crossing test) is one kernel launch; the reference's
        # per-iteration host synchronisations collapse into ONE read-back of the iteration count K at the end
        track, cnt, t_near, t_far, acc_hist = ops.sphere_trace_raw(
            _C.get(), self.field_spec(), self.table().detach(), self.SDF_MLP.theta().detach().contiguous(), o, d,
            self.sdf_threshold, self.iters_max)
        if self.st_sync_free:
            # K (the iteration count of the reference's loop) stays on the device: all iters_max track rows are evaluated, rows >= K
            # are masked out of the sum and replaced by the last valid point (so the field sees finite inputs).  d_pred, sdf_last,
            # finish_mask and their gradients are those of the synchronous form; only sampled_pts differs when K < iters_max (it
            # then carries iters_max instead of K rows per ray, the extra ones repeating the last point).  No host synchronisation:
            # the whole iteration becomes capturable in a CUDA graph (graph.GraphedStep).
            is_zero = cnt[:self.iters_max] == 0
            K_t = torch.where(is_zero.any(), is_zero.float().argmax(), torch.full((), self.iters_max, device=cnt.device)).long()
            Kc = K_t.clamp_min(1)
            mask = (torch.arange(self.iters_max, device=cnt.device) < Kc)
            last = track.gather(1, (Kc - 1).view(1, 1, 1).expand(track.shape[0], 1, 3))
            pts_tracks = torch.where(mask[None, :, None], track, last)
            sdf_tracks = self.infer_sdf(pts_tracks)                                     # [M, iters_max, 1], with grad
            d_sum = (sdf_tracks * mask[None, :, None]).sum(dim=-2)
            sdf_last = sdf_tracks.gather(1, (Kc - 1).view(1, 1, 1).expand(track.shape[0], 1, 1))[:, 0, :]
            acc_e = acc_hist.index_select(0, K_t.view(1))[0]
        else:
            done = (cnt == 0).nonzero()
            n_it = int(done[0, 0]) if done.numel() else self.iters_max          # iterations the reference's loop runs
            acc_e = acc_hist[n_it]
            pts_tracks = track[:, :max(n_it, 1)]                                # [M,K,3]; K = 0 keeps the start point
            sdf_tracks = self.infer_sdf(pts_tracks)                          # [M,K,1], with grad
            d_sum = sdf_tracks.sum(dim=-2)
            sdf_last = sdf_tracks[:, -1, :]
        if torch.is_grad_enabled() and (ray0.requires_grad or ray_direction.requires_grad):
            # d_pred = sum sdf(track points, constants) + t_near, clamped to t_far: the rays enter through the slab test only
            # (models/SDF.py:120-123,204-205).  The reference's RayAABBIntersector has no backward; ours has (SURVEY 8a defect
            # iii) unless opt.Renderer.aabb_grad is False.
            if getattr(getattr(self.opt, "Renderer", None), "aabb_grad", True) == False:   # noqa: E712
                raise NotImplementedError("RayAABBIntersector has no backward in the reference (utils/custom_functions.py:10-31); "
                                          "set opt.Renderer.aabb_grad = True for the analytic slab-test gradient")
            hits = ops.RayAABB.apply(ray0.reshape(-1, 3), ray_direction.reshape(-1, 3), [float(x) for x in self.opt.data.bound_min],
                                     [float(x) for x in self.opt.data.bound_max])
            t_near, t_far = hits[:, 0], hits[:, 1]
        d_pred = d_sum.view(*ray0.shape[:-1]) + t_near.view(*ray0.shape[:-1])
        d_pred = torch.minimum(d_pred, t_far.view(*d_pred.shape))
        thr2 = float(self.opt.data.bound_max[0] - self.opt.data.bound_min[0]) / 10 / self.opt.Res
        finish_mask = sdf_last.abs() < thr2
        # random points for the eikonal term (models/SDF.py:216-224)
        tf_v, tn_v = t_far.view(*d_pred.shape), t_near.view(*d_pred.shape)
        factor_rand = torch.rand_like(d_pred)
        d_up = torch.minimum(1.5 * acc_e.view(*d_pred.shape), tf_v)
        d_sample = (1 - factor_rand) * d_up + factor_rand * tn_v
        sampled_pts = ray0 + d_sample[..., None] * ray_direction          # (callers detach it, e.g. pipelines/Camera.py:243)
        pick = torch.randperm(pts_tracks.shape[0], device=pts_tracks.device)[:4096]
        sampled_pts = torch.cat([pts_tracks[pick].view(1, -1, 3), sampled_pts.view(1, -1, 3)], dim=1)
        return d_pred, sdf_last[:, 0], sampled_pts, finish_mask
