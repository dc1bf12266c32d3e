"""Oracle: functional CPU restatement of the reference hot path (the part that travels).

TEST INFRASTRUCTURE -- see oracle/__init__.py.  Plain PyTorch (fp32, or fp64 when the
parameters are fp64), autograd-differentiable to second order, no dependency on
/root/reference.  Pinned against the reference's own python in the build container
(tests/test_oracle_vs_reference.py, fixtures from oracle/make_golden.py).

Parameters are passed as *state dicts with the reference's key layout* (SURVEY.md 5):
  SDF : beta, embed_fn.embedder_obj.params, SDF_MLP.mlp.{i}.{bias,weight_g,weight_v}
  RadF: Rad_dec.mlp_radiance.{i}.{bias,weight_g,weight_v}
        (+ embed_fn.embedder_obj.params, Geo_enc.mlp.{i}.* when dual_field)

Each function cites the reference lines it restates (paths relative to /root/reference).
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence, Tuple

import torch

from . import aabb as _aabb
from . import hashgrid as _hg

Tensor = torch.Tensor


# --------------------------------------------------------------------------- config
@dataclass
class SceneCfg:
    """The options the hot path reads (SURVEY.md 5 'Config / flags', C.2 defaults)."""
    bound_min: Sequence[float] = (-1.0, -1.0, -1.0)
    bound_max: Sequence[float] = (1.0, 1.0, 1.0)
    inside: bool = True
    bg_sdf: bool = False
    bg_rad: float = 2.0
    bgcolor: Sequence[float] = (0.0, 0.0, 0.0)
    # SDF.NN_Init / VolSDF
    scale_mlp: float = 1.0
    rescale: float = 1.0
    beta_speed: float = 1.0
    beta_init: float = 0.05
    sdf_threshold: float = 1e-3
    iters_max_st: int = 20
    res: int = 100                      # opt.Res
    # sampler
    sample_intvs: int = 128
    final_sample_intvs: int = 64
    volsdf_sampling: bool = False
    max_upsample_iter: int = 6
    max_bisection_itr: int = 10
    eps: float = 0.1
    # hash grid (options/config_hash_sdf.json; per_level_scale is overridden, base.py:128-129)
    n_levels: int = 16
    n_features: int = 2
    log2_hashmap_size: int = 19
    base_resolution: int = 16
    # networks
    sdf_layers: Sequence[Optional[int]] = (None, 64, 16)
    rad_layers: Sequence[Optional[int]] = (None, 64, 64, 3)
    dual_field: bool = False
    n_fourier: int = 4
    softplus_beta: float = 100.0
    softplus_threshold: float = 20.0

    @property
    def per_level_scale(self) -> float:
        # models/base.py:128-129
        s = (self.bound_max[0] - self.bound_min[0]) / 2
        return math.exp(math.log(2048 * s / self.base_resolution) / (self.n_levels - 1))

    def grid(self) -> _hg.GridMeta:
        return _hg.grid_meta(self.n_levels, self.n_features, self.log2_hashmap_size,
                             self.base_resolution, self.per_level_scale)

    @property
    def k_geo(self) -> int:
        return int(self.sdf_layers[-1])

    @property
    def n_sdf_layers(self) -> int:
        return len(self.sdf_layers) - 1

    @property
    def n_rad_layers(self) -> int:
        return len(self.rad_layers) - 1

    @property
    def enc_dim(self) -> int:
        return 3 + self.n_levels * self.n_features

    @property
    def view_dim(self) -> int:
        return 3 + 3 * 2 * self.n_fourier

    @property
    def rad_in_dim(self) -> int:
        # models/RadF.py:52-56
        return 3 + self.view_dim + 3 + self.k_geo * (2 if self.dual_field else 1)


def cfg_from_opt(opt) -> SceneCfg:
    """Translate a reference ``opt`` (EasyDict) into a SceneCfg."""
    import json
    import os
    path = opt.SDF.Hash_config.config_file
    if not os.path.isabs(path) and not os.path.exists(path):
        path = os.path.join("/root/reference", path)
    with open(path) as f:
        enc = json.load(f)["encoding"]
    v = opt.SDF.VolSDF
    scene = opt.data.get(opt.data.scene, None) if hasattr(opt.data, "get") else None
    bg = getattr(scene, "bgcolor", None) if scene is not None else None
    if bg is None:
        bg = opt.data.bgcolor
    return SceneCfg(
        bound_min=tuple(float(x) for x in opt.data.bound_min),
        bound_max=tuple(float(x) for x in opt.data.bound_max),
        inside=bool(opt.data.inside == True), bg_sdf=bool(opt.data.bg_sdf == True),
        bg_rad=float(opt.data.bg_rad), bgcolor=tuple(float(x) for x in bg),
        scale_mlp=float(opt.SDF.NN_Init.scale_mlp), rescale=float(v.rescale),
        beta_speed=float(v.beta_speed), beta_init=float(v.beta_init),
        sdf_threshold=float(v.sdf_threshold), iters_max_st=int(v.iters_max_st), res=int(opt.Res),
        sample_intvs=int(v.sample_intvs), final_sample_intvs=int(v.final_sample_intvs),
        volsdf_sampling=bool(v.volsdf_sampling), max_upsample_iter=int(v.max_upsample_iter),
        max_bisection_itr=int(getattr(v, "max_bisection_itr", None) or 10), eps=float(v.eps),
        n_levels=int(enc["n_levels"]), n_features=int(enc["n_features_per_level"]),
        log2_hashmap_size=int(enc["log2_hashmap_size"]), base_resolution=int(enc["base_resolution"]),
        sdf_layers=tuple(opt.SDF.arch.layers), rad_layers=tuple(opt.RadF.arch.layers),
        dual_field=bool(opt.Ablate_config.dual_field == True),
    )


# --------------------------------------------------------------------------- parameters
def layer_dims(layers):
    """utils/util.py:273-275"""
    return list(zip(layers[:-1], layers[1:]))


def effective_weight(g: Tensor, v: Tensor) -> Tensor:
    """old-style nn.utils.weight_norm(dim=0): W = g * v / ||v||_row  (models/base.py:200,241)."""
    return torch._weight_norm(v, g, 0)


def mlp_from_sd(sd: Dict[str, Tensor], prefix: str, n: int) -> List[Tuple[Tensor, Tensor]]:
    out = []
    for i in range(n):
        if f"{prefix}.{i}.weight_g" in sd:
            W = effective_weight(sd[f"{prefix}.{i}.weight_g"], sd[f"{prefix}.{i}.weight_v"])
        else:
            W = sd[f"{prefix}.{i}.weight"]
        out.append((W, sd[f"{prefix}.{i}.bias"]))
    return out


def init_geometry_sd(cfg: SceneCfg, gen: torch.Generator, prefix: str, dtype=torch.float32) -> Dict[str, Tensor]:
    """Geometric (sphere) initialisation of the SDF MLP, models/base.py:184-199 + weight_norm."""
    sd = {}
    dims = layer_dims(list(cfg.sdf_layers))
    bias0 = None
    for li, (k_in, k_out) in enumerate(dims):
        if li == 0:
            k_in = cfg.enc_dim
        last = li == len(dims) - 1
        if last:
            k_out += 1
        if last:
            W = torch.randn(k_out, k_in, generator=gen) * 1e-4 + math.sqrt(math.pi) / math.sqrt(dims[li][0])
            b = torch.full((k_out,), -bias0 if bias0 is not None else 0.0)
        elif li == 0:
            W = torch.zeros(k_out, k_in)
            W[:, :3] = torch.randn(k_out, 3, generator=gen) * (math.sqrt(2) / math.sqrt(k_out))
            b = torch.zeros(k_out)
        else:
            W = torch.randn(k_out, k_in, generator=gen) * (math.sqrt(2) / math.sqrt(k_out))
            b = torch.zeros(k_out)
        sd[f"{prefix}.{li}.weight_g"] = W.norm(dim=1, keepdim=True).to(dtype)
        sd[f"{prefix}.{li}.weight_v"] = W.to(dtype)
        sd[f"{prefix}.{li}.bias"] = b.to(dtype)
    return sd


def random_state(cfg: SceneCfg, seed: int = 0, table_std: float = 0.05, sphere_bias: float = 0.5,
                 dtype=torch.float32, generic_weights: bool = True, hash_weight_std: float = 0.0):
    """Seeded synthetic parameters with the reference's key layout.

    generic_weights=True draws every MLP weight from N(0, .) so that all gradient paths
    (hash features -> SDF, second order terms) are exercised; False uses the reference's
    geometric init (level set = sphere of radius ``sphere_bias``).
    Returns (sdf_sd, rad_sd).
    """
    g = torch.Generator().manual_seed(seed)
    meta = cfg.grid()
    sdf_sd: Dict[str, Tensor] = {}
    sdf_sd["beta"] = torch.tensor([math.log(cfg.beta_init) / cfg.beta_speed], dtype=dtype)
    sdf_sd["embed_fn.embedder_obj.params"] = (torch.randn(meta.n_params, generator=g) * table_std).to(dtype)

    def rand_mlp(prefix, dims, first_in, last_extra):
        sd = {}
        for li, (k_in, k_out) in enumerate(dims):
            if li == 0:
                k_in = first_in
            if li == len(dims) - 1:
                k_out += last_extra
            W = torch.randn(k_out, k_in, generator=g) * (1.0 / math.sqrt(k_in))
            sd[f"{prefix}.{li}.weight_v"] = W.to(dtype)
            sd[f"{prefix}.{li}.weight_g"] = (W.norm(dim=1, keepdim=True) *
                                             (0.75 + 0.5 * torch.rand(k_out, 1, generator=g))).to(dtype)
            sd[f"{prefix}.{li}.bias"] = (torch.randn(k_out, generator=g) * 0.05).to(dtype)
        return sd

    if generic_weights:
        sd = rand_mlp("SDF_MLP.mlp", layer_dims(list(cfg.sdf_layers)), cfg.enc_dim, 1)
        # keep the level set near a sphere so that rays see both signs of the SDF
        sd[f"SDF_MLP.mlp.{cfg.n_sdf_layers - 1}.bias"][0] = -sphere_bias * 0.2
        sdf_sd.update(sd)
    else:
        c2 = SceneCfg(**{**cfg.__dict__})
        sd = init_geometry_sd(c2, g, "SDF_MLP.mlp", dtype)
        sd[f"SDF_MLP.mlp.{cfg.n_sdf_layers - 1}.bias"] = torch.full_like(
            sd[f"SDF_MLP.mlp.{cfg.n_sdf_layers - 1}.bias"], -sphere_bias)
        if hash_weight_std > 0:      # let the hash features perturb the sphere a little
            W0 = sd["SDF_MLP.mlp.0.weight_v"]
            W0[:, 3:] = (torch.randn(W0.shape[0], W0.shape[1] - 3, generator=g) * hash_weight_std).to(dtype)
            sd["SDF_MLP.mlp.0.weight_g"] = W0.norm(dim=1, keepdim=True)
        sdf_sd.update(sd)
    rad_sd = rand_mlp("Rad_dec.mlp_radiance", layer_dims(list(cfg.rad_layers)), cfg.rad_in_dim, 0)
    if cfg.dual_field:
        rad_sd["embed_fn.embedder_obj.params"] = (torch.randn(meta.n_params, generator=g) * table_std).to(dtype)
        rad_sd.update(rand_mlp("Geo_enc.mlp", layer_dims(list(cfg.sdf_layers)), cfg.enc_dim, 1))
    return sdf_sd, rad_sd


# --------------------------------------------------------------------------- fields
def _bounds(cfg: SceneCfg, like: Tensor):
    bmin = torch.tensor(cfg.bound_min, dtype=like.dtype, device=like.device)
    bmax = torch.tensor(cfg.bound_max, dtype=like.dtype, device=like.device)
    return bmin, bmax


def hash_embed(x: Tensor, table: Tensor, cfg: SceneCfg) -> Tensor:
    """Embedder_Hash.forward, models/base.py:23-40:  enc = cat([x / rescale, grid((x-bmin)/(bmax-bmin))])."""
    bmin, bmax = _bounds(cfg, x)
    u = (x - bmin) / (bmax - bmin)
    h = _hg.encode(u.reshape(-1, 3), table, cfg.grid())
    return torch.cat([x / cfg.rescale, h.view(*x.shape[:-1], -1)], dim=-1)


def softplus(z: Tensor, cfg: SceneCfg) -> Tensor:
    return torch.nn.functional.softplus(z, beta=cfg.softplus_beta, threshold=cfg.softplus_threshold)


def geometry_mlp(enc: Tensor, layers: List[Tuple[Tensor, Tensor]], cfg: SceneCfg) -> Tensor:
    """Geometry.forward, models/base.py:206-217 (skip=[] as in every shipped config)."""
    h = enc
    n = len(layers)
    for li, (W, b) in enumerate(layers):
        h = torch.nn.functional.linear(h, W, b)
        if li <= n - 2:
            h = softplus(h, cfg)
    return h


def field_out(x: Tensor, table: Tensor, layers, cfg: SceneCfg) -> Tensor:
    """hash embed + Geometry MLP -> [..., k_geo + 1]  (SDF.infer_sdf 'ret_feat', RadF.Geometry_feat)."""
    return geometry_mlp(hash_embed(x, table, cfg), layers, cfg)


def infer_sdf(x: Tensor, sdf_sd, cfg: SceneCfg, mode: str = "ret_sdf"):
    """SDF.infer_sdf, models/SDF.py:55-78."""
    layers = mlp_from_sd(sdf_sd, "SDF_MLP.mlp", cfg.n_sdf_layers)
    feat = field_out(x, sdf_sd["embed_fn.embedder_obj.params"], layers, cfg)
    if cfg.inside:
        sdf = feat[..., :1] / cfg.scale_mlp
        if cfg.bg_sdf:
            sdf = torch.min(sdf, cfg.bg_rad - x.norm(dim=-1, keepdim=True))
    else:
        sdf = -feat[..., :1] / cfg.scale_mlp
    if mode == "ret_sdf":
        return sdf
    if mode == "ret_feat":
        return feat
    return sdf, feat


def sdf_gradient(x: Tensor, sdf_sd, cfg: SceneCfg) -> Tensor:
    """SDF.gradient, models/SDF.py:102-114 (create_graph=True: result stays differentiable)."""
    with torch.enable_grad():
        if not x.requires_grad:
            x.requires_grad_(True)
        y = infer_sdf(x, sdf_sd, cfg, "ret_sdf")
        (g,) = torch.autograd.grad(y, x, torch.ones_like(y), create_graph=True, retain_graph=True)
    return g


def forward_ab(sdf_sd, cfg: SceneCfg):
    """SDF.forward_ab, models/SDF.py:80-82."""
    beta = torch.exp(sdf_sd["beta"] * cfg.beta_speed)
    return 1.0 / beta, beta


def sdf_to_sigma(sdf: Tensor, alpha, beta) -> Tensor:
    """Laplace-CDF density, models/SDF.py:84-87 == models/Renderer.py:164-167."""
    e = 0.5 * torch.exp(-torch.abs(sdf) / beta)
    return alpha * torch.where(sdf >= 0, e, 1 - e)


def get_surface_pts(pts: Tensor, sdf_sd, cfg: SceneCfg):
    """SDF.get_surface_pts, models/SDF.py:95-100."""
    sdf = infer_sdf(pts.detach(), sdf_sd, cfg)
    n = sdf_gradient(pts, sdf_sd, cfg)
    nv = n.norm(dim=-1, keepdim=True)
    return pts - n / nv.detach() * sdf, nv


def fourier_embed(v: Tensor, cfg: SceneCfg) -> Tensor:
    """Embedder_Fourier.forward, models/base.py:75-97 with the config of base.py:142-151."""
    out = [v]
    for k in range(cfg.n_fourier):
        f = 2.0 ** k
        out += [torch.sin(v * f), torch.cos(v * f)]
    return torch.cat(out, dim=-1)


def radiance_mlp(inp: Tensor, layers: List[Tuple[Tensor, Tensor]]) -> Tensor:
    """Radiance.forward, models/base.py:249-261.  No hidden activation: the reference tests
    ``li <= len(self.mlp) - 2`` on an EMPTY ModuleList (base.py:230,257) -- reproduced on purpose."""
    h = inp
    for W, b in layers:
        h = torch.nn.functional.linear(h, W, b)
    return torch.sigmoid(h)


# --------------------------------------------------------------------------- sampler
def ray_aabb(center: Tensor, ray: Tensor, cfg: SceneCfg):
    """RayAABBIntersector.apply -> (t_near, t_far), shapes [...]; models/Renderer.py:178-180."""
    bmin, bmax = _bounds(cfg, center)
    c, h = (bmax + bmin) / 2, (bmax - bmin) / 2
    tn, tf = _aabb.ray_aabb_t(center.reshape(-1, 3), ray.reshape(-1, 3), c, h)
    return tn.view(center.shape[:-1]), tf.view(center.shape[:-1])


def sample_depth(t_near: Tensor, t_far: Tensor, n: int) -> Tensor:
    """Renderer.sample_depth, models/Renderer.py:118-127: deterministic mid-points. -> [..., n]"""
    i = 0.5 + torch.arange(n, device=t_near.device).to(t_near.dtype)
    return i / n * (t_far[..., None] - t_near[..., None]) + t_near[..., None]


def error_bound(d: Tensor, sdf: Tensor, alpha, beta) -> Tensor:
    """Renderer.error_bound, models/Renderer.py:330-360.  d, sdf [..., M] -> [..., M-1]."""
    sigma = sdf_to_sigma(sdf, alpha, beta)
    a = sdf.abs()
    delta = d[..., 1:] - d[..., :-1]
    R = torch.cat([torch.zeros_like(sdf[..., :1]), torch.cumsum(sigma[..., :-1] * delta, -1)], -1)[..., :-1]
    dstar = torch.clamp_min(0.5 * (a[..., :-1] + a[..., 1:] - delta), 0.0)
    err = alpha / (4 * beta) * delta ** 2 * torch.exp(-dstar / beta)
    E = torch.cumsum(err, -1)
    b = torch.exp(-R) * (torch.exp(E) - 1.0)
    return torch.where(torch.isnan(b), torch.full_like(b, float("inf")), b)


def sample_pdf_det(bins: Tensor, weights: Tensor, n_imp: int, eps: float = 1e-5) -> Tensor:
    """Renderer.sample_pdf with det=True, models/Renderer.py:362-399."""
    w = weights + 1e-5
    pdf = w / w.sum(-1, keepdim=True)
    cdf = torch.cat([torch.zeros_like(pdf[..., :1]), torch.cumsum(pdf, -1)], -1)
    u = torch.linspace(0.0, 1.0, n_imp, device=bins.device, dtype=bins.dtype).expand(*cdf.shape[:-1], n_imp).contiguous()
    inds = torch.searchsorted(cdf.detach(), u, right=False)
    lo = (inds - 1).clamp_min(0)
    hi = inds.clamp_max(cdf.shape[-1] - 1)
    c0, c1 = cdf.gather(-1, lo), cdf.gather(-1, hi)
    b0, b1 = bins.gather(-1, lo), bins.gather(-1, hi)
    den = c1 - c0
    den = torch.where(den < eps, torch.ones_like(den), den)
    return b0 + (u - c0) / den * (b1 - b0)


def opacity_to_sample(d: Tensor, sdf: Tensor, alpha, beta, n_final: int) -> Tensor:
    """Renderer.opacity_to_sample + sample_depth_from_opacity, models/Renderer.py:129-162."""
    sigma = sdf_to_sigma(sdf, alpha, beta)
    delta = d[..., 1:] - d[..., :-1]
    R = torch.cat([torch.zeros_like(sdf[..., :1]), torch.cumsum(sigma[..., :-1] * delta, -1)], -1)[..., :-1]
    op = 1 - torch.exp(-R)                                                        # [..., M-1]
    op = torch.cat([torch.zeros_like(op[..., :1]), op], -1)                       # [..., M]
    grid = torch.linspace(0, 1, n_final + 1, device=d.device, dtype=d.dtype)
    unif = (0.5 * (grid[:-1] + grid[1:])).expand(*op.shape[:-1], n_final).contiguous()
    idx = torch.searchsorted(op, unif, right=False)
    lo = (idx - 1).clamp_min(0)
    hi = idx.clamp_max(op.shape[-1] - 1)
    d0, d1 = d.gather(-1, lo), d.gather(-1, hi)
    c0, c1 = op.gather(-1, lo), op.gather(-1, hi)
    t = (unif - c0) / (c1 - c0 + 1e-8)
    return d0 + t * (d1 - d0)


def volsdf_sampling(center: Tensor, ray: Tensor, sdf_sd, cfg: SceneCfg):
    """Renderer.volsdf_sampling, models/Renderer.py:169-328 -> (t [B,R,Nout], beta_plus [B,R], iters [B,R]).

    Default branch (volsdf_sampling False): uniform mid-points, Nout = sample_intvs.
    Error-bounded branch: the *intended* algorithm with the SURVEY 8(a) a5 fixes, written
    per-ray with masks instead of boolean-index compaction (same per-ray arithmetic).
    """
    t_near, t_far = ray_aabb(center, ray, cfg)
    N = cfg.sample_intvs
    if not cfg.volsdf_sampling:
        d = sample_depth(t_near, t_far, N)
        return d, None, None
    with torch.no_grad():
        shp = t_near.shape
        tn, tf = t_near.reshape(-1), t_far.reshape(-1)
        c, r = center.reshape(-1, 3), ray.reshape(-1, 3)
        R_ = tn.shape[0]
        max_d = tf.clone()
        if bool(torch.all(max_d == -1)):
            max_d = torch.zeros_like(max_d)
        # the reference computes log(1+eps) in float32 (torch.log of a float32 tensor)
        beta = torch.sqrt(max_d ** 2 / (4 * (N - 1) * torch.log(1 + torch.tensor([cfg.eps], dtype=tn.dtype, device=tn.device))))
        alpha = 1.0 / beta
        d = sample_depth(tn, tf, N)                                              # [R,N]
        pts = c[:, None, :] + r[:, None, :] * d[..., None]
        sdf = infer_sdf(pts, sdf_sd, cfg)[..., 0]
        a_net, b_net = forward_ab(sdf_sd, cfg)
        active = error_bound(d, sdf, a_net, b_net).max(-1).values > cfg.eps       # 'mask'
        bounds = error_bound(d, sdf, alpha[:, None], beta[:, None])
        fine = torch.zeros(R_, cfg.final_sample_intvs, dtype=tn.dtype, device=tn.device)
        iters = torch.zeros(R_, dtype=tn.dtype, device=tn.device)
        conv = ~active
        if conv.any():
            fine[conv] = opacity_to_sample(d[conv], sdf[conv], a_net, b_net, cfg.final_sample_intvs)
        it = 0
        while it < cfg.max_upsample_iter and active.any():
            it += 1
            idx = active.nonzero()[:, 0]
            new_d = sample_pdf_det(d[idx], bounds[idx], N + 2)[..., 1:-1]         # [A,N]
            new_pts = c[idx, None, :] + r[idx, None, :] * new_d[..., None]
            new_sdf = infer_sdf(new_pts, sdf_sd, cfg)[..., 0]
            d_cat = torch.cat([d[idx], new_d], -1)
            s_cat = torch.cat([sdf[idx], new_sdf], -1)
            d_sorted, order = torch.sort(d_cat, -1)
            s_sorted = s_cat.gather(-1, order)
            # grow the per-ray storage (inactive rays keep zeros in the tail, never read again)
            d = torch.cat([d, torch.zeros(R_, N, dtype=d.dtype, device=d.device)], -1)
            sdf = torch.cat([sdf, torch.zeros(R_, N, dtype=d.dtype, device=d.device)], -1)
            bounds = torch.cat([bounds, torch.zeros(R_, N, dtype=d.dtype, device=d.device)], -1)
            d[idx], sdf[idx] = d_sorted, s_sorted
            still = error_bound(d_sorted, s_sorted, a_net, b_net).max(-1).values > cfg.eps
            done_idx = idx[~still]
            if done_idx.numel():
                fine[done_idx] = opacity_to_sample(d[done_idx], sdf[done_idx], a_net, b_net, cfg.final_sample_intvs)
                iters[done_idx] = it
                conv[done_idx] = True
            idx2 = idx[still]
            active = torch.zeros_like(active)
            if idx2.numel() == 0:
                break
            active[idx2] = True
            b_r = beta[idx2].clone()
            b_l = b_net * torch.ones_like(b_r)
            dd, ss = d[idx2], sdf[idx2]
            for _ in range(cfg.max_bisection_itr):
                b_m = 0.5 * (b_l + b_r)
                mx = error_bound(dd, ss, (1.0 / b_m)[:, None], b_m[:, None]).max(-1).values
                ok = mx <= cfg.eps
                b_r = torch.where(ok, b_m, b_r)
                b_l = torch.where(~ok, b_m, b_l)
            beta[idx2] = b_r
            alpha[idx2] = 1.0 / b_r
            bounds[idx2] = torch.clamp(error_bound(dd, ss, alpha[idx2][:, None], beta[idx2][:, None]), 0, 1e5)
        if (~conv).any():
            nc = (~conv).nonzero()[:, 0]
            fine[nc] = opacity_to_sample(d[nc], sdf[nc], (1.0 / beta[nc])[:, None], beta[nc][:, None],
                                         cfg.final_sample_intvs)
            iters[nc] = -1
        beta = torch.where(conv, b_net.expand_as(beta), beta)
        coarse = sample_depth(tn, tf, N)
        final = torch.sort(torch.cat([fine, coarse], -1), -1).values
        return final.view(*shp, -1), beta.view(shp), iters.view(shp)


# --------------------------------------------------------------------------- renderer
def composite(ray: Tensor, rgb_s: Tensor, sigma: Tensor, t: Tensor):
    """Renderer.composite, models/Renderer.py:33-49.  t [B,R,N], sigma [B,R,N], rgb_s [B,R,N,3]."""
    ray_len = ray.norm(dim=-1, keepdim=True)
    dist = (t[..., 1:] - t[..., :-1]) * ray_len
    sd = sigma[..., :-1] * dist
    alpha = 1 - torch.exp(-sd)
    T = torch.exp(-torch.cat([torch.zeros_like(sd[..., :1]), sd], dim=2).cumsum(dim=2))[..., :-1]
    prob = (T * alpha)[..., None]
    return (rgb_s[..., :-1, :] * prob).sum(dim=2), prob


def render_forward(center: Tensor, ray: Tensor, sdf_sd, rad_sd, cfg: SceneCfg, t: Optional[Tensor] = None):
    """Renderer.forward, models/Renderer.py:51-116.  Returns the reference's 5-key dict
    (+ 'opacity', 't' for tests)."""
    if t is None:
        t, _, _ = volsdf_sampling(center, ray, sdf_sd, cfg)
    x = center[:, :, None, :] + ray[:, :, None, :] * t[..., None]               # utils/camera.py:262-266
    a, b = forward_ab(sdf_sd, cfg)
    sdf, feat = infer_sdf(x, sdf_sd, cfg, "ret_all")
    normals = sdf_gradient(x, sdf_sd, cfg)
    ray_enc = fourier_embed(ray[..., None, :].expand_as(x), cfg)
    geo = feat[..., 1:]
    if cfg.dual_field:
        g2 = field_out(x, rad_sd["embed_fn.embedder_obj.params"],
                       mlp_from_sd(rad_sd, "Geo_enc.mlp", cfg.n_sdf_layers), cfg)
        geo = torch.cat([geo, g2[..., 1:]], dim=-1)
    rgbs = radiance_mlp(torch.cat([x, normals, ray_enc, geo], dim=-1),
                        mlp_from_sd(rad_sd, "Rad_dec.mlp_radiance", cfg.n_rad_layers))
    sigma = sdf_to_sigma(sdf, a, b)[..., 0]
    rgb, prob = composite(ray, rgbs, sigma, t)
    opacity = prob.sum(dim=2)
    bg = torch.tensor(cfg.bgcolor, dtype=rgb.dtype, device=rgb.device)
    rgb = rgb + (1 - opacity) * bg
    depth = (t[..., :-1, None] * prob).sum(dim=2) + (1 - opacity) * t[..., -1:]
    nrm = (normals[..., :-1, :] * prob).sum(dim=2) + (1 - opacity) * normals[..., -1, :]
    return {"rgb": rgb, "sdfs_volume": sdf, "normals": normals, "depth_mlp": depth, "normal_mlp": nrm,
            "opacity": opacity, "t": t, "rgbs": rgbs}


# --------------------------------------------------------------------------- sphere tracing
def sphere_tracing(ray0: Tensor, ray_dir: Tensor, sdf_sd, cfg: SceneCfg):
    """SDF.sphere_tracing, models/SDF.py:116-226 -- the deterministic outputs.

    Returns dict(d_pred [B,M] (differentiable w.r.t. the SDF parameters), sdf_last [B*M],
    finish_mask [B*M,1], n_iters K, track [B*M,K,3], acc_end [B*M]).  The random
    ``sampled_pts`` (rand_like / randperm, SDF.py:216-224) is not reproduced.
    """
    o, d = ray0.reshape(-1, 3), ray_dir.reshape(-1, 3)
    t_near, t_far = ray_aabb(o, d, cfg)
    thr = cfg.sdf_threshold
    with torch.no_grad():
        acc_s, acc_e = t_near.clone(), t_far.clone()
        p_s, p_e = o + acc_s[:, None] * d, o + acc_e[:, None] * d
        s_s = infer_sdf(p_s, sdf_sd, cfg)[:, 0].clone()
        s_e = infer_sdf(p_e, sdf_sd, cfg)[:, 0].clone()
        un_s = un_e = None
        track = []
        iters = 0
        while True:
            s_s = torch.where(s_s.abs() <= thr, torch.zeros_like(s_s), s_s)
            s_e = torch.where(s_e.abs() <= thr, torch.zeros_like(s_e), s_e)
            if un_s is None:
                un_s, un_e = s_s.abs() > thr, s_e.abs() > thr
            else:
                un_s, un_e = un_s & (s_s.abs() > thr), un_e & (s_e.abs() > thr)
            if un_s.sum() == 0 or iters == cfg.iters_max_st:
                break
            iters += 1
            acc_s = torch.minimum(acc_s + s_s, t_far)       # where(x > max, max, x)
            acc_e = torch.minimum(acc_e + s_e, t_far)
            track.append(p_s.clone())
            p_s, p_e = o + acc_s[:, None] * d, o + acc_e[:, None] * d
            if un_s.any():
                s_s = s_s.clone()
                s_s[un_s] = infer_sdf(p_s[un_s], sdf_sd, cfg)[:, 0]
            if un_e.any():
                s_e = s_e.clone()
                s_e[un_e] = infer_sdf(p_e[un_e], sdf_sd, cfg)[:, 0]
            un_s, un_e = un_s & (acc_s < acc_e), un_e & (acc_s < acc_e)
        if not track:
            track = [p_s.clone()]
        pts = torch.stack(track, dim=1)                     # [M,K,3]
    sdf_tr = infer_sdf(pts, sdf_sd, cfg)                    # [M,K,1] with grad
    d_pred = sdf_tr.sum(dim=-2).view(ray0.shape[:-1]) + t_near.view(ray0.shape[:-1])
    d_pred = torch.minimum(d_pred, t_far.view(d_pred.shape))
    thr2 = (cfg.bound_max[0] - cfg.bound_min[0]) / 10 / cfg.res
    finish = sdf_tr[:, -1, :].abs() < thr2
    return {"d_pred": d_pred, "sdf_last": sdf_tr[:, -1, 0], "finish_mask": finish, "n_iters": pts.shape[1],
            "track": pts, "acc_end": acc_e}


# --------------------------------------------------------------------------- ray generation (caller side)
def get_center_and_ray(pose: Tensor, intr: Tensor, xy: Tensor):
    """utils/camera.py:230-252 (+ to_hom / img2cam / cam2world / Pose.invert, camera.py:200-217, 37-43):
    pose [B,3,4] world->camera, intr [B,3,3], xy [N,2] pixel centres -> center [B,N,3], ray [B,N,3] (un-normalised)."""
    B = pose.shape[0]
    xyb = xy.repeat(B, 1, 1)
    hom = torch.cat([xyb, torch.ones_like(xyb[..., :1])], dim=-1)
    grid = hom @ intr.inverse().transpose(-1, -2)
    R, t = pose[..., :3], pose[..., 3:]
    R_inv = R.transpose(-1, -2)
    t_inv = (-R_inv @ t)[..., 0]
    pose_inv = torch.cat([R_inv, t_inv[..., None]], dim=-1)

    def cam2world(X):
        Xh = torch.cat([X, torch.ones_like(X[..., :1])], dim=-1)
        return Xh @ pose_inv.transpose(-1, -2)
    gw = cam2world(grid)
    cw = cam2world(torch.zeros_like(grid))
    return cw, gw - cw


def se3_to_SE3(wu: Tensor) -> Tensor:
    """Lie.se3_to_SE3, utils/camera.py:85-96, with the Taylor coefficients of camera.py:119-142 (nth = 10):
    wu [...,6] -> Rt [...,3,4] = [I + A wx + B wx^2 | (I + B wx + C wx^2) u]."""
    w, u = wu.split([3, 3], dim=-1)
    w0, w1, w2 = w.unbind(dim=-1)
    O = torch.zeros_like(w0)
    wx = torch.stack([torch.stack([O, -w2, w1], dim=-1), torch.stack([w2, O, -w0], dim=-1), torch.stack([-w1, w0, O], dim=-1)], dim=-2)
    theta = w.norm(dim=-1)[..., None, None]
    eye = torch.eye(3, device=w.device, dtype=torch.float32)

    def taylor(x, first, step):
        ans, denom = torch.zeros_like(x), 1.0
        for i in range(11):
            if first is None:
                if i > 0:
                    denom *= (2 * i) * (2 * i + 1)
            else:
                denom *= (2 * i + first) * (2 * i + first + 1)
            ans = ans + (-1) ** i * x ** (2 * i) / denom
        return ans
    A, B, C = taylor(theta, None, 0), taylor(theta, 1, 0), taylor(theta, 2, 0)
    R = eye + A * wx + B * wx @ wx
    V = eye + B * wx + C * wx @ wx
    return torch.cat([R, V @ u[..., None]], dim=-1)
