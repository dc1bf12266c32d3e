"""Operators of the Level-S2fM render hot path on top of ``libls2fm_sm100.so``.

Low-level ``*_raw`` functions are 1:1 with the C ABI (``include/ls2fm.h``); the
``torch.autograd.Function`` classes give them the autograd behaviour of the reference's
eager graph (models/SDF.py, models/RadF.py, models/Renderer.py of the reference).
PyTorch is only the owner of device memory, streams and the autograd tape here.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field as _dc_field
from typing import List, Optional, Sequence, Tuple

import torch

from . import _C


# ------------------------------------------------------------------------- specs
@dataclass
class GridSpec:
    """tcnn-style multiresolution hash grid as the reference configures it (models/base.py:124-139)."""
    n_levels: int = 16
    n_features: int = 2
    log2_hashmap_size: int = 19
    base_resolution: int = 16
    per_level_scale: float = 1.3819
    levels: list = _dc_field(default_factory=list)   # filled by resolve()
    n_entries: int = 0

    def resolve(self, lib: Optional[_C.Lib] = None) -> "GridSpec":
        lib = lib or _C.get()
        self.levels, self.n_entries = lib.grid_meta(self.n_levels, self.n_features, self.log2_hashmap_size,
                                                    self.base_resolution, self.per_level_scale)
        return self

    @property
    def n_params(self) -> int:
        return self.n_entries * self.n_features

    @property
    def n_output_dims(self) -> int:
        return self.n_levels * self.n_features


@dataclass
class FieldSpec:
    """Hash grid + geometry MLP (SDF field or RadF's Geo_enc)."""
    grid: GridSpec
    bound_min: Sequence[float]
    bound_max: Sequence[float]
    dims: Sequence[int]                 # [3 + 2L, 64, ..., k_geo + 1]
    rescale: float = 1.0
    softplus_beta: float = 100.0
    softplus_threshold: float = 20.0
    sdf_sign: float = 1.0               # +1 inside, -1 otherwise (models/SDF.py:62-71)
    scale_mlp: float = 1.0

    @property
    def n_layers(self) -> int:
        return len(self.dims) - 1

    @property
    def dout(self) -> int:
        return int(self.dims[-1])

    @property
    def theta_size(self) -> int:
        return sum(self.dims[l] * self.dims[l + 1] + self.dims[l + 1] for l in range(self.n_layers))

    def c_field(self, lib: _C.Lib, table: torch.Tensor, theta: Optional[torch.Tensor], image: Optional[torch.Tensor] = None) -> _C.Field:
        f = _C.Field()
        f.table = lib.ptr(table)
        f.theta = lib.ptr(theta) if theta is not None else None
        f.tc_image = lib.ptr(image) if image is not None else None
        f.n_levels = self.grid.n_levels
        for i, lv in enumerate(self.grid.levels):
            f.levels[i] = lv
        for d in range(3):
            f.bound_min[d] = float(self.bound_min[d])
            f.bound_max[d] = float(self.bound_max[d])
        f.rescale = float(self.rescale)
        f.n_layers = self.n_layers
        for i, d in enumerate(self.dims):
            f.dims[i] = int(d)
        f.softplus_beta = float(self.softplus_beta)
        f.softplus_threshold = float(self.softplus_threshold)
        f.sdf_sign = float(self.sdf_sign)
        f.scale_mlp = float(self.scale_mlp)
        return f


@dataclass
class RadSpec:
    """Radiance decoder input layout (models/Renderer.py:75; models/RadF.py:52-56)."""
    n_freq: int = 4
    k_geo: int = 16
    k_geo2: int = 0

    @property
    def in_dim(self) -> int:
        return 3 + 3 + 3 + 6 * self.n_freq + self.k_geo + self.k_geo2


def pack_theta(layers: List[Tuple[torch.Tensor, torch.Tensor]]) -> torch.Tensor:
    """[(W [out,in], b [out])...] -> flat theta: for l: W_l^T row-major, then b_l (differentiable)."""
    parts = []
    for W, b in layers:
        parts.append(W.t().reshape(-1))
        parts.append(b.reshape(-1))
    return torch.cat(parts)


def compose_affine(layers: List[Tuple[torch.Tensor, torch.Tensor]]) -> Tuple[torch.Tensor, torch.Tensor]:
    """The reference's radiance decoder applies no hidden activation (models/base.py:230,257), i.e. it is
    one affine map: W_eff = W_n ... W_1, b_eff likewise (differentiable w.r.t. every layer)."""
    W, b = layers[0]
    for Wn, bn in layers[1:]:
        b = torch.mv(Wn, b) + bn
        W = Wn @ W
    return W.contiguous(), b.contiguous()


# ------------------------------------------------------------------------- launch log
class KernelLog:
    """Counts the launches of our kernels and, when ``timing`` is on, brackets each launch with CUDA events on the
    launching stream (used by bench.py for ``gpu_launches`` and the roofline of the dominant kernel).  ``units`` = the
    samples / rays that launch processes, so per-launch algorithmic bytes can be attached to per-launch durations."""

    def __init__(self):
        self.counts = {}
        self.timing = False
        self.events = []        # (name, start, end, units)

    def reset(self):
        self.counts = {}
        self.events = []

    def total(self):
        return sum(self.counts.values())

    def begin(self, name, n=1):
        self.counts[name] = self.counts.get(name, 0) + n
        if self.timing:
            ev = torch.cuda.Event(enable_timing=True)
            ev.record()
            return ev
        return None

    def end(self, name, start, units=None):
        if start is not None:
            ev = torch.cuda.Event(enable_timing=True)
            ev.record()
            self.events.append((name, start, ev, units))

    def durations_ms(self):
        """name -> list of per-launch durations (call after torch.cuda.synchronize())."""
        out = {}
        for name, a, b, _ in self.events:
            out.setdefault(name, []).append(a.elapsed_time(b))
        return out

    def launches(self):
        """[(name, ms, units)] in launch order (call after torch.cuda.synchronize())."""
        return [(name, a.elapsed_time(b), units) for name, a, b, units in self.events]


KLOG = KernelLog()


def _call(lib, name, fn, *args, launches=1, units=None):
    ev = KLOG.begin(name, launches)
    lib.check(fn(*args))
    KLOG.end(name, ev, units)


# ------------------------------------------------------------------------- raw calls
def _points(lib: _C.Lib, xyz=None, center=None, ray=None, t=None, t_offset=0, n_per_ray=None):
    p = _C.Points()
    if xyz is not None:
        p.xyz = lib.ptr(xyz)
        p.n = xyz.numel() // 3
        if ray is not None:                      # explicit points that still belong to rays (radiance needs the direction)
            p.ray = lib.ptr(ray)
            p.center = lib.ptr(center) if center is not None else None
            p.n_rays = ray.numel() // 3
            p.n_per_ray = int(n_per_ray if n_per_ray is not None else (p.n // max(p.n_rays, 1)))
    else:
        p.center, p.ray, p.t = lib.ptr(center), lib.ptr(ray), lib.ptr(t)
        p.n_rays = center.numel() // 3
        npr = int(n_per_ray if n_per_ray is not None else t.shape[-1] - t_offset)
        p.n_per_ray = npr
        p.t_stride = int(t.shape[-1])
        p.t_offset = int(t_offset)
        p.n = p.n_rays * npr
    return p


def _rad(lib: _C.Lib, rs: RadSpec, w_eff, b_eff, geo2):
    r = _C.Radiance()
    r.w_eff, r.b_eff = lib.ptr(w_eff), lib.ptr(b_eff)
    r.in_dim, r.n_freq, r.k_geo, r.k_geo2 = rs.in_dim, rs.n_freq, rs.k_geo, rs.k_geo2
    r.geo2 = lib.ptr(geo2) if geo2 is not None else None
    return r


def field_prepare_raw(lib, spec: FieldSpec, table, theta, rad: Optional[_C.Radiance]):
    """theta (+ W_eff) -> tensor-core operand image (hi/lo TF32, kernel shared-memory layout); reuse it for every launch
    that evaluates the same parameters (sampler rounds, render forward, sphere tracing)."""
    f = spec.c_field(lib, table, theta)
    n = int(lib.dll.ls2fm_field_image_floats(f, rad))
    image = torch.empty(n, device=table.device)
    _call(lib, "field_prepare", lib.dll.ls2fm_field_prepare, f, rad, lib.ptr(image), lib.stream())
    return image


def field_forward_raw(lib, spec: FieldSpec, table, theta, pts: _C.Points, rad: Optional[_C.Radiance],
                      want_y=False, want_sdf=True, want_nrm=False, want_rgb=False, image=None, simt=False):
    n, dev = int(pts.n), table.device
    y = torch.empty(n, spec.dout, device=dev) if want_y else None
    sdf = torch.empty(n, device=dev) if want_sdf else None
    nrm = torch.empty(n, 3, device=dev) if want_nrm else None
    rgb = torch.empty(n, 3, device=dev) if want_rgb else None
    f = spec.c_field(lib, table, theta, image)
    if FORWARD_WS and rad is None and not want_nrm and not want_rgb and not simt:      # experimental warp-specialised kernel (tests only)
        _call(lib, "field_forward_ws", lib.dll.ls2fm_field_forward_ws, f, pts, lib.ptr(y), lib.ptr(sdf), lib.stream())
        return y, sdf, nrm, rgb
    _call(lib, "field_forward_simt" if simt else "field_forward", lib.dll.ls2fm_field_forward_simt if simt else lib.dll.ls2fm_field_forward, f, pts, rad, lib.ptr(y), lib.ptr(sdf), lib.ptr(nrm), lib.ptr(rgb), lib.stream(), units=n)
    return y, sdf, nrm, rgb


def field_backward_raw(lib, spec: FieldSpec, table, theta, pts: _C.Points, rad: Optional[_C.Radiance],
                       g_y, g_sdf, g_nrm, g_rgb, saved_nrm, saved_rgb,
                       d_table, d_theta, d_w_eff=None, d_b_eff=None, d_geo2=None, image=None, mode="auto",
                       d_xyz=None, d_center=None, d_ray=None, d_t=None):
    """image: the operand image of field_prepare_raw for the SAME theta -- with it large launches run the tensor-core kernel.
    mode: "auto" (library dispatch) | "simt" (exact fp32 kernel) | "tc" (tensor-core kernel or an error).
    d_xyz (explicit points, written) / d_center, d_ray (+=), d_t (written) (ray mode): gradients w.r.t. the sample positions."""
    f = spec.c_field(lib, table, theta, image)
    ig = None
    if d_xyz is not None or d_center is not None or d_ray is not None or d_t is not None:
        ig = _C.InputGrads()
        ig.d_xyz, ig.d_center, ig.d_ray, ig.d_t = lib.ptr(d_xyz), lib.ptr(d_center), lib.ptr(d_ray), lib.ptr(d_t)
        if image is not None and mode != "simt" and (mode == "tc" or int(pts.n) >= TC_BACKWARD_MIN_SAMPLES):
            # scratch for the tensor-core route (encoding adjoints parked per sample, finished by ls_field_posgrad_kernel)
            ws = torch.empty(int(lib.dll.ls2fm_field_backward_workspace_floats(int(pts.n))), device=table.device)
            ig.workspace = lib.ptr(ws)
    if mode == "tc" and image is None:
        mode = "auto"       # no operand image: the tensor-core kernel has no weights to stream
    name = {"auto": "field_backward", "simt": "field_backward_simt", "tc": "field_backward_tc"}[mode]
    _call(lib, name, getattr(lib.dll, "ls2fm_" + name), f, pts, rad, lib.ptr(g_y), lib.ptr(g_sdf), lib.ptr(g_nrm), lib.ptr(g_rgb), lib.ptr(saved_nrm), lib.ptr(saved_rgb),
        lib.ptr(d_table), lib.ptr(d_theta), lib.ptr(d_w_eff), lib.ptr(d_b_eff), lib.ptr(d_geo2), ig, lib.stream(), units=int(pts.n))


def grid_encode_raw(lib, grid: GridSpec, table, u, want_idx=False):
    m = u.numel() // 3
    spec = FieldSpec(grid, (0, 0, 0), (1, 1, 1), [3 + 2 * grid.n_levels, 64, 1])
    f = spec.c_field(lib, table, None)
    enc = torch.empty(m, grid.n_output_dims, device=u.device)
    idx = torch.empty(m, grid.n_levels, 8, dtype=torch.int32, device=u.device) if want_idx else None
    _call(lib, "grid_encode", lib.dll.ls2fm_grid_encode, f, lib.ptr(u), m, lib.ptr(enc), lib.ptr(idx, torch.int32), lib.stream())
    return enc, idx


def grid_encode_backward_raw(lib, grid: GridSpec, table, u, g_enc, d_table, d_u=None):
    m = u.numel() // 3
    spec = FieldSpec(grid, (0, 0, 0), (1, 1, 1), [3 + 2 * grid.n_levels, 64, 1])
    f = spec.c_field(lib, table, None)
    _call(lib, "grid_encode_backward", lib.dll.ls2fm_grid_encode_backward, f, lib.ptr(u), m, lib.ptr(g_enc), lib.ptr(d_table), lib.ptr(d_u), lib.stream())


def ray_aabb_raw(lib, rays_o, rays_d, center, half_size):
    m = rays_o.numel() // 3
    hits = torch.empty(m, 2, device=rays_o.device)
    cnt = torch.empty(m, dtype=torch.int32, device=rays_o.device)
    _call(lib, "ray_aabb", lib.dll.ls2fm_ray_aabb, lib.ptr(rays_o), lib.ptr(rays_d), m, _C.f3(center), _C.f3(half_size),
                                     lib.ptr(hits), lib.ptr(cnt, torch.int32), lib.stream())
    return hits, cnt


def sample_uniform_raw(lib, center, ray, n_samples, bound_min, bound_max):
    r = center.numel() // 3
    t = torch.empty(r, n_samples, device=center.device)
    hits = torch.empty(r, 2, device=center.device)
    _call(lib, "sample_uniform", lib.dll.ls2fm_sample_uniform, lib.ptr(center), lib.ptr(ray), r, n_samples, _C.f3(bound_min), _C.f3(bound_max),
                                           lib.ptr(t), lib.ptr(hits), lib.stream())
    return t, hits


def composite_forward_raw(lib, ray, t, sdf, rgbs, nrm, beta_param, beta_speed, bgcolor):
    r, n = t.shape
    dev = t.device
    rgb = torch.empty(r, 3, device=dev) if rgbs is not None else None
    depth = torch.empty(r, device=dev)
    normal = torch.empty(r, 3, device=dev) if nrm is not None else None
    opacity = torch.empty(r, device=dev)
    _call(lib, "composite_forward", lib.dll.ls2fm_composite_forward, lib.ptr(ray), lib.ptr(t), lib.ptr(sdf), lib.ptr(rgbs), lib.ptr(nrm), lib.ptr(beta_param), float(beta_speed),
        _C.f3(bgcolor), r, n, lib.ptr(rgb), lib.ptr(depth), lib.ptr(normal), lib.ptr(opacity), lib.stream())
    return rgb, depth, normal, opacity


def composite_backward_raw(lib, ray, t, sdf, rgbs, nrm, beta_param, beta_speed, bgcolor, g_rgb, g_depth, g_normal,
                           want_d_ray=False, want_d_t=False, d_beta=None):
    r, n = t.shape
    dev = t.device
    d_sdf = torch.empty(r, n, device=dev)
    d_rgbs = torch.empty(r, n, 3, device=dev) if rgbs is not None else None
    d_nrm = torch.empty(r, n, 3, device=dev) if nrm is not None else None
    if d_beta is None:          # (else: the caller's accumulator, e.g. the gradient bucket's slot of SDF.beta -- the kernel adds)
        d_beta = torch.zeros(1, device=dev)
    d_ray = torch.zeros(r, 3, device=dev) if want_d_ray else None
    d_t = torch.empty(r, n, device=dev) if want_d_t else None
    _call(lib, "composite_backward", lib.dll.ls2fm_composite_backward, lib.ptr(ray), lib.ptr(t), lib.ptr(sdf), lib.ptr(rgbs), lib.ptr(nrm), lib.ptr(beta_param), float(beta_speed),
        _C.f3(bgcolor), r, n, lib.ptr(g_rgb), lib.ptr(g_depth), lib.ptr(g_normal),
        lib.ptr(d_sdf), lib.ptr(d_rgbs), lib.ptr(d_nrm), lib.ptr(d_beta), lib.ptr(d_ray), lib.ptr(d_t), lib.stream())
    return d_sdf, d_rgbs, d_nrm, d_beta, d_ray, d_t


# ------------------------------------------------------------------------- autograd
TC_BACKWARD_MIN_SAMPLES = 8192     # LS_BT_MIN_SAMPLES of the library's automatic dispatch
FORWARD_WS = False         # tests set this to route values-only evaluations through the experimental warp-specialised kernel
BACKWARD_MODE = "auto"     # tests set "simt" / "tc" to cross-check the tensor-core backward kernel against the fp32-SIMT one

def _c(t):
    return None if t is None else t.contiguous()


def grad_sink(param):
    """The buffer the fused backward may accumulate into directly for this parameter, or None.

    ``parallel.GradBucket`` (or any caller that owns ``param.grad`` as a persistent, pre-zeroed buffer and only ever calls
    ``loss.backward()``) sets ``param._ls2fm_grad_sink = param.grad``: the table-gradient scatter then lands in the bucket itself
    instead of a fresh zero-filled 49 MB tensor that autograd adds to ``.grad`` afterwards (two extra passes over the table
    per backward).  Not valid under ``torch.autograd.grad`` -- which is why it is opt-in."""
    sink = getattr(param, "_ls2fm_grad_sink", None)
    if sink is not None and (sink.shape != param.shape or sink.device != param.device or not sink.is_contiguous()
                             or sink.data_ptr() % 8 or sink.dtype != torch.float32):
        return None
    return sink


class FieldEval(torch.autograd.Function):
    """Fused hash grid + geometry MLP (+ analytic normals, + radiance).

    Differentiable w.r.t. ``table``, ``theta``, ``w_eff``, ``b_eff``, ``geo2`` -- including the second-order path through the
    normals (the reference's SDF.gradient uses create_graph=True, models/SDF.py:107-113) -- and w.r.t. the sample positions:
    ``xyz`` (explicit points) or ``center`` / ``ray`` / ``t`` (ray mode, x = center + ray * t), as tiny-cuda-nn's encoding is in
    the reference (pipelines/BA.py:123-125 feeds get_surface_pts' output back into infer_sdf; SDF.gradient is differentiated
    w.r.t. p).  The position gradient itself is first-order (no double backward through it).

    forward(spec, rad_spec, table, theta, w_eff, b_eff, geo2, xyz, center, ray, t, t_offset, n_per_ray,
            want_y, want_nrm, image=None, sdf_x_detached=False)
        -> (sdf [n], y [n,dout] | empty, nrm [n,3] | empty, rgb [n,3] | empty)
    sdf_x_detached: the sdf output is treated as a function of the DETACHED positions (get_surface_pts evaluates
    ``infer_sdf(pts.detach())`` next to ``gradient(pts)``, models/SDF.py:96-97) while sharing one launch with the normals.
    """

    @staticmethod
    def forward(ctx, spec, rad_spec, table, theta, w_eff, b_eff, geo2, xyz, center, ray, t, t_offset, n_per_ray,
                want_y, want_nrm, image=None, sdf_x_detached=False):
        lib = _C.get()
        ctx.table_sink = grad_sink(table)
        table, theta = table.detach().contiguous(), theta.detach().contiguous()
        xyz, center, ray, t = (_c(v.detach()) if v is not None else None for v in (xyz, center, ray, t))
        with_rad = w_eff is not None
        if with_rad:
            w_eff, b_eff = w_eff.detach().contiguous(), b_eff.detach().contiguous()
            geo2 = _c(geo2.detach()) if geo2 is not None else None
        pts = _points(lib, xyz, center, ray, t, t_offset, n_per_ray)
        rad = _rad(lib, rad_spec, w_eff, b_eff, geo2) if with_rad else None
        y, sdf, nrm, rgb = field_forward_raw(lib, spec, table, theta, pts, rad, want_y=want_y, want_sdf=True,
                                             want_nrm=want_nrm or with_rad, want_rgb=with_rad, image=image)
        ctx.spec, ctx.rad_spec = spec, rad_spec
        ctx.pt_args = (t_offset, n_per_ray)
        ctx.with_rad, ctx.want_y, ctx.want_nrm = with_rad, want_y, want_nrm
        ctx.image = image
        ctx.sdf_x_detached = bool(sdf_x_detached)
        keep_nrm = with_rad or (sdf_x_detached and nrm is not None)
        ctx.save_for_backward(table, theta, w_eff, b_eff, geo2, xyz, center, ray, t, nrm if keep_nrm else None, rgb)
        out_y = y if want_y else table.new_empty(0)
        out_nrm = nrm if (want_nrm or with_rad) else table.new_empty(0)
        out_rgb = rgb if with_rad else table.new_empty(0)
        ctx.mark_non_differentiable(*[o for o in (out_y, out_nrm, out_rgb) if o.numel() == 0])
        return sdf, out_y, out_nrm, out_rgb

    @staticmethod
    def backward(ctx, g_sdf, g_y, g_nrm, g_rgb):
        lib = _C.get()
        table, theta, w_eff, b_eff, geo2, xyz, center, ray, t, s_nrm, s_rgb = ctx.saved_tensors
        spec, rs = ctx.spec, ctx.rad_spec
        pts = _points(lib, xyz, center, ray, t, *ctx.pt_args)
        with_rad = ctx.with_rad and g_rgb is not None
        rad = _rad(lib, rs, w_eff, b_eff, geo2) if with_rad else None
        g_y = _c(g_y) if (ctx.want_y and g_y is not None) else None
        g_nrm = _c(g_nrm) if ((ctx.want_nrm or ctx.with_rad) and g_nrm is not None) else None
        g_rgb = _c(g_rgb) if with_rad else None
        g_sdf = _c(g_sdf) if g_sdf is not None else None
        need = ctx.needs_input_grad
        sink = ctx.table_sink if need[2] else None
        d_table = sink if sink is not None else (torch.zeros_like(table) if need[2] else None)
        # one zero-fill for the three small accumulators (the kernels add into them with atomics)
        n_t = theta.numel() if need[3] else 0
        n_w, n_b = (w_eff.numel(), b_eff.numel()) if with_rad else (0, 0)
        small = torch.zeros(n_t + n_w + n_b, device=table.device)
        d_theta = small[:n_t] if need[3] else None
        d_w = small[n_t:n_t + n_w].view_as(w_eff) if with_rad else None
        d_b = small[n_t + n_w:] if with_rad else None
        d_geo2 = torch.empty_like(geo2) if (with_rad and geo2 is not None) else None
        d_xyz = d_center = d_ray = d_t = None
        if xyz is not None:
            if need[7]:
                d_xyz = torch.empty_like(xyz)
            if need[9] and ray is not None and with_rad:
                d_ray = torch.zeros_like(ray)
        else:
            if need[8]:
                d_center = torch.zeros_like(center)
            if need[9]:
                d_ray = torch.zeros_like(ray)
            if need[10]:
                t_off, npr = ctx.pt_args
                if t_off != 0 or (npr is not None and npr != t.shape[-1]):
                    raise NotImplementedError("gradient w.r.t. a column slice of t")
                d_t = torch.empty_like(t)
        field_backward_raw(lib, spec, table, theta, pts, rad, g_y, g_sdf, g_nrm, g_rgb, s_nrm if with_rad else None, s_rgb,
                           d_table, d_theta, d_w, d_b, d_geo2, image=ctx.image, mode=BACKWARD_MODE,
                           d_xyz=d_xyz, d_center=d_center, d_ray=d_ray, d_t=d_t)
        if d_xyz is not None and ctx.sdf_x_detached and g_sdf is not None:
            d_xyz = d_xyz - g_sdf.reshape(-1, 1) * s_nrm          # d sdf / d x == the forward normal
        return (None, None, None if sink is not None else d_table, d_theta, d_w, d_b, d_geo2, d_xyz, d_center, d_ray, d_t,
                None, None, None, None, None, None)


class Composite(torch.autograd.Function):
    """Laplace-CDF density + alpha compositing + background fill (models/Renderer.py:33-49,80-107).

    forward(ray [R,3], t [R,N], sdf [R,N], rgbs [R,N,3] | None, nrm [R,N,3] | None, beta_param [1], beta_speed, bgcolor)
        -> rgb [R,3], depth [R], normal [R,3], opacity [R]
    Differentiable w.r.t. sdf, rgbs, nrm, beta_param, ray (interval lengths |ray| dt) and t.
    """

    @staticmethod
    def forward(ctx, ray, t, sdf, rgbs, nrm, beta_param, beta_speed, bgcolor):
        lib = _C.get()
        ray, t, sdf = _c(ray.detach()), _c(t.detach()), _c(sdf.detach())
        rgbs = _c(rgbs.detach()) if rgbs is not None else None
        nrm = _c(nrm.detach()) if nrm is not None else None
        ctx.beta_sink = grad_sink(beta_param)
        beta_param = beta_param.detach().contiguous()
        rgb, depth, normal, opacity = composite_forward_raw(lib, ray, t, sdf, rgbs, nrm, beta_param, beta_speed, bgcolor)
        ctx.save_for_backward(ray, t, sdf, rgbs, nrm, beta_param)
        ctx.misc = (beta_speed, tuple(bgcolor))
        ctx.mark_non_differentiable(opacity)
        if rgb is None:
            rgb = t.new_empty(0)
            ctx.mark_non_differentiable(rgb)
        if normal is None:
            normal = t.new_empty(0)
            ctx.mark_non_differentiable(normal)
        return rgb, depth, normal, opacity

    @staticmethod
    def backward(ctx, g_rgb, g_depth, g_normal, _g_op):
        lib = _C.get()
        ray, t, sdf, rgbs, nrm, beta_param = ctx.saved_tensors
        beta_speed, bg = ctx.misc
        want_ray, want_t = ctx.needs_input_grad[0], ctx.needs_input_grad[1]
        sink = ctx.beta_sink if ctx.needs_input_grad[5] else None
        d_sdf, d_rgbs, d_nrm, d_beta, d_ray, d_t = composite_backward_raw(
            lib, ray, t, sdf, rgbs, nrm, beta_param, beta_speed, bg, _c(g_rgb) if rgbs is not None else None, _c(g_depth),
            _c(g_normal) if nrm is not None else None, want_ray, want_t, d_beta=sink)
        return d_ray, d_t, d_sdf, d_rgbs, d_nrm, None if sink is not None else d_beta.view_as(beta_param), None, None


def _aabb_vjp(c, r, hits, g, n_samples, bmin, bmax):
    """VJP of the slab test the reference's RayAABBIntersector lacks (utils/custom_functions.py:10-31; SURVEY 8a defect iii), one
    launch (ls2fm_ray_aabb_backward): g [R, n_samples] on the uniform depths, or [R,2] on (t_near, t_far) when n_samples == 0."""
    lib = _C.get()
    d_c, d_r = torch.empty_like(c), torch.empty_like(r)
    _call(lib, "ray_aabb_backward", lib.dll.ls2fm_ray_aabb_backward, lib.ptr(c), lib.ptr(r), c.shape[0], int(n_samples), _C.f3(bmin), _C.f3(bmax),
          lib.ptr(hits), lib.ptr(g.contiguous().float()), lib.ptr(d_c), lib.ptr(d_r), lib.stream())
    return d_c, d_r


class RayAABB(torch.autograd.Function):
    """``RayAABBIntersector.apply`` for one box with the analytic backward: forward(rays_o [M,3], rays_d [M,3], bound_min, bound_max)
    -> hits_t [M,2] = (t_near, t_far) or (-1, -1)."""

    @staticmethod
    def forward(ctx, rays_o, rays_d, bound_min, bound_max):
        lib = _C.get()
        o, d = _c(rays_o.detach().float()), _c(rays_d.detach().float())
        center = [(float(a) + float(b)) / 2 for a, b in zip(bound_min, bound_max)]
        half = [(float(b) - float(a)) / 2 for a, b in zip(bound_min, bound_max)]
        hits, _ = ray_aabb_raw(lib, o, d, center, half)
        ctx.save_for_backward(o, d, hits)
        ctx.misc = (tuple(float(x) for x in bound_min), tuple(float(x) for x in bound_max))
        return hits

    @staticmethod
    def backward(ctx, g_hits):
        o, d, hits = ctx.saved_tensors
        d_o, d_d = _aabb_vjp(o, d, hits, g_hits, 0, *ctx.misc)
        return d_o, d_d, None, None


class UniformDepths(torch.autograd.Function):
    """Ray / AABB slab test + Renderer.sample_depth (models/Renderer.py:118-127,178-185) as ONE launch, with the analytic VJP the
    reference's RayAABBIntersector lacks (``_aabb_vjp``).

    forward(center [R,3], ray [R,3], n_samples, bound_min, bound_max) -> t [R,N]"""

    @staticmethod
    def forward(ctx, center, ray, n_samples, bound_min, bound_max):
        lib = _C.get()
        c, r = _c(center.detach()), _c(ray.detach())
        t, hits = sample_uniform_raw(lib, c, r, n_samples, bound_min, bound_max)
        ctx.save_for_backward(c, r, hits)
        ctx.misc = (n_samples, tuple(bound_min), tuple(bound_max))
        return t

    @staticmethod
    def backward(ctx, g_t):
        c, r, hits = ctx.saved_tensors
        n, bmin, bmax = ctx.misc
        d_c, d_r = _aabb_vjp(c, r, hits, g_t, n, bmin, bmax)
        return d_c, d_r, None, None, None


def grid_encode_tangent_raw(lib, grid: GridSpec, table, u, v, g_enc=None, want_t_enc=True, d_table=None, d_u2=None):
    m = u.numel() // 3
    spec = FieldSpec(grid, (0, 0, 0), (1, 1, 1), [3 + 2 * grid.n_levels, 64, 1])
    f = spec.c_field(lib, table, None)
    t_enc = torch.empty(m, grid.n_output_dims, device=u.device) if want_t_enc else None
    _call(lib, "grid_encode_tangent", lib.dll.ls2fm_grid_encode_tangent, f, lib.ptr(u), m, lib.ptr(v), lib.ptr(g_enc), lib.ptr(t_enc),
          lib.ptr(d_table), lib.ptr(d_u2), lib.stream())
    return t_enc


class GridEncode(torch.autograd.Function):
    """``tcnn.Encoding`` replacement (models/base.py:17,37): u [M,3] -> [M, L*F]; gradients to the table and to u, and -- like
    tcnn's -- differentiable a second time (``GridEncodeBackward``): the reference's ``SDF.gradient`` back-propagates THROUGH the
    input gradient of the encoding (create_graph=True, models/SDF.py:102-114)."""

    @staticmethod
    def forward(ctx, grid, table, u):
        lib = _C.get()
        enc, _ = grid_encode_raw(lib, grid, table.detach().contiguous(), u.detach().contiguous())
        ctx.grid = grid
        ctx.save_for_backward(table, u)
        return enc

    @staticmethod
    def backward(ctx, g_enc):
        table, u = ctx.saved_tensors
        d_table, d_u = GridEncodeBackward.apply(ctx.grid, table, u, g_enc, ctx.needs_input_grad[1], ctx.needs_input_grad[2])
        return None, d_table if ctx.needs_input_grad[1] else None, d_u if ctx.needs_input_grad[2] else None


class GridEncodeBackward(torch.autograd.Function):
    """(table, u, g_enc) -> (d_table, d_u) of GridEncode, itself differentiable:
       d(d_table)/d(g_enc) and d(d_u)/d(g_enc) : encode(gg_table, u) + J_enc(table, u) gg_u
       d(d_u)/d(table)                          : the tangent scatter along gg_u
       d(d_table)/d(u), d(d_u)/d(u)             : J_enc(gg_table, u)^T g_enc + the mixed second derivatives along gg_u."""

    @staticmethod
    def forward(ctx, grid, table, u, g_enc, want_table, want_u):
        lib = _C.get()
        t_c, u_c, g_c = table.detach().contiguous(), u.detach().contiguous(), g_enc.detach().contiguous()
        d_table = torch.zeros_like(t_c) if want_table else None
        d_u = torch.zeros_like(u_c) if want_u else None
        grid_encode_backward_raw(lib, grid, t_c, u_c, g_c, d_table, d_u)
        ctx.grid = grid
        ctx.save_for_backward(t_c, u_c, g_c)
        outs = (d_table if want_table else t_c.new_empty(0), d_u if want_u else t_c.new_empty(0))
        ctx.mark_non_differentiable(*[o for o in outs if o.numel() == 0])
        return outs

    @staticmethod
    def backward(ctx, gg_table, gg_u):
        lib = _C.get()
        table, u, g_enc = ctx.saved_tensors
        grid = ctx.grid
        need_t, need_u, need_g = ctx.needs_input_grad[1], ctx.needs_input_grad[2], ctx.needs_input_grad[3]
        have_t = gg_table is not None and gg_table.numel() > 0
        have_u = gg_u is not None and gg_u.numel() > 0
        d_g = d_t = d_uu = None
        if need_g:
            d_g = torch.zeros_like(g_enc)
            if have_t:
                d_g = d_g + grid_encode_raw(lib, grid, gg_table.detach().contiguous(), u)[0]
            if have_u:
                d_g = d_g + grid_encode_tangent_raw(lib, grid, table, u, gg_u.detach().contiguous())
        if need_t and have_u:
            d_t = torch.zeros_like(table)
            grid_encode_tangent_raw(lib, grid, table, u, gg_u.detach().contiguous(), g_enc, want_t_enc=False, d_table=d_t)
        if need_u:
            d_uu = torch.zeros_like(u)
            if have_t:
                grid_encode_backward_raw(lib, grid, gg_table.detach().contiguous(), u, g_enc, None, d_uu)
            if have_u:
                grid_encode_tangent_raw(lib, grid, table, u, gg_u.detach().contiguous(), g_enc, want_t_enc=False, d_u2=d_uu)
        return None, d_t, d_uu, d_g, None, None


def sample_error_bounded_raw(lib, spec: FieldSpec, table, theta, beta_param, center, ray, n_samples, n_final,
                             max_upsample_iter, max_bisection_itr, eps, beta_speed, image=None):
    """Renderer.volsdf_sampling's error-bounded branch -> (t [R, N+Nf], beta_plus [R], iters [R])."""
    r = center.numel() // 3
    dev = center.device
    cfg = _C.SamplerCfg(int(n_samples), int(n_final), int(max_upsample_iter), int(max_bisection_itr), float(eps), float(beta_speed))
    nbytes = int(lib.dll.ls2fm_sampler_workspace_bytes(cfg, r))
    if nbytes < 0:
        raise RuntimeError("bad sampler configuration")
    ws = torch.empty(max(nbytes, 16) // 4 + 4, dtype=torch.float32, device=dev)
    t = torch.empty(r, n_samples + n_final, device=dev)
    beta_plus = torch.empty(r, device=dev)
    iters = torch.empty(r, device=dev)
    f = spec.c_field(lib, table, theta, image)
    _call(lib, "sample_error_bounded", lib.dll.ls2fm_sample_error_bounded, f, lib.ptr(beta_param), cfg, lib.ptr(center), lib.ptr(ray),
          r, lib.ptr(ws), lib.ptr(t), lib.ptr(beta_plus), lib.ptr(iters), lib.stream(),
          launches=2 + 2 * (int(max_upsample_iter) + 1))     # init + (field forward + round) per round + finalize
    return t, beta_plus, iters


def sphere_trace_raw(lib, spec: FieldSpec, table, theta, ray0, ray_dir, sdf_threshold, iters_max):
    """No-grad march of SDF.sphere_tracing -> (track [M,iters_max,3], n_unfinished [iters_max+1] int32,
    t_near [M], t_far [M], acc_end [iters_max+1, M])."""
    m = ray0.numel() // 3
    dev = ray0.device
    track = torch.empty(m, iters_max, 3, device=dev)
    cnt = torch.empty(iters_max + 1, dtype=torch.int32, device=dev)
    t_near, t_far = torch.empty(m, device=dev), torch.empty(m, device=dev)
    acc_end = torch.empty(iters_max + 1, m, device=dev)
    f = spec.c_field(lib, table, theta)
    _call(lib, "sphere_trace", lib.dll.ls2fm_sphere_trace, f, lib.ptr(ray0), lib.ptr(ray_dir), m, float(sdf_threshold), int(iters_max),
          lib.ptr(track), lib.ptr(cnt, torch.int32), lib.ptr(t_near), lib.ptr(t_far), lib.ptr(acc_end), lib.stream())
    return track, cnt, t_near, t_far, acc_end


def grid_points_raw(lib, n, step, origin, begin, count, device):
    """utils/util.py:392-411 query points [count,3] for flat grid indices [begin, begin+count) (float64 arithmetic on the device)."""
    xyz = torch.empty(count, 3, device=device)
    o = (_C.C.c_double * 3)(float(origin[0]), float(origin[1]), float(origin[2]))
    _call(lib, "grid_points", lib.dll.ls2fm_grid_points, int(n), float(step), o, int(begin), int(count), lib.ptr(xyz), lib.stream())
    return xyz


class RenderLoss(torch.autograd.Function):
    """w_rgb * mean|rgb - gt| + w_eik * mean| ||normals|| - 1 | in one kernel (forward value and both gradients in the
    same pass); the rendering-loss tail of pipelines/rendering_refine.py:99-121 / BA.py:190-204.
    forward(rgb [...,3], gt [...,3], normals [...,3], w_rgb, w_eik) -> (loss, rgb_l1_mean, eikonal_mean)"""

    @staticmethod
    def forward(ctx, rgb, gt, normals, w_rgb, w_eik):
        lib = _C.get()
        rgb_c, gt_c, nrm_c = rgb.detach().contiguous(), gt.detach().contiguous(), normals.detach().contiguous()
        n_rays, n_samples = rgb_c.numel() // 3, nrm_c.numel() // 3
        sums = torch.empty(2, device=rgb.device)
        g_rgb = torch.empty_like(rgb_c)
        g_nrm = torch.empty_like(nrm_c)
        _call(lib, "render_loss", lib.dll.ls2fm_render_loss, lib.ptr(rgb_c), lib.ptr(gt_c), n_rays, lib.ptr(nrm_c), n_samples,
              float(w_rgb), float(w_eik), lib.ptr(sums), lib.ptr(g_rgb), lib.ptr(g_nrm), lib.stream())
        ctx.save_for_backward(g_rgb, g_nrm)
        l1 = sums[0] / max(3 * n_rays, 1)
        eik = sums[1] / max(n_samples, 1)
        out = w_rgb * l1 + w_eik * eik
        ctx.mark_non_differentiable(l1, eik)
        return out, l1, eik

    @staticmethod
    def backward(ctx, g_out, _g1, _g2):
        g_rgb, g_nrm = ctx.saved_tensors
        return g_rgb * g_out, None, g_nrm * g_out, None, None


class RenderTail(torch.autograd.Function):
    """The loss tail of ``CameraSet.render`` + the stage's ``compute_loss`` (pipelines/Camera.py:506-537, BA.py:190-204,
    rendering_refine.py:99-107) in two launches: background / finish masks, rgb L1, PSNR, smooth-L1 depth consistency between the
    sphere-traced and the volume-rendered depth, (masked) eikonal -- values and every gradient in the same passes.

    forward(rgb [...,3], gt [...,3], normals [..., N, 3] | None, depth_mlp [...,1] | None, d_points [...] | None,
            mask_finish [...] bool | None, w_rgb, w_eik, w_dc, eik_masked)
        -> (total, rgb_loss, eikonal, DC_loss, PSNR, mask_bg [R] bool, mask_finish [R] bool);  only ``total`` carries gradient:
           total = w_rgb * rgb_loss + w_eik * eikonal + w_dc * DC_loss."""

    @staticmethod
    def forward(ctx, rgb, gt, normals, depth_mlp, d_points, mask_finish, w_rgb, w_eik, w_dc, eik_masked):
        lib = _C.get()
        dev = rgb.device
        rgb_c, gt_c = rgb.detach().reshape(-1, 3).contiguous().float(), gt.detach().reshape(-1, 3).contiguous().float()
        R = rgb_c.shape[0]
        nrm_c = normals.detach().contiguous().float() if normals is not None else None
        npr = (nrm_c.numel() // 3) // max(R, 1) if nrm_c is not None else 0
        dep_c = depth_mlp.detach().reshape(-1).contiguous().float() if depth_mlp is not None else None
        dpt_c = d_points.detach().reshape(-1).contiguous().float() if d_points is not None else None
        fin_c = mask_finish.detach().reshape(-1).to(torch.uint8).contiguous() if mask_finish is not None else None
        sums = torch.empty(8, device=dev)
        m_bg = torch.empty(R, dtype=torch.uint8, device=dev)
        m_fin = torch.empty(R, dtype=torch.uint8, device=dev)
        g_rgb = torch.empty_like(rgb_c)
        g_nrm = torch.empty_like(nrm_c) if nrm_c is not None else None
        g_dep = torch.empty_like(dep_c) if dep_c is not None else None
        g_dpt = torch.empty_like(dpt_c) if dpt_c is not None else None
        _call(lib, "render_tail", lib.dll.ls2fm_render_tail, lib.ptr(rgb_c), lib.ptr(gt_c), lib.ptr(dep_c), lib.ptr(dpt_c),
              lib.ptr(fin_c, torch.uint8), lib.ptr(nrm_c), R, int(npr), int(bool(eik_masked)), float(w_rgb), float(w_eik), float(w_dc),
              lib.ptr(sums), lib.ptr(m_bg, torch.uint8), lib.ptr(m_fin, torch.uint8), lib.ptr(g_rgb), lib.ptr(g_dep), lib.ptr(g_dpt),
              lib.ptr(g_nrm), lib.stream(), launches=2)
        ctx.save_for_backward(g_rgb, g_nrm, g_dep, g_dpt)
        ctx.shapes = (rgb.shape, normals.shape if normals is not None else None, depth_mlp.shape if depth_mlp is not None else None,
                      d_points.shape if d_points is not None else None)
        l1 = sums[0] / max(3 * R, 1)
        cnt_eik = sums[2] * npr if eik_masked else torch.full((), float(max(R * npr, 1)), device=dev)
        eik = sums[1] / cnt_eik if nrm_c is not None else torch.zeros((), device=dev)
        dc = torch.where(sums[5] > 0, sums[4] / sums[5].clamp_min(1.0), torch.zeros((), device=dev))
        psnr = -10.0 * torch.log10(sums[3] / (3.0 * sums[2]))
        total = w_rgb * l1 + w_eik * eik + w_dc * dc
        outs = (total, l1, eik, dc, psnr, m_bg.bool(), m_fin.bool())
        ctx.mark_non_differentiable(*outs[1:])
        return outs

    @staticmethod
    def backward(ctx, g_out, *_):
        g_rgb, g_nrm, g_dep, g_dpt = ctx.saved_tensors
        s_rgb, s_nrm, s_dep, s_dpt = ctx.shapes
        return (
            (g_rgb * g_out).view(s_rgb), None, (g_nrm * g_out).view(s_nrm) if g_nrm is not None else None,
            (g_dep * g_out).view(s_dep) if g_dep is not None else None, (g_dpt * g_out).view(s_dpt) if g_dpt is not None else None,
            None, None, None, None, None)


def _param_layers(lib, layers, grads=None):
    arr = (_C.ParamLayer * len(layers))()
    for i, (g, v, b) in enumerate(layers):
        arr[i].g, arr[i].v, arr[i].b = lib.ptr(g), lib.ptr(v), lib.ptr(b)
        arr[i].dout, arr[i].din = int(v.shape[0]), int(v.shape[1])
        if grads is not None:
            arr[i].dg, arr[i].dv, arr[i].db = lib.ptr(grads[i][0]), lib.ptr(grads[i][1]), lib.ptr(grads[i][2])
    return arr


class ParamPrep(torch.autograd.Function):
    """Weight norm of every layer + packing of the geometry MLP (theta) + composition of the radiance decoder
    (W_eff, b_eff) in ONE launch; its backward (to every weight_g / weight_v / bias) in one more.

    forward(n_geo, *tensors) with tensors = (g, v, b) per geometry layer followed by (g, v, b) of the 3 radiance layers
    (or nothing) -> (theta, w_eff, b_eff)   (w_eff / b_eff are empty tensors without a radiance decoder)."""

    @staticmethod
    def forward(ctx, n_geo, *tensors):
        lib = _C.get()
        sinks = [grad_sink(t) for t in tensors]
        ctx.sinks = sinks if all(s is not None for s in sinks) else None
        ts = [t.detach().contiguous() for t in tensors]
        geo = [tuple(ts[3 * i:3 * i + 3]) for i in range(n_geo)]
        rad = [tuple(ts[3 * (n_geo + i):3 * (n_geo + i) + 3]) for i in range((len(ts) - 3 * n_geo) // 3)]
        assert len(rad) in (0, 3)
        dev = ts[0].device
        n_theta = sum(v.shape[0] * v.shape[1] + v.shape[0] for _, v, _ in geo)
        theta = torch.empty(n_theta, device=dev)
        w_eff = torch.empty(3, rad[0][1].shape[1], device=dev) if rad else torch.empty(0, device=dev)
        b_eff = torch.empty(3, device=dev) if rad else torch.empty(0, device=dev)
        _call(lib, "params_forward", lib.dll.ls2fm_params_forward, _param_layers(lib, geo) if geo else None, n_geo,
              _param_layers(lib, rad) if rad else None, lib.ptr(theta), lib.ptr(w_eff), lib.ptr(b_eff), lib.stream())
        ctx.n_geo, ctx.n_rad = n_geo, len(rad)
        ctx.save_for_backward(*ts)
        return theta, w_eff, b_eff

    @staticmethod
    def backward(ctx, d_theta, d_w_eff, d_b_eff):
        lib = _C.get()
        ts = list(ctx.saved_tensors)
        n_geo, n_rad = ctx.n_geo, ctx.n_rad
        use_rad = n_rad == 3 and d_w_eff is not None and d_w_eff.numel() > 0
        # with gradient sinks on every parameter (parallel.GradBucket(direct=True)) the kernel adds straight into the .grad buffers
        direct = ctx.sinks is not None and all(ctx.needs_input_grad[1:]) and (n_rad == 0 or use_rad)
        grads = list(ctx.sinks) if direct else [torch.empty_like(t) for t in ts]
        geo = [tuple(ts[3 * i:3 * i + 3]) for i in range(n_geo)]
        rad = [tuple(ts[3 * (n_geo + i):3 * (n_geo + i) + 3]) for i in range(n_rad)]
        ggeo = [tuple(grads[3 * i:3 * i + 3]) for i in range(n_geo)]
        grad_ = [tuple(grads[3 * (n_geo + i):3 * (n_geo + i) + 3]) for i in range(n_rad)]
        if n_rad == 3 and not use_rad:
            for g in grads[3 * n_geo:]:
                g.zero_()
        if d_theta is None and n_geo:
            d_theta = torch.zeros(sum(v.shape[0] * v.shape[1] + v.shape[0] for _, v, _ in geo), device=ts[0].device)
        if use_rad and d_b_eff is None:
            d_b_eff = torch.zeros(3, device=ts[0].device)
        _call(lib, "params_backward", lib.dll.ls2fm_params_backward, _param_layers(lib, geo, ggeo) if geo else None, n_geo,
              _param_layers(lib, rad, grad_) if use_rad else None,
              lib.ptr(d_theta.contiguous()) if n_geo else None, lib.ptr(d_w_eff.contiguous()) if use_rad else None,
              lib.ptr(d_b_eff.contiguous()) if use_rad else None, 1 if direct else 0, lib.stream())
        if direct:
            return (None,) * (1 + len(ts))
        return (None, *grads)


class GenerateRays(torch.autograd.Function):
    """utils/camera.py:230-252 (get_center_and_ray) as one kernel: pose [B,3,4], kinv [B,3,3], xy [N,2] ->
    center [B,N,3], ray [B,N,3]; differentiable w.r.t. the pose."""

    @staticmethod
    def forward(ctx, pose, kinv, xy):
        lib = _C.get()
        pose_c, kinv_c, xy_c = pose.detach().contiguous().float(), kinv.detach().contiguous().float(), xy.detach().contiguous().float()
        B, N = pose_c.shape[0], xy_c.shape[0]
        if tuple(kinv_c.shape) != (B, 3, 3) or tuple(pose_c.shape) != (B, 3, 4):
            raise ValueError(f"GenerateRays: pose {tuple(pose_c.shape)} / kinv {tuple(kinv_c.shape)}: need [B,3,4] and [B,3,3]")
        center = torch.empty(B, N, 3, device=pose.device)
        ray = torch.empty(B, N, 3, device=pose.device)
        _call(lib, "generate_rays", lib.dll.ls2fm_generate_rays, lib.ptr(pose_c), lib.ptr(kinv_c), lib.ptr(xy_c), B, N,
              lib.ptr(center), lib.ptr(ray), lib.stream())
        ctx.save_for_backward(pose_c, kinv_c, xy_c)
        return center, ray

    @staticmethod
    def backward(ctx, g_center, g_ray):
        lib = _C.get()
        pose, kinv, xy = ctx.saved_tensors
        d_pose = torch.zeros_like(pose)
        _call(lib, "generate_rays_backward", lib.dll.ls2fm_generate_rays_backward, lib.ptr(pose), lib.ptr(kinv), lib.ptr(xy),
              pose.shape[0], xy.shape[0], lib.ptr(_c(g_center)), lib.ptr(_c(g_ray)), lib.ptr(d_pose), lib.stream())
        return d_pose, None, None


class Se3ToSE3(torch.autograd.Function):
    """utils/camera.py:85-96 (Lie.se3_to_SE3, Taylor coefficients of camera.py:119-142) as one kernel each way:
    wu [...,6] -> Rt [...,3,4]."""

    @staticmethod
    def forward(ctx, wu):
        lib = _C.get()
        w = wu.detach().reshape(-1, 6).contiguous().float()
        Rt = torch.empty(w.shape[0], 3, 4, device=w.device)
        _call(lib, "se3_to_SE3", lib.dll.ls2fm_se3_to_SE3, lib.ptr(w), w.shape[0], lib.ptr(Rt), lib.stream())
        ctx.save_for_backward(w)
        ctx.shape = wu.shape
        return Rt.view(*wu.shape[:-1], 3, 4)

    @staticmethod
    def backward(ctx, g):
        lib = _C.get()
        (w,) = ctx.saved_tensors
        d = torch.empty_like(w)
        _call(lib, "se3_to_SE3_backward", lib.dll.ls2fm_se3_to_SE3_backward, lib.ptr(w), w.shape[0], lib.ptr(g.reshape(-1, 12).contiguous().float()),
              lib.ptr(d), lib.stream())
        return d.view(ctx.shape)


class ReprojLoss(torch.autograd.Function):
    """The reprojection term of the BA "sfm" loop (pipelines/BA.py:126-141) in two launches, gradients included.
    forward(xyz [n,3], Rt [n,3,4], K [3,3], kypts [n,2], sdf [n] or [n,1], sdf_band, eps=1e-6)
        -> (loss, uv [n,2], mask_surf [n] bool, n_kept);  loss is 0 when no point lies in the band (BA.py:147-148)."""

    @staticmethod
    def forward(ctx, xyz, Rt, K, kypts, sdf, sdf_band, eps=1e-6):
        lib = _C.get()
        x, P = xyz.detach().reshape(-1, 3).contiguous().float(), Rt.detach().reshape(-1, 12).contiguous().float()
        Kc, kp = K.detach().reshape(9).contiguous().float(), kypts.detach().reshape(-1, 2).contiguous().float()
        s = sdf.detach().reshape(-1).contiguous().float()
        n = x.shape[0]
        dev = x.device
        sums = torch.empty(4, device=dev)
        uv = torch.empty(n, 2, device=dev)
        mask = torch.empty(n, dtype=torch.uint8, device=dev)
        g_x, g_P = torch.empty_like(x), torch.empty_like(P)
        _call(lib, "reproj_loss", lib.dll.ls2fm_reproj_loss, lib.ptr(x), lib.ptr(P), lib.ptr(Kc), lib.ptr(kp), lib.ptr(s), n, float(sdf_band),
              float(eps), lib.ptr(sums), lib.ptr(uv), lib.ptr(mask, torch.uint8), lib.ptr(g_x), lib.ptr(g_P), lib.stream(), launches=2)
        ctx.save_for_backward(g_x, g_P)
        ctx.shapes = (xyz.shape, Rt.shape)
        nk = sums[2]
        loss = torch.where(nk > 0, 0.5 * (sums[0] + sums[1]) / nk.clamp_min(1.0), torch.zeros((), device=dev))
        ctx.mark_non_differentiable(uv, mask, nk)
        return loss, uv, mask.bool(), nk

    @staticmethod
    def backward(ctx, g_loss, *_):
        g_x, g_P = ctx.saved_tensors
        sx, sP = ctx.shapes
        return (g_x * g_loss).view(sx), (g_P * g_loss).view(sP), None, None, None, None, None
