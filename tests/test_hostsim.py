"""Kernel ARITHMETIC checked on the CPU: the very same csrc/*.cu* sources compiled by g++ against the SIMT emulator
(tests/hostsim/simt.h) and driven through the same C ABI + python layer as on the GPU.  This is test infrastructure
for a container without a GPU; the product library is the nvcc build and the -m gpu tests are the parity tests proper."""
import pytest
import torch

from oracle import port

from . import common
from . import golden_checks as gc


@pytest.fixture(scope="module", autouse=True)
def hostsim_lib():
    from levels2fm_b200 import _C
    from .hostsim import harness
    old = _C._lib
    _C._lib = harness.get()
    yield _C._lib
    _C._lib = old


def test_hash_indices_bit_exact(hostsim_lib):
    from levels2fm_b200 import ops
    from oracle import hashgrid
    cfg = port.SceneCfg()
    meta = cfg.grid()
    grid = ops.GridSpec(16, 2, 19, 16, cfg.per_level_scale).resolve(hostsim_lib)
    g = torch.Generator().manual_seed(0)
    u = torch.rand(3000, 3, generator=g)
    u[:300] = u[:300] * 3 - 1
    u[300:600] = torch.round(u[300:600] * 64) / 64
    table = torch.randn(meta.n_params, generator=g)
    enc, idx = ops.grid_encode_raw(hostsim_lib, grid, table, u, want_idx=True)
    for l, lv in enumerate(meta.levels):
        ref_idx, _ = hashgrid.corner_indices(u, lv)
        assert torch.equal(idx[:, l, :].to(torch.int64) & 0xFFFFFFFF, ref_idx + lv.offset)
    ref = hashgrid.encode(u, table, meta)
    assert (enc - ref).abs().max() <= 1e-5 * ref.abs().max()


@pytest.mark.parametrize("dataset,n_levels,layers,n_samples,n_rays,dual", [
    ("DTU", 4, (None, 64, 16), 16, 9, False),
    ("ETH3D", 16, (None, 64, 64, 64, 16), 24, 5, False),
    ("bmvs", 16, (None, 64, 16), 12, 4, True),
    ("DTU", 16, (None, 64, 64, 16), 33, 3, False),
])
@pytest.mark.parametrize("mode", ["auto", "tc"])     # auto: these sizes run the exact SIMT backward; tc: force the tensor-core kernel
def test_render_forward_backward_matches_oracle(dataset, n_levels, layers, n_samples, n_rays, dual, mode):
    from levels2fm_b200 import ops
    opt = common.make_opt(dataset, "cpu", n_levels, layers, n_samples, dual)
    ops.BACKWARD_MODE = mode
    try:
        outs, grads = common.render_parity_case(opt, n_levels, 2, n_rays)
    finally:
        ops.BACKWARD_MODE = "auto"
    for k, (a, b) in outs.items():
        assert common.rel_err(a, b) < 1e-4, k
    for k, (a, b) in grads.items():
        assert common.cosine(a, b) > 1 - 1e-6, k
        assert common.rel_err(a, b) < 2e-3, (k, common.rel_err(a, b))


def test_tensor_core_backward_agrees_with_simt_backward():
    """same sources under the emulator: the tcgen05 backward kernel (emulated TMEM / MMA / bulk copies) vs the SIMT one."""
    from levels2fm_b200 import ops
    res = {}
    for simt in (True, False):
        ops.BACKWARD_MODE = "simt" if simt else "tc"
        try:
            opt = common.make_opt("DTU", "cpu", 16, (None, 64, 64, 16), 21, False)
            _, res[simt] = common.render_parity_case(opt, 16, 2, 5)
        finally:
            ops.BACKWARD_MODE = "auto"
    for k in res[True]:
        a, b = res[False][k][0], res[True][k][0]
        assert common.rel_err(a, b) < 5e-5, (k, common.rel_err(a, b))


def test_forced_tensor_core_backward_refuses_what_it_cannot_do(hostsim_lib):
    """ls2fm_field_backward_tc is strict: without the operand image (its weights stream from it) it reports an error instead of
    silently running something else; with it, first-order launches (no gradient on the normals) run on its single-channel variant."""
    from levels2fm_b200 import ops
    opt = common.make_opt("DTU", "cpu", 4, (None, 64, 16), 16)
    sdf, _, _ = common.build_models(opt)
    spec, table = sdf.field_spec(), sdf.table().detach()
    theta = sdf.SDF_MLP.theta().detach().contiguous()
    x = torch.rand(70, 3) * 1.6 - 0.8
    pts = ops._points(hostsim_lib, x.contiguous(), None, None, None)
    g = torch.ones(70)
    d_table, d_theta = torch.zeros_like(table), torch.zeros_like(theta)
    gn0 = torch.ones(70, 3)

    def forced_tc(image, g_nrm):
        f = spec.c_field(hostsim_lib, table, theta, image)
        ops._call(hostsim_lib, "field_backward_tc", hostsim_lib.dll.ls2fm_field_backward_tc, f, pts, None, None, hostsim_lib.ptr(g),
                  hostsim_lib.ptr(g_nrm), None, None, None, hostsim_lib.ptr(d_table), hostsim_lib.ptr(d_theta), None, None, None,
                  None, hostsim_lib.stream())

    with pytest.raises(RuntimeError, match="field_backward_tc"):      # no operand image
        forced_tc(None, gn0)
    image = ops.field_prepare_raw(hostsim_lib, spec, table, theta, None)
    assert float(d_table.abs().max()) == 0.0 and float(d_theta.abs().max()) == 0.0      # the refused launch wrote nothing
    # first-order only (no gradient on the normals: RadF.Geo_enc under dual_field): the single-channel tensor-core kernel
    forced_tc(image, None)
    d_t1, d_th1 = torch.zeros_like(table), torch.zeros_like(theta)
    ops.field_backward_raw(hostsim_lib, spec, table, theta, pts, None, None, g, None, None, None, None, d_t1, d_th1, mode="simt")
    assert common.rel_err(d_table, d_t1) < 5e-5 and common.rel_err(d_theta, d_th1) < 5e-5
    d_table.zero_()
    d_theta.zero_()
    # and with both it agrees with the SIMT kernel
    gn = torch.randn(70, 3)
    ops.field_backward_raw(hostsim_lib, spec, table, theta, pts, None, None, g, gn, None, None, None, d_table, d_theta, image=image, mode="tc")
    d_table2, d_theta2 = torch.zeros_like(table), torch.zeros_like(theta)
    ops.field_backward_raw(hostsim_lib, spec, table, theta, pts, None, None, g, gn, None, None, None, d_table2, d_theta2, mode="simt")
    assert common.rel_err(d_table, d_table2) < 5e-5 and common.rel_err(d_theta, d_theta2) < 5e-5


@pytest.mark.parametrize("n", [1, 63, 64, 65, 200])
def test_tensor_core_backward_ragged_sizes(hostsim_lib, n):
    """Tile edges of the tcgen05 backward kernel (64-sample tiles, two per emulated CTA sweep): a single point, one short of a
    tile, exactly one, one over, several with a ragged tail -- forced through ls2fm_field_backward_tc, against the SIMT kernel."""
    from levels2fm_b200 import ops
    opt = common.make_opt("DTU", "cpu", 16, (None, 64, 64, 16), 16)
    cfg = common.cfg_of(opt, 16)
    sdf_sd, _ = port.random_state(cfg, seed=4, table_std=0.2)
    sdf, _, _ = common.build_models(opt)
    sdf.load_state_dict(sdf_sd)
    spec, table = sdf.field_spec(), sdf.table().detach()
    theta = sdf.SDF_MLP.theta().detach().contiguous()
    g = torch.Generator().manual_seed(n)
    x = (torch.rand(n, 3, generator=g) * 1.6 - 0.8).contiguous()
    g_sdf, g_nrm = torch.randn(n, generator=g), torch.randn(n, 3, generator=g)
    pts = ops._points(hostsim_lib, x, None, None, None)
    image = ops.field_prepare_raw(hostsim_lib, spec, table, theta, None)
    out = {}
    for mode in ("tc", "simt"):
        d_table, d_theta = torch.zeros_like(table), torch.zeros_like(theta)
        ops.field_backward_raw(hostsim_lib, spec, table, theta, pts, None, None, g_sdf, g_nrm, None, None, None, d_table, d_theta,
                               image=image, mode=mode)
        out[mode] = (d_table, d_theta)
    for a, b, name in zip(out["tc"], out["simt"], ("d_table", "d_theta")):
        assert float(b.abs().max()) > 0
        assert common.rel_err(a, b) < 5e-5, (name, common.rel_err(a, b))


@pytest.mark.parametrize("gather_warps,depth", [(4, 2), (8, 2), (4, 4), (8, 4)])
@pytest.mark.parametrize("n,n_levels,layers", [(1, 16, (None, 64, 64, 64, 16)), (127, 16, (None, 64, 16)), (128, 4, (None, 64, 16)),
                                               (129, 16, (None, 64, 64, 16)), (700, 16, (None, 64, 64, 64, 16))])
def test_experimental_warp_specialised_forward(hostsim_lib, monkeypatch, n, n_levels, layers, gather_warps, depth):
    """ls2fm_field_forward_ws (gather warps feeding MLP warps through a shared-memory MMA operand, counted mbarriers, named
    barrier) against the default values-only kernel and the oracle, on tile-edge sizes.  Emulator only: the kernel is round-2
    groundwork (one hardware run so far: bit-identical, not yet faster).  gather_warps: one or two threads per sample."""
    from levels2fm_b200 import ops
    monkeypatch.setenv("LS2FM_WS_GATHER_WARPS", str(gather_warps))
    monkeypatch.setenv("LS2FM_WS_DEPTH", str(depth))
    opt = common.make_opt("DTU", "cpu", n_levels, layers, 16)
    cfg = common.cfg_of(opt, n_levels)
    sdf_sd, _ = port.random_state(cfg, seed=6, table_std=0.2)
    sdf, _, _ = common.build_models(opt)
    sdf.load_state_dict(sdf_sd)
    spec, table = sdf.field_spec(), sdf.table().detach()
    theta = sdf.SDF_MLP.theta().detach().contiguous()
    x = (torch.rand(n, 3, generator=torch.Generator().manual_seed(n)) * 1.6 - 0.8).contiguous()
    pts = ops._points(hostsim_lib, x, None, None, None)
    image = ops.field_prepare_raw(hostsim_lib, spec, table, theta, None)
    ref_y, ref_sdf, _, _ = ops.field_forward_raw(hostsim_lib, spec, table, theta, pts, None, want_y=True, image=image)
    ops.FORWARD_WS = True
    try:
        y, s, _, _ = ops.field_forward_raw(hostsim_lib, spec, table, theta, pts, None, want_y=True, image=image)
        y2, s2, _, _ = ops.field_forward_raw(hostsim_lib, spec, table, theta, pts, None, want_y=True)       # no operand image
    finally:
        ops.FORWARD_WS = False
    assert common.rel_err(y, ref_y) < 2e-6 and common.rel_err(s, ref_sdf) < 2e-6
    assert common.rel_err(y2, ref_y) < 2e-6 and common.rel_err(s2, ref_sdf) < 2e-6
    o_sdf, o_feat = port.infer_sdf(x, sdf_sd, cfg, "ret_all")
    assert common.rel_err(s, o_sdf.reshape(-1)) < 1e-4


def test_golden_c1_through_kernels():
    gold = gc.load("c1_render.npz")
    out, grads, loss = gc.run_c1_product(gold, "cpu")
    gc.check_c1(out, grads, loss, gold)


def test_golden_sphere_tracing_through_kernels():
    gold = gc.load("st_dtu.npz")
    gc.check_st(*gc.run_st_product(gold, "cpu"), gold)


def test_ragged_and_tiny_inputs():
    """n not a multiple of the 8-sample warp tile / 64-sample CTA step; a single point; zero points."""
    opt = common.make_opt("DTU", "cpu", 4, (None, 64, 16), 16)
    cfg = common.cfg_of(opt, 4)
    sdf_sd, _ = port.random_state(cfg, seed=2, table_std=0.2)
    sdf, _, _ = common.build_models(opt)
    sdf.load_state_dict(sdf_sd)
    g = torch.Generator().manual_seed(0)
    for n in (1, 7, 65, 130):
        x = torch.rand(n, 3, generator=g) * 1.6 - 0.8
        s, f = sdf.infer_sdf(x, mode="ret_all")
        rs, rf = port.infer_sdf(x, sdf_sd, cfg, "ret_all")
        assert common.rel_err(s, rs) < 1e-5 and common.rel_err(f, rf) < 1e-5
        n_ours = sdf.gradient(x.clone())
        n_ref = port.sdf_gradient(x.clone(), sdf_sd, cfg)
        assert common.rel_err(n_ours, n_ref.detach()) < 1e-5
    assert sdf.infer_sdf(torch.zeros(0, 3)).shape == (0, 1)
    assert sdf.infer_sdf(torch.zeros(2, 5, 3)).shape == (2, 5, 1)


def test_golden_error_bounded_sampler_through_kernels():
    gold = gc.load("c2_sampler.npz")
    gc.check_c2(*gc.run_c2_product(gold, "cpu"), gold)


@pytest.mark.parametrize("eps,N,std", [(0.002, 16, 0.05), (0.02, 32, 0.1)])
def test_error_bounded_sampler_hard_cases(eps, N, std):
    gc.sampler_hard_case("cpu", eps, N, std)


def test_fused_sphere_trace_kernel_internals():
    gc.sphere_trace_internals("cpu")


def test_fused_render_loss_matches_torch():
    from levels2fm_b200 import ops, synthetic
    g = torch.Generator().manual_seed(0)
    rgb = torch.rand(2, 7, 3, generator=g, requires_grad=True)
    gt = torch.rand(2, 7, 3, generator=g)
    nrm = (torch.randn(2, 7, 5, 3, generator=g) * 1.3).requires_grad_(True)
    ref = synthetic.render_loss({"rgb": rgb, "normals": nrm}, gt)
    gr_ref = torch.autograd.grad(ref, [rgb, nrm])
    out = synthetic.render_loss_fused({"rgb": rgb, "normals": nrm}, gt)
    gr = torch.autograd.grad(out * 1.0, [rgb, nrm])
    assert abs(out.item() - ref.item()) < 1e-4 * abs(ref.item())
    for a, b in zip(gr, gr_ref):
        assert common.rel_err(a, b) < 1e-5


@pytest.mark.parametrize("layers,dual", [((None, 64, 64, 64, 16), False), ((None, 64, 16), True)])
def test_fused_param_prep_matches_torch_weight_norm_and_composition(layers, dual):
    from levels2fm_b200 import ops
    from levels2fm_b200.models import base
    opt = common.make_opt("DTU", "cpu", 16, layers, 16, dual)
    sdf, rad, _ = common.build_models(opt)
    g = torch.Generator().manual_seed(0)
    with torch.no_grad():
        for p in list(sdf.SDF_MLP.parameters()) + list(rad.Rad_dec.parameters()):
            p.copy_(torch.randn(p.shape, generator=g) * 0.3 + 0.1)
    theta, w_eff, b_eff = base.prepare_params(sdf.SDF_MLP, rad.Rad_dec)
    theta_ref = ops.pack_theta(base.effective_layers(sdf.SDF_MLP.mlp))
    w_ref, b_ref = ops.compose_affine(base.effective_layers(rad.Rad_dec.mlp_radiance))
    assert common.rel_err(theta, theta_ref) < 1e-6 and common.rel_err(w_eff, w_ref) < 1e-5 and common.rel_err(b_eff, b_ref) < 1e-5
    ct, cw, cb = torch.randn(theta.shape, generator=g), torch.randn(w_eff.shape, generator=g), torch.randn(3, generator=g)
    params = list(sdf.SDF_MLP.parameters()) + list(rad.Rad_dec.parameters())
    ours = torch.autograd.grad((theta * ct).sum() + (w_eff * cw).sum() + (b_eff * cb).sum(), params)
    ref = torch.autograd.grad((theta_ref * ct).sum() + (w_ref * cw).sum() + (b_ref * cb).sum(), params)
    for a, b in zip(ours, ref):
        assert common.rel_err(a, b) < 2e-5, (a.shape, common.rel_err(a, b))
    # geometry only / radiance only
    assert common.rel_err(sdf.SDF_MLP.theta(), theta_ref) < 1e-6
    w2, b2 = rad.Rad_dec.effective_affine()
    (w2.sum() + b2.sum()).backward()
    assert all(p.grad is not None for p in rad.Rad_dec.parameters())


def test_ray_generation_kernel_matches_oracle():
    gc.rays_case("cpu")


# ----------------------------------------------------------------------------- gradients w.r.t. the sample positions
def test_position_gradient_first_order():
    from . import input_grad_checks as ig
    ig.first_order("cpu", n=120)


def test_position_gradient_of_the_normals():
    from . import input_grad_checks as ig
    ig.hessian_vector("cpu", n=90)


@pytest.mark.parametrize("dataset", ["DTU", "ETH3D"])
def test_ba_surface_point_pattern_gradients(dataset):
    from . import input_grad_checks as ig
    ig.ba_surface_pattern("cpu", n=150, dataset=dataset)


@pytest.mark.parametrize("dataset,dual", [("DTU", False), ("bmvs", True)])
def test_pose_gradient_through_renderer(dataset, dual):
    from . import input_grad_checks as ig
    ig.pose_gradient_through_renderer("cpu", dataset, dual, n_pix=6, n_samples=10)


def test_aabb_grad_flag_reproduces_reference_error():
    from . import input_grad_checks as ig
    ig.aabb_grad_flag("cpu")


def test_radf_geometry_feat_position_gradient():
    from . import input_grad_checks as ig
    ig.radf_geometry_feat_input_grad("cpu", n=60)


# ----------------------------------------------------------------------------- forward-only entry points
@pytest.mark.parametrize("use_bounds,dataset", [(False, "DTU"), (True, "ETH3D")])
def test_sdf_grid_volume_matches_reference_point_arithmetic(use_bounds, dataset):
    from . import inference_checks as ic
    ic.grid_case("cpu", N=9, dataset=dataset, volume_size=2.0 if not use_bounds else 10.0, use_bounds=use_bounds, chunk=200)


@pytest.mark.parametrize("dual", [False, True])
def test_render_image_in_slices(dual):
    from . import inference_checks as ic
    ic.image_case("cpu", H=6, W=8, dual=dual, slice_rays=20)


# ----------------------------------------------------------------------------- fused loss tail
@pytest.mark.parametrize("eik_masked,none_finished", [(True, False), (False, False), (True, True)])
def test_fused_loss_tail_matches_reference_tail(eik_masked, none_finished):
    from . import loss_checks as lc
    lc.tail_case("cpu", eik_masked=eik_masked, none_finished=none_finished)


def test_gradient_bucket_direct_scatter_equals_autograd_accumulation():
    """parallel.GradBucket(direct=True): the backward kernels scatter the table gradient straight into the bucket (no zero-filled
    temporary + autograd add); same numbers as the plain autograd route, and a second backward accumulates."""
    from levels2fm_b200 import parallel, synthetic
    opt = common.make_opt("DTU", "cpu", 16, (None, 64, 16), 8, True)
    res = {}
    for direct in (True, False):
        torch.manual_seed(0)
        sdf, rad, ren = common.build_models(opt)
        cfg = common.cfg_of(opt, 16)
        sdf_sd, rad_sd = port.random_state(cfg, seed=3, table_std=0.2)
        sdf.load_state_dict(sdf_sd)
        rad.load_state_dict(rad_sd)
        bucket = parallel.GradBucket(list(sdf.parameters()) + list(rad.parameters()), direct=direct)
        center, ray = common.make_rays(1, 6, 1.0)
        gt = torch.rand(1, 6, 3, generator=torch.Generator().manual_seed(1))
        for _ in range(2):
            synthetic.render_loss_fused(ren.forward(opt, center, ray, sdf, rad), gt).backward()
        assert all(p.grad.data_ptr() == bucket.flat[o:o + 1].data_ptr() for p, o in zip(bucket.params, bucket.offsets))
        res[direct] = bucket.flat.clone()
    assert float(res[False].abs().max()) > 0
    assert common.rel_err(res[True], res[False]) < 1e-6


def test_se3_to_SE3_kernel_matches_oracle():
    gc.se3_case("cpu")


def test_ba_surface_terms_match_reference_block():
    from . import ba_checks
    ba_checks.ba_terms_case("cpu", n=40)


def test_error_bounded_sampler_against_reference_hard_case_golden():
    gc.sampler_golden_hard_case("cpu")


def test_standalone_grid_encoding_is_double_differentiable():
    gc.grid_encode_double_backward_case("cpu", m=120)


@pytest.mark.parametrize("dataset,dual", [("DTU", False), ("bmvs", True)])
def test_position_gradients_on_the_tensor_core_route(dataset, dual):
    from . import input_grad_checks as ig
    ig.tensor_core_route_matches_simt_route("cpu", n_rays=9, n_samples=15, dataset=dataset, dual=dual)


def test_sphere_tracing_sync_free_form_equals_default():
    gc.sphere_trace_sync_free_case("cpu")


@pytest.mark.parametrize("n,layers,ray_mode", [(1500, (None, 64, 64, 64, 16), False), (1100, (None, 64, 64, 16), True), (900, (None, 64, 16), False),
                                               (257, (None, 64, 64, 64, 16), True)])
def test_values_only_kernel_many_pairs_all_depths(n, layers, ray_mode):
    from . import inference_checks as ic
    ic.values_only_kernel_case("cpu", n, layers, ray_mode=ray_mode)


@pytest.mark.parametrize("n,layers,ray_mode", [(1, (None, 64, 16), False), (127, (None, 64, 16), True), (128, (None, 64, 64, 16), False), (129, (None, 64, 64, 64, 16), False),
                                               (700, (None, 64, 16), False), (520, (None, 64, 64, 64, 16), True)])
def test_feature_only_tensor_core_backward(n, layers, ray_mode):
    from . import input_grad_checks as ig
    ig.feature_only_tensor_core_backward("cpu", n, layers, ray_mode=ray_mode)


def test_empty_ray_and_point_batches():
    from . import inference_checks as ic
    ic.empty_batch_case("cpu")
