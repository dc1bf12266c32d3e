"""CPU tests of the oracle itself: pinned against the fixtures the reference's own python produced, against the
scalar C restatement of the hash-grid arithmetic, and against fp64 finite differences."""
import ctypes
import os
import subprocess

import numpy as np
import pytest
import torch

from oracle import aabb, hashgrid, port

from . import golden_checks as gc

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def cref():
    subprocess.run(["make", "-s", "-C", os.path.join(ROOT, "oracle")], check=True)
    lib = ctypes.CDLL(os.path.join(ROOT, "oracle", "_ref", "libhashgrid_ref.so"))

    class L(ctypes.Structure):
        _fields_ = [("scale", ctypes.c_float), ("resolution", ctypes.c_uint32), ("offset", ctypes.c_uint32),
                    ("size", ctypes.c_uint32), ("hashed", ctypes.c_uint32)]
    lib.ref_grid_meta.restype = ctypes.c_uint32
    return lib, L


@pytest.mark.parametrize("half,n_levels", [(1.0, 16), (2.0, 16), (5.0, 16), (1.0, 4)])
def test_grid_meta_matches_c(cref, half, n_levels):
    lib, L = cref
    b = port.SceneCfg(bound_min=(-half,) * 3, bound_max=(half,) * 3, n_levels=n_levels).per_level_scale
    meta = hashgrid.grid_meta(n_levels, 2, 19, 16, b)
    out = (L * n_levels)()
    total = lib.ref_grid_meta(n_levels, 19, 16, ctypes.c_float(b), out)
    assert total == meta.n_entries
    for lv, c in zip(meta.levels, out):
        assert (np.float32(lv.scale), lv.resolution, lv.offset, lv.size, int(lv.hashed)) == \
               (np.float32(c.scale), c.resolution, c.offset, c.size, c.hashed)
    if half == 1.0 and n_levels == 16:
        assert meta.n_entries == 6098120 and meta.levels[-1].resolution == 2048


def test_corner_indices_bit_exact_vs_c(cref):
    lib, L = cref
    meta = hashgrid.grid_meta(16, 2, 19, 16, port.SceneCfg().per_level_scale)
    g = torch.Generator().manual_seed(0)
    u = torch.rand(20000, 3, generator=g)
    u[:2000] = u[:2000] * 3 - 1                                    # out-of-range coordinates wrap through the casts
    u[2000:3000] = torch.round(u[2000:3000] * 64) / 64             # exactly on cell boundaries of coarse levels
    un = np.ascontiguousarray(u.numpy())
    for lv in meta.levels:
        idx, w = hashgrid.corner_indices(u, lv)
        c = L(lv.scale, lv.resolution, lv.offset, lv.size, int(lv.hashed))
        ci = np.zeros((u.shape[0], 8), dtype=np.uint32)
        cw = np.zeros((u.shape[0], 3), dtype=np.float32)
        lib.ref_grid_corners(ctypes.byref(c), un.ctypes.data_as(ctypes.c_void_p), ctypes.c_int64(u.shape[0]),
                             ci.ctypes.data_as(ctypes.c_void_p), cw.ctypes.data_as(ctypes.c_void_p))
        assert np.array_equal(idx.numpy().astype(np.uint32), ci), f"level res {lv.resolution}"
        assert np.array_equal(w.numpy(), cw)


def test_grid_gradcheck_fp64():
    meta = hashgrid.grid_meta(2, 2, 5, 3, 1.7)
    g = torch.Generator().manual_seed(0)
    table = torch.randn(meta.n_params, dtype=torch.float64, generator=g, requires_grad=True)
    u = (torch.rand(3, 3, dtype=torch.float64, generator=g) * 0.9 + 0.05).requires_grad_(True)
    assert torch.autograd.gradcheck(lambda uu, tt: hashgrid.encode(uu, tt, meta), (u, table), eps=1e-7, atol=1e-6)
    assert torch.autograd.gradgradcheck(lambda uu, tt: hashgrid.encode(uu, tt, meta), (u, table), eps=1e-7, atol=1e-6)


def test_aabb_cases():
    o = torch.tensor([[0.0, 0.0, -3.0], [0.0, 0.0, 0.0], [0.0, 5.0, -3.0], [0.0, 0.0, 3.0]])
    d = torch.tensor([[0.0, 0.0, 1.0], [0.0, 0.0, 2.0], [0.0, 0.0, 1.0], [0.0, 0.0, 1.0]])
    cnt, hits, idx = aabb.ray_aabb_intersect(o, d, torch.zeros(1, 3), torch.ones(1, 3), 1)
    assert hits[:, 0].tolist() == [[2.0, 4.0], [0.0, 0.5], [-1.0, -1.0], [-1.0, -1.0]]
    assert cnt.tolist() == [1, 1, 0, 0]


def test_sphere_init_is_a_sphere():
    cfg = port.SceneCfg()
    sd, _ = port.random_state(cfg, seed=0, table_std=1e-4, generic_weights=False, sphere_bias=0.5)
    d = torch.nn.functional.normalize(torch.randn(512, 3, generator=torch.Generator().manual_seed(0)), dim=-1)
    s_out = port.infer_sdf(d * 0.9, sd, cfg)[:, 0]
    s_in = port.infer_sdf(d * 0.1, sd, cfg)[:, 0]
    # geometric init (models/base.py:184-199): sdf ~ |x| - bias in expectation over the random first layer
    assert abs(s_out.mean().item() - 0.4) < 0.1 and abs(s_in.mean().item() + 0.4) < 0.1
    assert (s_out > 0).all() and (s_in < 0).all()


def test_constant_density_compositing_closed_form():
    R, N = 3, 50
    t = torch.linspace(1.0, 2.0, N).expand(1, R, N).contiguous()
    ray = torch.tensor([[[0.0, 0.0, 2.0]]]).expand(1, R, 3)
    sigma = torch.full((1, R, N), 0.7)
    rgb, prob = port.composite(ray, torch.ones(1, R, N, 3), sigma, t)
    opacity = prob.sum(dim=2)[..., 0]
    expect = 1 - np.exp(-0.7 * 2.0 * (2.0 - 1.0))      # the last sample carries no weight -> integrates to t_{N-1}
    assert torch.allclose(opacity, torch.full_like(opacity, expect), atol=1e-5)
    assert torch.allclose(rgb[..., 0], opacity, atol=1e-6)


def test_port_matches_reference_golden_c1():
    gold = gc.load("c1_render.npz")
    out, grads, loss = gc.run_c1_oracle(gold)
    gc.check_c1(out, grads, loss, gold)


def test_port_matches_reference_golden_sphere_tracing():
    gold = gc.load("st_dtu.npz")
    gc.check_st(*gc.run_st_oracle(gold), gold, exact=True)


def test_port_matches_reference_golden_sampler():
    gold = gc.load("c2_sampler.npz")
    cfg = port.SceneCfg(n_levels=16, sample_intvs=64, final_sample_intvs=64, volsdf_sampling=True,
                        sdf_layers=(None, 64, 64, 64, 16))
    sdf_sd, rad_sd = port.random_state(cfg, seed=6, table_std=0.02, generic_weights=False, hash_weight_std=0.05)
    t, beta_plus, iters = port.volsdf_sampling(gold["center"], gold["ray"], sdf_sd, cfg)
    assert torch.equal(iters, gold["iters"])
    gc.assert_close(t, gold["t"], tol=1e-5, what="sampler t")
    gc.assert_close(beta_plus, gold["beta_plus"], tol=1e-5, what="beta plus")
    out = port.render_forward(gold["center"], gold["ray"], sdf_sd, rad_sd, cfg)
    gc.assert_close(out["rgb"], gold["out.rgb"], what="c2 rgb")
    gc.assert_close(out["depth_mlp"], gold["out.depth_mlp"], what="c2 depth")


def test_port_matches_reference_golden_sampler_hard_case():
    """The bisection on beta+ and the give-up path (iters = -1) of the port against the reference's own run of
    models/Renderer.py:281-321 (fixture: oracle/make_golden.c2_sampler_hard, 40 of 48 rays never converge)."""
    gold = gc.load("c2_sampler_hard.npz")
    cfg = port.SceneCfg(n_levels=16, sample_intvs=16, final_sample_intvs=24, volsdf_sampling=True, eps=0.002, max_upsample_iter=4)
    sdf_sd, _ = port.random_state(cfg, seed=8, table_std=0.05, generic_weights=False, hash_weight_std=0.1)
    t, beta_plus, iters = port.volsdf_sampling(gold["center"], gold["ray"], sdf_sd, cfg)
    assert (gold["iters"] == -1).sum() >= 30 and (gold["iters"] >= 0).sum() >= 5
    assert torch.equal(iters, gold["iters"])
    gc.assert_close(t, gold["t"], tol=1e-5, what="hard sampler t")
    gc.assert_close(beta_plus, gold["beta_plus"], tol=1e-5, what="hard sampler beta plus")


def test_renderer_api_compat_pieces_match_the_oracle():
    """Renderer.composite / error_bound / sample_pdf stay callable with the reference's semantics (pure tensor math)."""
    from levels2fm_b200.models.Renderer import Renderer
    from . import common
    ren = Renderer(common.make_opt("DTU", "cpu"))
    g = torch.Generator().manual_seed(0)
    d = torch.sort(torch.rand(2, 5, 20, generator=g) * 3, dim=-1).values
    sdf = torch.randn(2, 5, 20, generator=g) * 0.3
    a, b = torch.tensor(20.0), torch.tensor(0.05)
    assert torch.allclose(ren.error_bound(d, sdf, a, b), port.error_bound(d, sdf, a, b), equal_nan=True)
    w = torch.rand(2, 5, 19, generator=g)
    assert torch.allclose(ren.sample_pdf(d, w, 12, det=True), port.sample_pdf_det(d, w, 12))
    ray = torch.randn(2, 5, 3, generator=g)
    rgbs = torch.rand(2, 5, 20, 3, generator=g)
    sig = torch.rand(2, 5, 20, generator=g) * 5
    r1, p1 = ren.composite(ray, rgbs, sig, d[..., None])
    r2, p2 = port.composite(ray, rgbs, sig, d)
    assert torch.allclose(r1, r2) and torch.allclose(p1, p2)
    assert torch.allclose(ren.sdf_to_sigma(sdf, a, b), port.sdf_to_sigma(sdf, a, b))
