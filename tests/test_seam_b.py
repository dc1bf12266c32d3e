"""Seam B of INTEGRATION.md as a fact: the reference's UNMODIFIED models/{SDF,RadF,Renderer,base}.py running on
levels2fm_b200.compat.{tinycudann,vren} (our hash-grid and ray/AABB kernels, SIMT-emulator build here) instead of the oracle's
restatements of the two third-party packages: same outputs and gradients.  CPU only (/root/reference is not on the GPU box)."""
import sys

import pytest
import torch

from oracle import ref_shim

from . import common

pytestmark = pytest.mark.skipif(not ref_shim.available(), reason="reference tree not present")


@pytest.fixture(scope="module", autouse=True)
def hostsim_lib():
    from levels2fm_b200 import _C
    from .hostsim import harness
    old = _C._lib
    _C._lib = harness.get()
    yield _C._lib
    _C._lib = old


def _run(ref, opt, state, center, ray, gw):
    torch.manual_seed(0)
    sdf, rad, ren = ref_shim.build_models(opt)
    if state is not None:
        sdf.load_state_dict(state[0])
        rad.load_state_dict(state[1])
    else:
        g = torch.Generator().manual_seed(7)
        with torch.no_grad():
            for p in list(sdf.parameters()) + list(rad.parameters()):
                if p.numel() > 4096:
                    p.copy_(torch.randn(p.shape, generator=g) * 0.05)
            w0 = sdf.SDF_MLP.mlp[0]
            w0.weight_v[:, 3:] = torch.randn(w0.weight_v[:, 3:].shape, generator=g) * 0.05
    out = ren.forward(opt, center, ray, sdf, rad)
    loss = common.loss_fn(out, gw) if gw is not None else None
    if loss is not None:
        loss.backward()
    grads = {k: p.grad for k, p in list(sdf.named_parameters()) + list(rad.named_parameters())}
    return out, grads, (sdf.state_dict(), rad.state_dict())


def test_reference_models_on_our_native_packages():
    ref = ref_shim.load()
    from levels2fm_b200 import compat
    opt = ref_shim.make_opt("DTU", device="cpu", **{"SDF.VolSDF.sample_intvs": 12,
                                                      "SDF.Hash_config.config_file": ref_shim.REFERENCE_ROOT + "/options/config_hash_sdf.json"})
    center, ray = common.make_rays(1, 10, 1.0)
    center[0, 0] += 10.0                              # one ray that misses the box
    g = torch.Generator().manual_seed(3)
    # (a) the oracle's restatements of tcnn / vren underneath (what every other reference-side test uses)
    out_a, grads_a, state = _run(ref, opt, None, center, ray, None)
    gw = {k: torch.randn(out_a[k].shape, generator=g) for k in ["rgb", "depth_mlp", "normal_mlp", "sdfs_volume"]}
    out_a, grads_a, _ = _run(ref, opt, state, center, ray, gw)
    # (b) our kernels underneath: swap the two packages the reference imported
    tcnn_mod, vren_mod = sys.modules["tinycudann"], sys.modules["vren"]
    old = (tcnn_mod.Encoding, vren_mod.ray_aabb_intersect)
    tcnn_mod.Encoding, vren_mod.ray_aabb_intersect = compat.tinycudann.Encoding, compat.vren.ray_aabb_intersect
    try:
        out_b, grads_b, _ = _run(ref, opt, state, center, ray, gw)
    finally:
        tcnn_mod.Encoding, vren_mod.ray_aabb_intersect = old
    for k in ("rgb", "sdfs_volume", "normals", "depth_mlp", "normal_mlp"):
        assert common.rel_err(out_b[k].detach(), out_a[k].detach()) < 1e-4, k
    n = 0
    for k in grads_a:
        if grads_a[k] is None:
            continue
        n += 1
        assert common.cosine(grads_b[k], grads_a[k]) > 1 - 1e-6, (k, common.cosine(grads_b[k], grads_a[k]))
        assert common.rel_err(grads_b[k], grads_a[k]) < 2e-3, (k, common.rel_err(grads_b[k], grads_a[k]))
    assert n >= 10
