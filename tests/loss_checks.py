"""The fused loss tail (levels2fm_b200.losses / ls2fm_render_tail) against a line-by-line torch restatement of the reference's
CameraSet.render tail and the stages' compute_loss.  Shared by emulator (cpu) and GPU tests."""
import torch
import torch.nn.functional as torch_F

from . import common


def reference_tail(rgb, normals, depth_mlp, rgbs_gt, d_points, mask_finish, eik_masked, w):
    """/root/reference/pipelines/Camera.py:510-537 (dataset in the listed set: only the mask_finish smooth-L1 term),
    /root/reference/pipelines/BA.py:190-204 (eikonal masked by mask_bg) / rendering_refine.py:99-107 (all samples),
    summarize_loss: sum of 10**w * term."""
    mask_finish = mask_finish.view(*depth_mlp.shape)
    mask_bg = (rgbs_gt.mean(dim=-1) < 0.95) & (rgbs_gt.mean(dim=-1) > 0.05)
    mask_finish = mask_finish & mask_bg.view(*mask_finish.shape)
    if mask_finish.sum() > 0:
        d_consistent = torch_F.smooth_l1_loss(d_points.view(*depth_mlp.shape)[mask_finish], depth_mlp[mask_finish], reduction="mean")
    else:
        d_consistent = torch.zeros_like(d_points).mean()
    PSNR = -10 * torch_F.mse_loss(rgb[mask_bg], rgbs_gt[mask_bg]).log10()
    rgb_loss = torch_F.l1_loss(rgb, rgbs_gt)
    nn = torch.norm(normals[mask_bg], dim=-1) if eik_masked else torch.norm(normals, dim=-1)
    eik = torch_F.l1_loss(nn, torch.ones_like(nn))
    total = 10 ** w[0] * rgb_loss + 10 ** w[1] * eik + 10 ** w[2] * d_consistent
    return {"loss": total, "rgb_loss": rgb_loss, "eikonal_loss": eik, "DC_loss": d_consistent, "PSNR": PSNR, "mask_bg": mask_bg,
            "mask_finish": mask_finish[..., 0]}


def tail_case(device, B=2, R=37, N=9, eik_masked=True, none_finished=False, seed=0):
    from levels2fm_b200 import losses
    g = torch.Generator().manual_seed(seed)
    gt = torch.rand(B, R, 3, generator=g)
    gt[0, :5] = 0.99                  # background-white rays (mean > 0.95)
    gt[1, :4] = 0.01                  # background-black rays
    w = (3.0, 2.0, 1.0)
    leaves = {"rgb": torch.rand(B, R, 3, generator=g), "normals": torch.randn(B, R, N, 3, generator=g) * 1.3,
              "depth_mlp": torch.rand(B, R, 1, generator=g) * 3, "d_points": torch.rand(B, R, generator=g) * 5}
    leaves["d_points"][0, 7] = leaves["depth_mlp"][0, 7, 0] + 2.5          # |d| > 1: the linear branch of smooth-L1
    mask_finish = torch.rand(B * R, 1, generator=g) > (2.0 if none_finished else 0.4)
    res = {}
    for who in ("ours", "ref"):
        dev = device if who == "ours" else "cpu"
        t = {k: v.clone().to(dev).requires_grad_(True) for k, v in leaves.items()}
        if who == "ours":
            out = losses.camera_render_tail({"rgb": t["rgb"], "normals": t["normals"], "depth_mlp": t["depth_mlp"]}, gt.to(dev),
                                            t["d_points"], mask_finish.to(dev), w, eik_masked)
        else:
            out = reference_tail(t["rgb"], t["normals"], t["depth_mlp"], gt, t["d_points"], mask_finish, eik_masked, w)
        grads = torch.autograd.grad(out["loss"] * 0.7, list(t.values()), allow_unused=True)
        res[who] = (out, grads)
    o, r = res["ours"][0], res["ref"][0]
    assert torch.equal(o["mask_bg"].cpu(), r["mask_bg"]) and torch.equal(o["mask_finish"].cpu(), r["mask_finish"])
    for k in ("loss", "rgb_loss", "eikonal_loss", "DC_loss", "PSNR"):
        a_, b_ = float(o[k].detach()), float(r[k].detach())
        assert abs(a_ - b_) <= 1e-5 * max(abs(b_), 1e-6), (k, a_, b_)
    for name, a, b in zip(leaves, res["ours"][1], res["ref"][1]):
        if b is None:
            assert a is None or float(a.abs().max()) == 0.0, name
            continue
        assert common.rel_err(a.cpu(), b) < 1e-5, (name, common.rel_err(a.cpu(), b))
