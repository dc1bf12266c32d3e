"""GPU diagnostic: error-bounded sampler determinism and render parity on the c2 golden fixture."""
import sys, os, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import port
from tests import golden_checks as gc, common
dev = "cuda"
gold = gc.load("c2_sampler.npz")
opt = gc.c2_opt(dev)
cfg = common.cfg_of(opt, 16)
sdf_sd, rad_sd = port.random_state(cfg, seed=6, table_std=0.02, generic_weights=False, hash_weight_std=0.05)
sdf, rad, ren = common.build_models(opt)
sdf.load_state_dict(sdf_sd); rad.load_state_dict(rad_sd)
c, r = gold["center"].to(dev), gold["ray"].to(dev)
ts = []
for k in range(3):
    t, bp, it = ren.volsdf_sampling(opt, c, r, sdf)
    torch.cuda.synchronize()
    ts.append(t.cpu())
    print(k, "t err vs gold", (ts[-1] - gold["t"]).abs().max().item(), "iters eq", torch.equal(it.cpu(), gold["iters"]))
print("call0 == call1", torch.equal(ts[0], ts[1]), "call1 == call2", torch.equal(ts[1], ts[2]))
out = ren.forward(opt, c, r, sdf, rad)
e = (out["rgb"].cpu() - gold["out.rgb"]).abs().amax(dim=-1)[0]
print("rgb err per ray top", e.sort(descending=True).values[:8].tolist(), e.argsort(descending=True)[:8].tolist())
# render with the golden t through our kernels
from levels2fm_b200 import ops
t2 = gold["t"].reshape(-1, 128).contiguous().to(dev)
w_eff, b_eff = rad.Rad_dec.effective_affine()
c2, r2 = c.reshape(-1, 3).contiguous(), r.reshape(-1, 3).contiguous()
s, _, n, rgbs = ops.FieldEval.apply(sdf.field_spec(), rad.rad_spec(), sdf.table(), sdf.SDF_MLP.theta(), w_eff, b_eff, None, None, c2, r2, t2, 0, None, False, True)
rgb, depth, normal, op = ops.Composite.apply(r2, t2, s.view(-1, 128), rgbs.view(-1, 128, 3), n.view(-1, 128, 3), sdf.beta, 1.0, (0, 0, 0))
e2 = (rgb.cpu() - gold["out.rgb"][0]).abs().amax(dim=-1)
print("rgb err with golden t", e2.max().item())
ref = port.render_forward(gold["center"], gold["ray"], sdf_sd, rad_sd, cfg, t=gold["t"])
print("sdf err", (s.cpu() - ref["sdfs_volume"].reshape(-1)).abs().max().item(), "rgbs err", (rgbs.cpu() - ref["rgbs"].reshape(-1, 3)).abs().max().item())
print("oracle rgb with golden t vs gold", (ref["rgb"] - gold["out.rgb"]).abs().max().item())
i = int(e.argmax())
print("worst ray", i, "t ours", ts[0][0, i, ::16].tolist(), "t gold", gold["t"][0, i, ::16].tolist())
