"""The slice of the reference's option tree that the render hot path reads.

The reference parses ``options/*.yaml`` into an EasyDict (utils/options.py) and hands it to every hot-path
constructor.  That parser is outside the hot path; this module only provides an attribute-dict with the same
keys and the shipped defaults (options/LevelS2fM.yaml:3-40, options/{DTU,ETH3D,bmvs}.yaml) so that the drop-in
``models`` package can be constructed without the reference tree.  A real reference ``opt`` works unchanged.
"""
from __future__ import annotations

import copy
import os

HERE = os.path.dirname(os.path.abspath(__file__))
DEFAULT_HASH_CONFIG = os.path.join(HERE, "options", "config_hash_sdf.json")


class AttrDict(dict):
    """dict with attribute access, recursive (the subset of easydict.EasyDict the hot path relies on)."""

    def __init__(self, d=None, **kw):
        super().__init__()
        d = dict(d or {})
        d.update(kw)
        for k, v in d.items():
            self[k] = v

    def __setitem__(self, k, v):
        if isinstance(v, dict) and not isinstance(v, AttrDict):
            v = AttrDict(v)
        super().__setitem__(k, v)

    def __setattr__(self, k, v):
        self[k] = v

    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError:
            raise AttributeError(k)

    def __deepcopy__(self, memo):
        return AttrDict({k: copy.deepcopy(v, memo) for k, v in self.items()})


_DATASETS = {
    #            bound  inside  bias scale_mlp bgcolor       iters_max_st image_size
    "DTU":   dict(bound=1.0, inside=True, bias=0.5, scale_mlp=1, bgcolor=[0, 0, 0], iters_max_st=10, image_size=[1200, 1600]),
    "ETH3D": dict(bound=5.0, inside=False, bias=2.5, scale_mlp=5, bgcolor=[0, 0, 0], iters_max_st=20, image_size=[1033, 1551]),
    "bmvs":  dict(bound=2.0, inside=True, bias=1.0, scale_mlp=3, bgcolor=[1, 1, 1], iters_max_st=20, image_size=[576, 768]),
}


def default_opt(dataset: str = "DTU", device: str = "cuda", **overrides) -> AttrDict:
    """Hot-path options with the reference's shipped defaults.  ``overrides`` use dotted keys,
    e.g. ``default_opt("DTU", **{"SDF.VolSDF.sample_intvs": 64})``."""
    ds = _DATASETS[dataset]
    b = ds["bound"]
    opt = AttrDict(
        device=device, Res=100,
        Ablate_config=dict(dual_field=False),
        SDF=dict(
            arch=dict(layers=[None, 64, 16], skip=[]),
            NN_Init=dict(scale_mlp=ds["scale_mlp"], bias=ds["bias"], tf_init=True),
            VolSDF=dict(max_upsample_iter=6, sample_intvs=128, final_sample_intvs=64, volsdf_sampling=False,
                        iters_max_st=ds["iters_max_st"], eps=0.1, beta_init=0.05, rescale=1.0, beta_speed=1.0,
                        sdf_threshold=1e-3, max_bisection_itr=10),
            Hash_config=dict(config_file=DEFAULT_HASH_CONFIG)),
        RadF=dict(arch=dict(layers=[None, 64, 64, 3], skip=[])),
        Renderer=dict(rand_rays=8192),      # options/LevelS2fM.yaml:133-134; aabb_grad (ours, default on): see Renderer.volsdf_sampling
        data=dict(dataset=dataset, inside=ds["inside"], bg_sdf=False, bg_rad=2,
                  image_size=ds["image_size"], bound_max=[b, b, b], bound_min=[-b, -b, -b], bgcolor=ds["bgcolor"]),
    )
    opt.data.scene = "synthetic"
    opt.data["synthetic"] = AttrDict()    # per-scene overrides live under opt.data[<scene name>] (models/Renderer.py:25-29)
    opt.H, opt.W = ds["image_size"]
    for dotted, v in overrides.items():
        node = opt
        keys = dotted.split(".")
        for k in keys[:-1]:
            node = node[k]
        node[keys[-1]] = v
    return opt
