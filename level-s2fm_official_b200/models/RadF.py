"""Radiance field of Level-S2fM on the fused sm_100a kernels (reference: models/RadF.py).

State-dict keys as in the reference: ``Rad_dec.mlp_radiance.{i}.{bias,weight_g,weight_v}`` (+
``embed_fn.embedder_obj.params`` and ``Geo_enc.mlp.{i}.*`` with ``dual_field``).
"""
from __future__ import annotations

import numpy as np
import torch
import torch.nn as nn

from .. import ops
from .base import Geometry, Radiance, get_Embedder, get_layer_dims


class RadF(nn.Module):
    def __init__(self, opt):
        super().__init__()
        self.opt = opt
        dev = opt.device
        self.bound_max = torch.tensor(np.array(opt.data.bound_max), dtype=torch.float32, device=dev)[None, None, :]
        self.bound_min = torch.tensor(np.array(opt.data.bound_min), dtype=torch.float32, device=dev)[None, None, :]
        self.center = (self.bound_max + self.bound_min) / 2
        self.half_size = (self.bound_max - self.bound_min) / 2
        self.rescale = opt.SDF.VolSDF.rescale
        self.define_network(opt)

    @property
    def dual_field(self) -> bool:
        return self.opt.Ablate_config.dual_field == True   # noqa: E712

    def define_network(self, opt):
        k_geo = get_layer_dims(opt.SDF.arch.layers)[-1][-1]
        if self.dual_field:
            self.embed_fn = get_Embedder(opt=opt, input_dim=3, input_choice="Hash")
            self.Geo_enc = Geometry(opt=opt, input_dim=self.embed_fn.out_dim, skip=opt.SDF.arch.skip,
                                    tf_init=opt.SDF.NN_Init.tf_init, layers=get_layer_dims(opt.SDF.arch.layers))
        self.embed_fn_v = get_Embedder(opt=opt, input_dim=3, input_choice="Fourier")
        input_enc_dim = 3 + self.embed_fn_v.out_dim + 3 + k_geo * (2 if self.dual_field else 1)
        self.k_geo = k_geo
        self.Rad_dec = Radiance(opt=opt, input_dim=input_enc_dim, skip=opt.SDF.arch.skip,
                                tf_init=opt.SDF.NN_Init.tf_init, layers=get_layer_dims(opt.RadF.arch.layers))

    # ------------------------------------------------------------------ kernel plumbing
    def rad_spec(self) -> ops.RadSpec:
        return ops.RadSpec(self.embed_fn_v.N_freqs, self.k_geo, self.k_geo if self.dual_field else 0)

    def field_spec(self) -> ops.FieldSpec:
        # Geo_enc shares the SDF field's layout but its raw outputs are used as features (no sign / scale)
        return ops.FieldSpec(self.embed_fn.embedder_obj.grid, [float(x) for x in self.opt.data.bound_min],
                             [float(x) for x in self.opt.data.bound_max], self.Geo_enc.dims(), float(self.rescale),
                             100.0, 20.0, 1.0, 1.0)

    # ------------------------------------------------------------------ reference surface
    def Geometry_feat(self, xyz):
        shp = xyz.shape[:-1]
        flat = xyz.reshape(-1, 3).float()
        _, y, _, _ = ops.FieldEval.apply(self.field_spec(), None, self.embed_fn.embedder_obj.params, self.Geo_enc.theta(),
                                         None, None, None, flat, None, None, None, 0, None, True, False)
        return y.view(*shp, -1)

    def infer_embed_v(self, ray_utils):
        return self.embed_fn_v(ray_utils)

    def infer_app(self, geo_enc):
        """rgbs [...,3] from an explicit [x, n, view_enc, geo] input (API compatibility; Renderer.forward fuses
        this into the field kernel)."""
        return self.Rad_dec(geo_enc)

    forward = infer_app
