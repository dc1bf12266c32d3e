"""Encoders and MLP parameter holders of the hot path (reference: models/base.py).

The classes keep the reference's names, constructor arguments and ``state_dict`` keys
(``embedder_obj.params``, ``mlp.{i}.{bias,weight_g,weight_v}``, ``mlp_radiance.{i}.*``).  On the hot path
they are *parameter holders*: ``SDF`` / ``RadF`` / ``Renderer`` hand their effective weights to the fused
sm_100a kernels (``levels2fm_b200.ops``) instead of calling ``forward`` layer by layer.
"""
from __future__ import annotations

import json
import math
import warnings

import numpy as np
import torch
import torch.nn as nn

from .. import ops


def _weight_norm(linear: nn.Linear) -> nn.Linear:
    # old-style weight norm: parameters weight_g [out,1], weight_v [out,in] (models/base.py:200,241)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        return nn.utils.weight_norm(linear)


def effective_layers(mlp: nn.ModuleList):
    """[(W_eff [out,in], bias [out])] with W_eff = g * v / ||v||_row, differentiable w.r.t. g and v."""
    out = []
    for layer in mlp:
        if hasattr(layer, "weight_g"):
            W = torch._weight_norm(layer.weight_v, layer.weight_g, 0)
        else:
            W = layer.weight
        out.append((W, layer.bias))
    return out


def _wn_triples(mlp: nn.ModuleList):
    """[(weight_g, weight_v, bias)] if every layer is weight-normed, else None."""
    out = []
    for layer in mlp:
        if not hasattr(layer, "weight_g"):
            return None
        out.append((layer.weight_g, layer.weight_v, layer.bias))
    return out


def prepare_params(geometry=None, radiance=None):
    """(theta, w_eff, b_eff) of a Geometry MLP and/or a Radiance decoder.  One fused launch (ops.ParamPrep: weight norm +
    packing + composition, with its own fused backward) when every layer is weight-normed and the decoder has the shipped
    in -> 64 -> 64 -> 3 shape; the layer-by-layer torch path otherwise."""
    geo = _wn_triples(geometry.mlp) if geometry is not None else []
    rad = _wn_triples(radiance.mlp_radiance) if radiance is not None else []
    rad_ok = radiance is None or (rad is not None and len(rad) == 3 and rad[0][1].shape[0] == 64 and
                                  tuple(rad[1][1].shape) == (64, 64) and tuple(rad[2][1].shape) == (3, 64) and rad[0][1].shape[1] <= 68)
    if geo is not None and rad_ok and (geo or rad):
        flat = [t for triple in geo for t in triple] + [t for triple in (rad or []) for t in triple]
        theta, w_eff, b_eff = ops.ParamPrep.apply(len(geo), *flat)
        return (theta if geometry is not None else None, w_eff if radiance is not None else None,
                b_eff if radiance is not None else None)
    theta = ops.pack_theta(effective_layers(geometry.mlp)) if geometry is not None else None
    w_eff, b_eff = ops.compose_affine(effective_layers(radiance.mlp_radiance)) if radiance is not None else (None, None)
    return theta, w_eff, b_eff


# --------------- Hash encoding (tcnn.Encoding replacement) -------------------------------
class Encoding(nn.Module):
    """Stand-in for ``tinycudann.Encoding`` (otype Grid / type Hash / Linear interpolation) with the surface
    the reference uses (models/base.py:17,37): ``n_output_dims``, flat fp32 ``params``, ``forward(x [M,3])``.
    Output is fp32 (tcnn's default is fp16 -- see DESIGN.md)."""

    def __init__(self, n_input_dims, encoding_config, seed=1337):
        super().__init__()
        assert n_input_dims == 3, "hash grid is built for 3-D inputs"
        cfg = dict(encoding_config)
        assert cfg.get("otype", "Grid") in ("Grid", "HashGrid") and cfg.get("interpolation", "Linear") == "Linear"
        self.grid = ops.GridSpec(int(cfg["n_levels"]), int(cfg["n_features_per_level"]), int(cfg["log2_hashmap_size"]),
                                 int(cfg["base_resolution"]), float(cfg["per_level_scale"])).resolve()
        self.n_input_dims = 3
        self.n_output_dims = self.grid.n_output_dims
        g = torch.Generator().manual_seed(seed)
        self.params = nn.Parameter((torch.rand(self.grid.n_params, generator=g) * 2 - 1) * 1e-4)

    def forward(self, x):
        return ops.GridEncode.apply(self.grid, self.params, x.reshape(-1, 3).float())


class Embedder_Hash(nn.Module):
    def __init__(self, kwargs, include_input=True, input_dim=3):
        super().__init__()
        self.embedder_obj = Encoding(n_input_dims=input_dim, encoding_config=kwargs)
        self.input_dim = input_dim
        self.out_dim = self.embedder_obj.n_output_dims + self.input_dim
        self.include_input = include_input

    def forward(self, input, bound_min, bound_max, rescale=1.0):
        """Unfused encoding [x / rescale, grid((x - bmin) / (bmax - bmin))]; the fused field kernels do this
        on chip -- this entry point exists for callers that want the encoding itself."""
        assert input.shape[-1] == self.input_dim
        norm_input = (input - bound_min.to(input.device)) / ((bound_max - bound_min).to(input.device))
        out = self.embedder_obj(norm_input.reshape(-1, 3))
        if self.include_input:
            out = torch.cat([input / rescale, out.view(*input.shape[:-1], -1)], dim=-1)
        return out


class Embedder_Fourier(nn.Module):
    """[v, sin(2^k v), cos(2^k v)]_k (models/base.py:43-97).  Per-ray, tiny; the fused kernels evaluate it on chip."""

    def __init__(self, input_dim, max_freq_log2, N_freqs, log_sampling=True, include_input=True,
                 periodic_fns=(torch.sin, torch.cos)):
        super().__init__()
        self.input_dim = input_dim
        self.include_input = include_input
        self.periodic_fns = periodic_fns
        self.N_freqs = N_freqs
        self.out_dim = (input_dim if include_input else 0) + input_dim * N_freqs * len(periodic_fns)
        if log_sampling:
            self.freq_bands = 2.0 ** torch.linspace(0.0, max_freq_log2, N_freqs)
        else:
            self.freq_bands = torch.linspace(2.0 ** 0.0, 2.0 ** max_freq_log2, N_freqs)

    def forward(self, input, bound_min=None, bound_max=None, rescale=1.0):
        """Same signature as the reference (models/base.py:75-97): the bounds are accepted and unused, the raw input is divided
        by ``rescale``."""
        assert input.shape[-1] == self.input_dim
        out = [input / rescale] if self.include_input else []
        for f in self.freq_bands.tolist():
            for fn in self.periodic_fns:
                out.append(fn(input * f))
        return torch.cat(out, dim=-1)


def get_Embedder(opt, input_dim=3, input_choice="Hash", choices=("Hash", "Fourier")):
    if input_choice not in choices:
        raise ValueError(f"Invalid input option. Valid choices are: {choices}")
    if input_choice == "Hash":
        with open(opt.SDF.Hash_config.config_file) as f:
            enc_cfg = json.load(f)["encoding"]
        L, N_min = enc_cfg["n_levels"], enc_cfg["base_resolution"]
        scale = (opt.data.bound_max[0] - opt.data.bound_min[0]) / 2
        # the json's per_level_scale is overridden so the finest level is 2048 * half-extent (models/base.py:128-129)
        b_ = np.exp(np.log(2048 * scale / N_min) / (L - 1))
        kwargs = {"otype": "Grid", "type": "Hash", "n_levels": L, "n_features_per_level": enc_cfg["n_features_per_level"],
                  "log2_hashmap_size": enc_cfg["log2_hashmap_size"], "base_resolution": N_min, "per_level_scale": b_,
                  "interpolation": "Linear"}
        return Embedder_Hash(kwargs=kwargs, input_dim=input_dim)
    return Embedder_Fourier(input_dim=input_dim, max_freq_log2=4 - 1, N_freqs=4, log_sampling=True, include_input=True,
                            periodic_fns=(torch.sin, torch.cos))


def get_layer_dims(layers):
    """utils/util.py:273-275 of the reference."""
    return list(zip(layers[:-1], layers[1:]))


# --------------- Geometry MLP -------------------------------
class Geometry(nn.Module):
    """Weight-normed softplus(beta=100) MLP with geometric (sphere) initialisation (models/base.py:164-217)."""

    def __init__(self, opt, input_dim, layers, skip=[], tf_init=True):
        super().__init__()
        if len(skip):
            raise NotImplementedError("skip connections are not used by any shipped config (options/*.yaml: skip: [])")
        self.mlp = nn.ModuleList()
        self.skip = skip
        bias = opt.SDF.NN_Init.bias
        for li, (k_in, k_out) in enumerate(layers):
            if li == 0:
                k_in = input_dim
            last = li == len(layers) - 1
            if last:
                k_out += 1
            linear = nn.Linear(k_in, k_out)
            if tf_init:
                with torch.no_grad():
                    if last:
                        nn.init.normal_(linear.weight, mean=math.sqrt(math.pi) / math.sqrt(layers[li][0]), std=0.0001)
                        nn.init.constant_(linear.bias, -bias)
                    elif li == 0:
                        nn.init.constant_(linear.bias, 0.0)
                        nn.init.constant_(linear.weight[:, 3:], 0.0)
                        nn.init.normal_(linear.weight[:, :3], 0.0, math.sqrt(2) / math.sqrt(k_out))
                    else:
                        nn.init.constant_(linear.bias, 0.0)
                        nn.init.normal_(linear.weight, 0.0, math.sqrt(2) / math.sqrt(k_out))
            self.mlp.append(_weight_norm(linear))
        self.softplus = nn.Softplus(beta=100, threshold=20)

    def dims(self):
        return [self.mlp[0].weight_v.shape[1]] + [l.weight_v.shape[0] for l in self.mlp]

    def theta(self):
        """Packed effective parameters for the fused kernels (differentiable w.r.t. g, v, bias)."""
        return prepare_params(geometry=self)[0]

    def forward(self, points_enc):
        """Layer-by-layer evaluation of an already-built encoding (API compatibility; the hot path uses the
        fused field kernels through SDF.infer_sdf / RadF.Geometry_feat instead)."""
        feat = points_enc
        for li, (W, b) in enumerate(effective_layers(self.mlp)):
            feat = torch.nn.functional.linear(feat, W, b)
            if li <= len(self.mlp) - 2:
                feat = self.softplus(feat)
        return feat


# --------------- Radiance decoder -------------------------------
class Radiance(nn.Module):
    """Colour decoder (models/base.py:221-261).  The reference tests ``li <= len(self.mlp) - 2`` on an EMPTY
    ModuleList, so no hidden activation is ever applied: the decoder is affine o sigmoid.  Reproduced."""

    def __init__(self, opt, input_dim, layers, skip=[], tf_init=True):
        super().__init__()
        self.mlp = nn.ModuleList()          # stays empty, as in the reference
        self.skip = skip
        self.mlp_radiance = nn.ModuleList()
        for li, (k_in, k_out) in enumerate(layers):
            if li == 0:
                k_in = input_dim
            linear = nn.Linear(k_in, k_out)
            if tf_init:
                linear = _weight_norm(linear)
            self.mlp_radiance.append(linear)
        self.sigmoid = nn.Sigmoid()

    def effective_affine(self):
        _, w_eff, b_eff = prepare_params(radiance=self)
        return w_eff, b_eff

    def forward(self, geo_enc):
        W, b = self.effective_affine()
        return self.sigmoid(torch.nn.functional.linear(geo_enc, W, b))
