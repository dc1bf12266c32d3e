"""Oracle: multiresolution hash grid (tiny-cuda-nn 1.7 ``GridEncoding``). [EXT]

TEST INFRASTRUCTURE -- see oracle/__init__.py.  PARITY UNPINNED for this op:
tiny-cuda-nn (pinned ``tinycudann==1.7`` in the reference's env.yaml:241) is
not under /root/reference; this restates its published algorithm
 and is anchored on the reference's call sites
``models/base.py:17`` (construction) and ``models/base.py:37`` (forward), with
the config the reference builds at ``models/base.py:124-139``.

Everything is written with differentiable torch ops (index_select + arithmetic)
so autograd supplies first AND second order derivatives, as the reference needs
for ``SDF.gradient`` (models/SDF.py:102-114, create_graph=True).
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import List

import numpy as np
import torch
import torch.nn as nn

PRIME1 = 2654435761
PRIME2 = 805459861
MASK32 = 0xFFFFFFFF


@dataclass
class Level:
    scale: float        # exact float32 value, held as python float
    resolution: int
    offset: int         # in entries (not floats)
    size: int           # hashmap_size of this level, in entries
    hashed: bool


@dataclass
class GridMeta:
    n_levels: int
    n_features: int
    log2_hashmap_size: int
    base_resolution: int
    per_level_scale: float
    levels: List[Level]
    n_entries: int

    @property
    def n_params(self) -> int:
        return self.n_entries * self.n_features

    @property
    def n_output_dims(self) -> int:
        return self.n_levels * self.n_features


def grid_meta(n_levels: int, n_features: int, log2_hashmap_size: int,
              base_resolution: int, per_level_scale: float) -> GridMeta:
    """Per-level (scale, resolution, offset, size) in float32 arithmetic.

    scale_l = exp2f(l * log2f(b)) * N_min - 1 ; res_l = ceilf(scale_l) + 1 ;
    params_l = min(next_multiple_of_8(min(res^3, 2^31-1)), 2^log2_T).
    """
    f32 = np.float32
    b = f32(per_level_scale)                 # the json/py float lands in a C float
    # log2f / exp2f evaluated in float64 and rounded once to float32 (== a correctly rounded
    # libm); tests/test_oracle.py checks this table against oracle/hashgrid_ref.c (glibc) and
    # tests/test_cabi.py against the product's ls2fm_grid_meta -- one shared table.
    log2b = f32(math.log2(float(b)))
    levels: List[Level] = []
    offset = 0
    for l in range(n_levels):
        e = f32(2.0 ** float(f32(f32(l) * log2b)))
        scale = f32(e * f32(base_resolution)) - f32(1.0)
        scale = f32(scale)
        res = int(np.ceil(scale)) + 1
        max_params = (2 ** 32 - 1) // 2
        dense = res ** 3
        params = max_params if float(f32(res) ** 3) > float(max_params) else dense
        params = (params + 7) // 8 * 8
        params = min(params, 1 << log2_hashmap_size)
        # "hashed" as decided inside grid_index(): stride after the dense loop
        stride, d = 1, 0
        while d < 3 and stride <= params:
            stride *= res
            d += 1
        hashed = params < stride
        levels.append(Level(float(scale), res, offset, params, hashed))
        offset += params
    return GridMeta(n_levels, n_features, log2_hashmap_size, base_resolution,
                    float(b), levels, offset)


def corner_indices(u: torch.Tensor, lvl: Level):
    """Integer part of A.1 for one level.

    u: [M,3] float32 (or float64).  Returns (idx [M,8] int64 -- entry index
    inside the level, before adding lvl.offset -- and the fractional position
    w [M,3], differentiable w.r.t. u).
    """
    if u.dtype == torch.float32:
        # fmaf(scale, u, 0.5f): the product of two float32 is exact in float64,
        # so one float64 add followed by the float32 rounding reproduces the
        # single-rounding fused result (up to a ~2^-29-probability double
        # rounding, irrelevant to floor()).
        p = (u.double() * lvl.scale + 0.5).float()
    else:
        p = u * lvl.scale + 0.5
    pf = torch.floor(p.detach())
    w = p - pf
    g = pf.to(torch.int64) & MASK32          # (uint32_t)(int)floorf(p)
    idx = []
    for c in range(8):
        q = [(g[:, d] + ((c >> d) & 1)) & MASK32 for d in range(3)]
        stride, index, d = 1, torch.zeros_like(q[0]), 0
        while d < 3 and stride <= lvl.size:
            index = (index + q[d] * stride) & MASK32
            stride *= lvl.resolution
            d += 1
        if lvl.size < stride:
            # int64 products wrap mod 2^64; only the low 32 bits are kept
            index = (q[0] ^ ((q[1] * PRIME1) & MASK32) ^ ((q[2] * PRIME2) & MASK32)) & MASK32
        idx.append(index % lvl.size)
    return torch.stack(idx, dim=1), w


def encode(u: torch.Tensor, table: torch.Tensor, meta: GridMeta) -> torch.Tensor:
    """u [M,3] in (nominally) [0,1]^3, table flat [n_params] -> [M, L*F].

    Output channel order: level-major, feature-minor.
    """
    F = meta.n_features
    tab = table.view(-1, F)
    outs = []
    for lvl in meta.levels:
        idx, w = corner_indices(u, lvl)
        acc = None
        for c in range(8):
            wc = None
            for d in range(3):
                f = w[:, d] if (c >> d) & 1 else (1.0 - w[:, d])
                wc = f if wc is None else wc * f
            val = tab.index_select(0, idx[:, c] + lvl.offset)          # [M,F]
            term = wc[:, None] * val
            acc = term if acc is None else acc + term
        outs.append(acc)
    return torch.cat(outs, dim=1)


class Encoding(nn.Module):
    """Stand-in for ``tinycudann.Encoding`` (the surface models/base.py uses).

    ``Encoding(n_input_dims, encoding_config)`` with config keys otype="Grid",
    type="Hash", n_levels, n_features_per_level, log2_hashmap_size,
    base_resolution, per_level_scale, interpolation="Linear"
    (reference models/base.py:130-139).  Attributes: ``n_output_dims``,
    ``params`` (flat fp32 Parameter, init U(-1e-4, 1e-4) [EXT]).
    Deviation from real tcnn (accepted): output is fp32, not fp16.
    """

    def __init__(self, n_input_dims, encoding_config, seed=1337, dtype=torch.float32):
        super().__init__()
        assert n_input_dims == 3
        cfg = dict(encoding_config)
        assert cfg.get("otype", "Grid") in ("Grid", "HashGrid")
        assert cfg.get("interpolation", "Linear") == "Linear"
        self.meta = grid_meta(int(cfg["n_levels"]), int(cfg["n_features_per_level"]),
                              int(cfg["log2_hashmap_size"]), int(cfg["base_resolution"]),
                              float(cfg["per_level_scale"]))
        self.n_input_dims = 3
        self.n_output_dims = self.meta.n_output_dims
        g = torch.Generator().manual_seed(seed)
        p = (torch.rand(self.meta.n_params, generator=g, dtype=torch.float32) * 2 - 1) * 1e-4
        self.params = nn.Parameter(p.to(dtype))

    def forward(self, x):
        return encode(x, self.params, self.meta)
