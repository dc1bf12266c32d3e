"""ctypes binding of the C ABI in ``include/ls2fm.h`` (``libls2fm_sm100.so``).

The library is hand-written CUDA for sm_100a (``csrc/``), built in-tree by ``build.py``.
There is NO CPU path: every tensor handed to :class:`Lib` must live on a CUDA device and
:func:`get` raises if the library has not been built.  PyTorch owns all buffers; the
library only sees raw device pointers and the current CUDA stream.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional, Sequence

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
# LS2FM_LIB: an alternative build of the SAME sources (e.g. the -DLS_ABLATE timing experiment of tools/ablate.sh); never a CPU path
LIB_PATH = os.environ.get("LS2FM_LIB") or os.path.join(HERE, "lib", "libls2fm_sm100.so")

MAX_LEVELS = 16
MAX_LAYERS = 4
HIDDEN = 64
MAX_OUT = 20
MAX_RAD_IN = 68
ABI_VERSION = 2


class Level(C.Structure):
    _fields_ = [("scale", C.c_float), ("resolution", C.c_uint32), ("offset", C.c_uint32),
                ("size", C.c_uint32), ("hashed", C.c_uint32)]


class GridCfg(C.Structure):
    _fields_ = [("n_levels", C.c_int32), ("n_features", C.c_int32), ("log2_hashmap_size", C.c_int32),
                ("base_resolution", C.c_int32), ("per_level_scale", C.c_float)]


class Field(C.Structure):
    _fields_ = [("table", C.c_void_p), ("theta", C.c_void_p), ("n_levels", C.c_int32),
                ("levels", Level * MAX_LEVELS), ("bound_min", C.c_float * 3), ("bound_max", C.c_float * 3),
                ("rescale", C.c_float), ("n_layers", C.c_int32), ("dims", C.c_int32 * (MAX_LAYERS + 1)),
                ("softplus_beta", C.c_float), ("softplus_threshold", C.c_float),
                ("sdf_sign", C.c_float), ("scale_mlp", C.c_float), ("tc_image", C.c_void_p)]


class Points(C.Structure):
    _fields_ = [("xyz", C.c_void_p), ("center", C.c_void_p), ("ray", C.c_void_p), ("t", C.c_void_p),
                ("ray_index", C.c_void_p), ("n_active", C.c_void_p), ("n", C.c_int64),
                ("n_rays", C.c_int32), ("n_per_ray", C.c_int32), ("t_stride", C.c_int32), ("t_offset", C.c_int32),
                ("out_stride", C.c_int32), ("out_offset", C.c_int32)]


class ParamLayer(C.Structure):
    _fields_ = [("g", C.c_void_p), ("v", C.c_void_p), ("b", C.c_void_p), ("dg", C.c_void_p), ("dv", C.c_void_p), ("db", C.c_void_p),
                ("din", C.c_int32), ("dout", C.c_int32)]


class SamplerCfg(C.Structure):
    _fields_ = [("n_samples", C.c_int32), ("n_final", C.c_int32), ("max_upsample_iter", C.c_int32),
                ("max_bisection_itr", C.c_int32), ("eps", C.c_float), ("beta_speed", C.c_float)]


class Radiance(C.Structure):
    _fields_ = [("w_eff", C.c_void_p), ("b_eff", C.c_void_p), ("in_dim", C.c_int32), ("n_freq", C.c_int32),
                ("k_geo", C.c_int32), ("k_geo2", C.c_int32), ("geo2", C.c_void_p)]


class InputGrads(C.Structure):
    _fields_ = [("d_xyz", C.c_void_p), ("d_center", C.c_void_p), ("d_ray", C.c_void_p), ("d_t", C.c_void_p), ("workspace", C.c_void_p)]


_F3 = C.c_float * 3
_VP = C.c_void_p

# name -> (restype, argtypes); every symbol include/ls2fm.h declares
# (SamplerCfg is defined above Radiance)
SIGNATURES = {
    "ls2fm_abi_version": (C.c_int, []),
    "ls2fm_last_error": (C.c_char_p, []),
    "ls2fm_grid_meta": (C.c_int, [C.POINTER(GridCfg), C.POINTER(Level), C.POINTER(C.c_uint32)]),
    "ls2fm_smem_bytes": (C.c_int, [C.POINTER(Field), C.c_int, C.c_int]),
    "ls2fm_ray_aabb": (C.c_int, [_VP, _VP, C.c_int64, _F3, _F3, _VP, _VP, _VP]),
    "ls2fm_ray_aabb_backward": (C.c_int, [_VP, _VP, C.c_int32, C.c_int32, _F3, _F3, _VP, _VP, _VP, _VP, _VP]),
    "ls2fm_grid_encode": (C.c_int, [C.POINTER(Field), _VP, C.c_int64, _VP, _VP, _VP]),
    "ls2fm_grid_encode_backward": (C.c_int, [C.POINTER(Field), _VP, C.c_int64, _VP, _VP, _VP, _VP]),
    "ls2fm_grid_encode_tangent": (C.c_int, [C.POINTER(Field), _VP, C.c_int64, _VP, _VP, _VP, _VP, _VP, _VP]),
    "ls2fm_params_forward": (C.c_int, [C.POINTER(ParamLayer), C.c_int32, C.POINTER(ParamLayer), _VP, _VP, _VP, _VP]),
    "ls2fm_params_backward": (C.c_int, [C.POINTER(ParamLayer), C.c_int32, C.POINTER(ParamLayer), _VP, _VP, _VP, C.c_int32, _VP]),
    "ls2fm_field_image_floats": (C.c_int64, [C.POINTER(Field), C.POINTER(Radiance)]),
    "ls2fm_field_prepare": (C.c_int, [C.POINTER(Field), C.POINTER(Radiance), _VP, _VP]),
    "ls2fm_field_forward": (C.c_int, [C.POINTER(Field), C.POINTER(Points), C.POINTER(Radiance), _VP, _VP, _VP, _VP, _VP]),
    "ls2fm_field_forward_simt": (C.c_int, [C.POINTER(Field), C.POINTER(Points), C.POINTER(Radiance), _VP, _VP, _VP, _VP, _VP]),
    "ls2fm_field_forward_ws": (C.c_int, [C.POINTER(Field), C.POINTER(Points), _VP, _VP, _VP]),
    "ls2fm_field_backward_workspace_floats": (C.c_int64, [C.c_int64]),
    "ls2fm_field_backward": (C.c_int, [C.POINTER(Field), C.POINTER(Points), C.POINTER(Radiance)] + [_VP] * 11 + [C.POINTER(InputGrads), _VP]),
    "ls2fm_field_backward_simt": (C.c_int, [C.POINTER(Field), C.POINTER(Points), C.POINTER(Radiance)] + [_VP] * 11 + [C.POINTER(InputGrads), _VP]),
    "ls2fm_field_backward_tc": (C.c_int, [C.POINTER(Field), C.POINTER(Points), C.POINTER(Radiance)] + [_VP] * 11 + [C.POINTER(InputGrads), _VP]),
    "ls2fm_composite_forward": (C.c_int, [_VP] * 6 + [C.c_float, _F3, C.c_int32, C.c_int32] + [_VP] * 5),
    "ls2fm_composite_backward": (C.c_int, [_VP] * 6 + [C.c_float, _F3, C.c_int32, C.c_int32] + [_VP] * 10),
    "ls2fm_sample_uniform": (C.c_int, [_VP, _VP, C.c_int32, C.c_int32, _F3, _F3, _VP, _VP, _VP]),
    "ls2fm_sampler_workspace_bytes": (C.c_int64, [C.POINTER(SamplerCfg), C.c_int32]),
    "ls2fm_render_tail": (C.c_int, [_VP] * 6 + [C.c_int64, C.c_int32, C.c_int32, C.c_float, C.c_float, C.c_float] + [_VP] * 8),
    "ls2fm_grid_points": (C.c_int, [C.c_int32, C.c_double, C.c_double * 3, C.c_int64, C.c_int64, _VP, _VP]),
    "ls2fm_render_loss": (C.c_int, [_VP, _VP, C.c_int64, _VP, C.c_int64, C.c_float, C.c_float, _VP, _VP, _VP, _VP]),
    "ls2fm_generate_rays": (C.c_int, [_VP, _VP, _VP, C.c_int32, C.c_int64, _VP, _VP, _VP]),
    "ls2fm_generate_rays_backward": (C.c_int, [_VP, _VP, _VP, C.c_int32, C.c_int64, _VP, _VP, _VP, _VP]),
    "ls2fm_se3_to_SE3": (C.c_int, [_VP, C.c_int64, _VP, _VP]),
    "ls2fm_se3_to_SE3_backward": (C.c_int, [_VP, C.c_int64, _VP, _VP, _VP]),
    "ls2fm_reproj_loss": (C.c_int, [_VP] * 5 + [C.c_int64, C.c_float, C.c_float] + [_VP] * 6),
    "ls2fm_sphere_trace": (C.c_int, [C.POINTER(Field), _VP, _VP, C.c_int64, C.c_float, C.c_int32, _VP, _VP, _VP, _VP, _VP, _VP]),
    "ls2fm_sample_error_bounded": (C.c_int, [C.POINTER(Field), _VP, C.POINTER(SamplerCfg), _VP, _VP, C.c_int32, _VP, _VP, _VP, _VP, _VP]),
}


class Lib:
    """One loaded copy of the shared library."""
    _multi_device = None        # more than one CUDA device visible to this process (decided at the first launch)

    def __init__(self, path: str = LIB_PATH):
        if not os.path.exists(path):
            raise RuntimeError(
                f"{path} is missing: the sm_100a CUDA library has not been built "
                "(run `python -c 'import __graft_entry__ as g; g.build()'`). There is no CPU fallback.")
        self.path = path
        self.dll = C.CDLL(path)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(self.dll, name)          # AttributeError if the symbol is not exported
            fn.restype = res
            fn.argtypes = args
        if self.dll.ls2fm_abi_version() != ABI_VERSION:
            raise RuntimeError("libls2fm ABI version mismatch")

    # ---------------------------------------------------------------- plumbing
    def _check_device(self, t: torch.Tensor):
        if not t.is_cuda:
            raise RuntimeError("levels2fm_b200 kernels need CUDA tensors (no CPU fallback)")
        if self._multi_device is None:
            self._multi_device = torch.cuda.device_count() > 1
        if self._multi_device and t.device.index != torch.cuda.current_device():
            # (launches go to the CURRENT device's stream: one process per GPU, or torch.cuda.device(...) around the call)
            raise RuntimeError(f"tensor on cuda:{t.device.index} but the current device is cuda:{torch.cuda.current_device()}")

    def ptr(self, t: Optional[torch.Tensor], dtype=torch.float32):
        if t is None:
            return None
        if t.dtype != dtype:
            raise TypeError(f"expected {dtype}, got {t.dtype}")
        if not t.is_contiguous():
            raise ValueError("tensor must be contiguous")
        self._check_device(t)
        return C.c_void_p(t.data_ptr()) if t.numel() else None

    def stream(self):
        return C.c_void_p(torch.cuda.current_stream().cuda_stream)

    def check(self, rc: int):
        if rc != 0:
            raise RuntimeError("libls2fm: " + self.dll.ls2fm_last_error().decode())

    # ---------------------------------------------------------------- host helpers
    def grid_meta(self, n_levels, n_features, log2_hashmap_size, base_resolution, per_level_scale):
        cfg = GridCfg(n_levels, n_features, log2_hashmap_size, base_resolution, per_level_scale)
        levels = (Level * MAX_LEVELS)()
        n = C.c_uint32(0)
        self.check(self.dll.ls2fm_grid_meta(C.byref(cfg), levels, C.byref(n)))
        return [levels[i] for i in range(n_levels)], int(n.value)


_lib: Optional[Lib] = None


def get() -> Lib:
    """The product library (loaded once).  Raises if it has not been built."""
    global _lib
    if _lib is None:
        _lib = Lib(LIB_PATH)
    return _lib


def f3(v: Sequence[float]):
    return _F3(float(v[0]), float(v[1]), float(v[2]))
