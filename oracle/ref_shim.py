"""Import the UNMODIFIED reference python from /root/reference (build container only).

TEST INFRASTRUCTURE -- see oracle/__init__.py.  /root/reference does not exist on
the GPU box, so nothing reachable from ``pytest -m gpu``, ``smoke()`` or
``bench.py`` may call :func:`load`; it is used by ``oracle/make_golden.py`` (which
writes ``tests/golden/*.npz``) and by the container-only tests that validate
``oracle/port.py`` against the real reference code.

What is stubbed: the reference's missing third-party imports
``easydict, ipdb, termcolor, plyfile, skimage, open3d, torch_scatter`` (dummy
modules), ``tinycudann`` (-> oracle.hashgrid.Encoding) and ``vren``
(-> oracle.aabb.ray_aabb_intersect).  Reference files are never copied; they are
executed where they lie.

Defect handling mirrored from SURVEY.md 8(a):
 (ii)  error-bounded sampler typos: ``opt.VolSDF = opt.SDF.VolSDF``,
       ``max_bisection_itr = 10``, ``SDF.forward = SDF.infer_sdf``  (apply_c2_fixes)
 (iv)  ``sphere_tracing`` hard-codes ``.cuda()``: torch.Tensor.cuda -> identity.
"""
from __future__ import annotations

import contextlib
import os
import sys
import types

import torch

REFERENCE_ROOT = "/root/reference"


class EasyDict(dict):
    """Minimal easydict.EasyDict: attribute access, recursive on nested dicts."""

    def __init__(self, d=None, **kwargs):
        super().__init__()
        d = {} if d is None else dict(d)
        d.update(kwargs)
        for k, v in d.items():
            setattr(self, k, v)

    def __setattr__(self, name, value):
        if isinstance(value, (list, tuple)):
            value = type(value)(self.__class__(x) if isinstance(x, dict) and not isinstance(x, EasyDict) else x
                                for x in value)
        elif isinstance(value, dict) and not isinstance(value, EasyDict):
            value = self.__class__(value)
        super().__setattr__(name, value)
        super().__setitem__(name, value)

    __setitem__ = __setattr__

    def __getattr__(self, name):
        try:
            return self[name]
        except KeyError:
            raise AttributeError(name)

    def update(self, e=None, **f):
        d = e or dict()
        d.update(f)
        for k in d:
            setattr(self, k, d[k])


def available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "models"))


def _install_stubs():
    from . import aabb, hashgrid

    def mod(name, **attrs):
        m = types.ModuleType(name)
        m.__dict__.update(attrs)
        sys.modules[name] = m
        return m

    mod("easydict", EasyDict=EasyDict)
    mod("ipdb", set_trace=lambda *a, **k: None)
    mod("termcolor", colored=lambda s, *a, **k: str(s))
    mod("plyfile")
    mod("skimage")
    mod("open3d")
    mod("torch_scatter", segment_csr=lambda *a, **k: (_ for _ in ()).throw(NotImplementedError()))
    mod("tinycudann", Encoding=hashgrid.Encoding)
    mod("vren", ray_aabb_intersect=aabb.ray_aabb_intersect)


_loaded = None


def load():
    """Returns a namespace with the reference's modules (SDF, RadF, Renderer, base, camera, options)."""
    global _loaded
    if _loaded is not None:
        return _loaded
    if not available():
        raise RuntimeError("reference tree not present (only available in the build container)")
    _install_stubs()
    import warnings
    # the reference resolves `models`, `utils`, `options/*.json` relative to its root
    for name in ("models", "utils"):
        if name in sys.modules and not getattr(sys.modules[name], "__file__", "").startswith(REFERENCE_ROOT):
            raise RuntimeError(f"a non-reference module named '{name}' is already imported; "
                               "load the reference shim in a fresh process")
    sys.path.insert(0, REFERENCE_ROOT)
    cwd = os.getcwd()
    os.chdir(REFERENCE_ROOT)
    try:
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            import models.SDF as m_sdf
            import models.RadF as m_radf
            import models.Renderer as m_ren
            import models.base as m_base
            import utils.camera as u_camera
            import utils.options as u_options
            import utils.util as u_util
    finally:
        os.chdir(cwd)
        sys.path.remove(REFERENCE_ROOT)
    # defect (iv): .cuda() hard-coded in sphere_tracing; identity on a CPU-only box
    if not torch.cuda.is_available():
        torch.Tensor.cuda = lambda self, *a, **k: self
    _loaded = types.SimpleNamespace(SDF=m_sdf.SDF, RadF=m_radf.RadF, Renderer=m_ren.Renderer,
                                    base=m_base, camera=u_camera, options=u_options, util=u_util,
                                    EasyDict=EasyDict)
    return _loaded


_pipelines = None


def load_pipelines():
    """The reference's UNMODIFIED pipelines/Camera.py, BA.py, Point3D.py (callers of the hot path) next to the modules of load().
    ``utils.util_vis`` (matplotlib / imageio / trimesh plotting helpers, absent here) is replaced by an empty module: nothing
    on the render / loss path touches it."""
    global _pipelines
    if _pipelines is not None:
        return _pipelines
    ref = load()
    import importlib
    import warnings
    vis = types.ModuleType("utils.util_vis")
    sys.modules["utils.util_vis"] = vis
    sys.modules["utils"].util_vis = vis
    sys.path.insert(0, REFERENCE_ROOT)
    cwd = os.getcwd()
    os.chdir(REFERENCE_ROOT)
    try:
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            cam = importlib.import_module("pipelines.Camera")
            p3d = importlib.import_module("pipelines.Point3D")
            ba = importlib.import_module("pipelines.BA")
    finally:
        os.chdir(cwd)
        sys.path.remove(REFERENCE_ROOT)
    _pipelines = types.SimpleNamespace(Camera=cam, Point3D=p3d, BA=ba, ref=ref)
    return _pipelines


@contextlib.contextmanager
def in_reference_cwd():
    """The reference opens ``options/config_hash_sdf.json`` relative to its root (models/base.py:120)."""
    cwd = os.getcwd()
    os.chdir(REFERENCE_ROOT)
    try:
        yield
    finally:
        os.chdir(cwd)


def make_opt(yaml_name="DTU", device="cpu", **overrides):
    """Build the reference ``opt`` the way utils/options.py would, without its stdin prompts."""
    ref = load()
    with in_reference_cwd():
        opt = ref.options.load_options(f"options/{yaml_name}.yaml")
    opt.device = device
    if getattr(opt.data, "image_size", None) and opt.data.image_size[0]:
        opt.H, opt.W = opt.data.image_size
    if not getattr(opt.data, "scene", None):
        opt.data.scene = {"DTU": "scan24", "ETH3D": "courtyard", "bmvs": "scan1"}.get(yaml_name, "scene")
    if opt.data.scene not in opt.data:
        opt.data[opt.data.scene] = EasyDict()
    if not getattr(opt.data, "dataset", None):
        opt.data.dataset = yaml_name
    for dotted, v in overrides.items():
        node = opt
        keys = dotted.split(".")
        for k in keys[:-1]:
            node = node[k]
        node[keys[-1]] = v
    return opt


def apply_c2_fixes(opt):
    """Defect (ii): what the dormant error-bounded sampler needs to run as intended."""
    ref = load()
    opt.VolSDF = opt.SDF.VolSDF
    if getattr(opt.SDF.VolSDF, "max_bisection_itr", None) is None:
        opt.SDF.VolSDF.max_bisection_itr = 10
    ref.SDF.forward = ref.SDF.infer_sdf
    return opt


def build_models(opt, hash_config=None):
    """Construct reference SDF / RadF / Renderer.  ``hash_config`` optionally overrides the
    json the reference reads (written to a temp file because the reference reads a path)."""
    import json
    import tempfile
    ref = load()
    if hash_config is not None:
        f = tempfile.NamedTemporaryFile("w", suffix=".json", delete=False)
        json.dump({"encoding": hash_config}, f)
        f.close()
        opt.SDF.Hash_config.config_file = f.name
    import warnings
    with in_reference_cwd(), warnings.catch_warnings():
        warnings.simplefilter("ignore")
        sdf = ref.SDF(opt)
        rad = ref.RadF(opt)
        ren = ref.Renderer(opt)
    return sdf, rad, ren
