"""The C-ABI library: builds for sm_100a, loads, exports every symbol include/ls2fm.h declares, host-side helpers and
argument validation work without a GPU (no compute calls here)."""
import ctypes as C
import os
import re

import pytest
import torch

from oracle import hashgrid, port

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    import __graft_entry__ as ge
    ge.build()
    from levels2fm_b200 import _C
    return _C.get()


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "ls2fm.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(ls2fm_[A-Za-z0-9_]+)\s*\(", src)))


def test_every_declared_symbol_is_exported_and_bound(lib):
    from levels2fm_b200 import _C
    syms = declared_symbols()
    assert len(syms) >= 12
    for s in syms:
        assert hasattr(lib.dll, s), f"{s} not exported"
    assert sorted(_C.SIGNATURES) == syms, "ctypes binding and header disagree"


def test_library_is_sm100a_sass(lib):
    import shutil
    import subprocess
    cuobjdump = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not os.path.exists(cuobjdump):
        pytest.skip("cuobjdump not available")
    out = subprocess.run([cuobjdump, "-lelf", lib.path], capture_output=True, text=True).stdout
    assert "sm_100a" in out, out


@pytest.mark.parametrize("half,n_levels", [(1.0, 16), (2.0, 16), (5.0, 16), (1.0, 4)])
def test_grid_meta_is_the_oracles_table(lib, half, n_levels):
    b = port.SceneCfg(bound_min=(-half,) * 3, bound_max=(half,) * 3, n_levels=n_levels).per_level_scale
    levels, n_entries = lib.grid_meta(n_levels, 2, 19, 16, b)
    meta = hashgrid.grid_meta(n_levels, 2, 19, 16, b)
    assert n_entries == meta.n_entries
    for a, o in zip(levels, meta.levels):
        assert (a.scale, a.resolution, a.offset, a.size, bool(a.hashed)) == \
               (C.c_float(o.scale).value, o.resolution, o.offset, o.size, o.hashed)


def test_argument_errors_are_reported_not_crashed(lib):
    from levels2fm_b200 import _C
    f = _C.Field()
    p = _C.Points()
    rc = lib.dll.ls2fm_field_forward(C.byref(f), C.byref(p), None, None, None, None, None, None)
    assert rc != 0 and b"NULL" in lib.dll.ls2fm_last_error()
    with pytest.raises(RuntimeError):
        lib.check(rc)
    cfg = _C.GridCfg(40, 2, 19, 16, 1.5)
    assert lib.dll.ls2fm_grid_meta(C.byref(cfg), (_C.Level * 16)(), None) != 0


def test_no_cpu_fallback(lib):
    from levels2fm_b200 import ops
    with pytest.raises(RuntimeError, match="CUDA"):
        ops.sample_uniform_raw(lib, torch.zeros(4, 3), torch.ones(4, 3), 8, [-1, -1, -1], [1, 1, 1])


def test_missing_library_fails_loudly(tmp_path):
    from levels2fm_b200 import _C
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        _C.Lib(str(tmp_path / "nope.so"))


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "level-s2fm_official_b200")
    for dp, _, files in os.walk(pkg):
        for fn in files:
            if fn.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dp, fn)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", txt, flags=re.M), fn
                assert "hostsim import" not in txt and "from tests" not in txt, fn


def test_new_entry_points_validate_arguments(lib):
    """Round-2 entry points: NULL / out-of-range arguments are reported through the error string, empty inputs are no-ops
    (host-side checks only: nothing is launched without a GPU)."""
    d = lib.dll
    assert d.ls2fm_se3_to_SE3(None, 0, None, None) == 0
    assert d.ls2fm_se3_to_SE3(None, 3, None, None) != 0 and b"se3_to_SE3" in d.ls2fm_last_error()
    assert d.ls2fm_se3_to_SE3_backward(None, 2, None, None, None) != 0
    org = (C.c_double * 3)(0.0, 0.0, 0.0)
    assert d.ls2fm_grid_points(8, 0.1, org, 0, 0, None, None) == 0
    assert d.ls2fm_grid_points(8, 0.1, org, 500, 100, None, None) != 0 and b"grid_points" in d.ls2fm_last_error()
    assert d.ls2fm_grid_points(1, 0.1, org, 0, 1, None, None) != 0
    assert d.ls2fm_reproj_loss(None, None, None, None, None, 5, 0.1, 1e-6, None, None, None, None, None, None) != 0
    assert b"reproj_loss" in d.ls2fm_last_error()
    assert d.ls2fm_render_tail(None, None, None, None, None, None, 4, 8, 1, 1.0, 1.0, 1.0, None, None, None, None, None, None, None, None) != 0
    assert b"render_tail" in d.ls2fm_last_error()
