"""CPU oracle for the Level-S2fM per-ray SDF volume-rendering hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is product code: only
``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl
reference`` legs of ``bench.py`` may import it, and only as the checker or the
timed CPU baseline.  The product path (``level-s2fm_official_b200``) never
imports this package and raises if its CUDA library is missing.

Layout
------
``hashgrid.py``   pure-PyTorch, double-differentiable restatement of the
                  tiny-cuda-nn 1.7 multiresolution hash grid (``tcnn.Encoding``
                  with otype=Grid/type=Hash/interpolation=Linear).  [EXT]
``aabb.py``       restatement of ``vren.ray_aabb_intersect`` (ngp_pl csrc). [EXT]
``port.py``       functional restatement of the reference's
                  ``models/{Renderer,SDF,RadF,base}.py`` -- this is the oracle
                  that travels to the GPU box.
``ref_shim.py``   imports the UNMODIFIED reference python from /root/reference
                  with stub modules for its missing third-party imports.  Only
                  usable in the build container; used to validate ``port.py``
                  and to generate ``tests/golden/*.npz`` (``make_golden.py``).
``hashgrid_ref.c``scalar C restatement of the hash-grid index/weight arithmetic
                  (fmaf / floorf / uint32 casts) used to pin the integer part.

PARITY PIN STATUS: the reference ships no tests, golden vectors or fixtures
(SURVEY.md section 4).  ``port.py`` is pinned against the reference's own Python
(run here through ``ref_shim.py``; fixtures committed under ``tests/golden``).
The two third-party native ops the reference calls (tiny-cuda-nn 1.7 hash grid,
vren 2.0 ray/AABB) are absent from /root/reference and cannot be installed:
for those two ops PARITY IS UNPINNED -- they are restated from the published
algorithms and anchored on the reference's call sites
(models/base.py:17,37; utils/custom_functions.py:31).
"""
