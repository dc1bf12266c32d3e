"""Importable alias of the ``level-s2fm_official_b200/`` source directory.

The directory name required by the repo layout contains a hyphen and cannot be imported
directly; this package simply points its ``__path__`` at it, so
``import levels2fm_b200.models.SDF`` loads ``level-s2fm_official_b200/models/SDF.py``.
"""
import os as _os

_SRC = _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))), "level-s2fm_official_b200")
__path__.insert(0, _SRC)
__version__ = "0.1.0"
