"""Drop-in replacement of the reference's ``models`` package for the render hot path.

``pipelines/LevelS2fM.py:38-43`` of the reference builds its fields with
``importlib.import_module("models.SDF").SDF(opt)``, ``models.RadF.RadF(opt)`` and
``models.Renderer.Renderer(opt)``; this package exports the same module and class names with the
same constructor/method signatures and state-dict layout, backed by the sm_100a kernels.
"""
