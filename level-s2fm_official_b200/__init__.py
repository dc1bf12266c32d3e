# Source directory of the ``levels2fm_b200`` package (see ../levels2fm_b200/__init__.py).
